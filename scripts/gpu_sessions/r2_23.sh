#!/bin/bash
# round 2, session 23 (8 GPUs): the scaling lines of configs 2 (+ 3 as `secondary`) and 5, multi-device engines behind the C ABI
mkdir -p gpurun_out/r2_23
O=gpurun_out/r2_23
nvidia-smi -L | head -8
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
run 8 29511 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "n8 rc=$?"; tail -c 300 $O/bench_n8.err
run 4 29512 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n4.json 2> $O/bench_n4.err; echo "n4 rc=$?"
run 8 29513 --workload super_500part_100taxa --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg5_n8.json 2> $O/cfg5_n8.err; echo "cfg5 n8 rc=$?"; tail -c 300 $O/cfg5_n8.err
python - <<PY
import json
def last(f):
    return json.loads([l for l in open(f) if l.startswith("{")][-1])
for f in ("bench_n8","bench_n4","cfg5_n8"):
    try:
        d=last("$O/%s.json"%f)
        print(f, "value %.2f e2e %.2f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("self_check"))
        s=d.get("secondary")
        if s: print("   secondary value %.2f e2e %.2f" % (s["value"], s["e2e"]["value"]))
    except Exception as ex:
        print(f, "failed", ex)
PY
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_site.py -q -x -k "multi" 2>&1 | tail -5
