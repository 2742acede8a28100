#!/bin/bash
# round 2, session 13 (2 GPUs): multi-device engines behind the C ABI, the sharded bench with its self-check, reference arm under torchrun
mkdir -p gpurun_out/r2_13
O=gpurun_out/r2_13
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi_device.py -q 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -c 400 $O/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload super_500part_100taxa --steps 10 --warmup 3 > $O/cfg5_n2.json 2> $O/cfg5_n2.err; echo "cfg5 n2 rc=$?"; tail -c 300 $O/cfg5_n2.err
python - <<PY
import json
for f in ("bench_n2","bench_ref_n2","cfg5_n2"):
    try:
        d=json.loads([l for l in open("$O/%s.json"%f) if l.startswith("{")][-1])
        print(f, "value %.3f e2e %.3f ms %.3f n_gpus %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["n_gpus"]), "cores", (d.get("cpu_baseline") or {}).get("cores"), json.dumps(d.get("check"))[:600])
        if "secondary" in d: s=d["secondary"]; print("  secondary value %.3f e2e %.3f ms %.3f" % (s["value"], s["e2e"]["value"], s["ms_per_step"]), json.dumps(s.get("check"))[:400])
    except Exception as ex: print(f, "failed", ex)
PY
