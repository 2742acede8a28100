#!/bin/bash
# round 2, session 31 (8 GPUs): scaling lines of the default workload (config 2 + config 3 as `secondary`) with the final kernels
mkdir -p gpurun_out/r2_31
O=gpurun_out/r2_31
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
run 8 29511 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "n8 rc=$?"; tail -c 200 $O/bench_n8.err
run 4 29512 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_n4.json 2> $O/bench_n4.err; echo "n4 rc=$?"


python - <<PY
import json
def last(f):
    return json.loads([l for l in open(f) if l.startswith("{")][-1])
for f in ("bench_n4","bench_n8"):
    try:
        d=last("$O/%s.json"%f)
        c=d.get("check") or {}
        print(f, "value %.2f e2e %.2f ms %.3f launch_ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms"]), c.get("sum_of_ranks_equals_allreduce"), (c.get("slice_recomputed_on_rank0") or {}).get("max_rel_diff"))
        s=d.get("secondary")
        if s: print("   secondary value %.2f e2e %.2f" % (s["value"], s["e2e"]["value"]))
    except Exception as ex:
        print(f, "failed", ex)
PY
