#!/bin/bash
# round 2, session 28 (2 GPUs): multi-device engines with mixed branch lengths / site-specific models; N = 2 bench line + reference arm
mkdir -p gpurun_out/r2_28
O=gpurun_out/r2_28
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_site.py -q -x -k "multi" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"; tail -c 300 $O/bench_n2.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_n2.json") if l.startswith("{")][-1])
print("n2 value %.2f e2e %.2f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("check",{}).get("sum_of_ranks_equals_allreduce"))
s=d.get("secondary")
if s: print("   secondary value %.2f e2e %.2f" % (s["value"], s["e2e"]["value"]))
PY
