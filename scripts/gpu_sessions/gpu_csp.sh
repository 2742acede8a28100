#!/bin/bash
mkdir -p gpurun_out/s8
for n in 12544 4096 1024; do for v in 0 4; do
  timeout 300 python bench.py --workload aa_200x100k_lg_g4 --patterns $n --steps 10 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/s8/csp_${n}_$v.json 2> gpurun_out/s8/csp_${n}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/csp_${n}_$v.json"))
    print("patterns $n variant $v (0 = category split when small, 4 = whole blocks per warp): launch_ms %.3f evals/s %.1f logl %.6f" % (d["roofline"]["launch_ms"], d["value"], d["result"]["logl"]))
except Exception as ex:
    print("failed", ex, open("gpurun_out/s8/csp_${n}_$v.err").read()[-400:])
PY
done; done
python scripts/bench_sweep.py --taxa 100 --patterns 20000 --kind aa --no-reference | cut -c1-200
python scripts/bench_sweep.py --taxa 100 --patterns 2000 --kind aa --no-reference | cut -c1-200
HT=8 bash scripts/gpu_sessions/gpu_super.sh
