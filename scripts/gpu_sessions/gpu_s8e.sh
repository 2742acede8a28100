#!/bin/bash
mkdir -p gpurun_out/s8
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 )
for v in ${VARIANTS:-0 1 2 3}; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/s8/n4_v$v.json 2> gpurun_out/s8/n4_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/n4_v$v.json"))
    print("n4 variant $v cfg2: launch_ms %.3f frac %.3f evals/s %.2f logl %.6f df %.6f clocks %s" % (d["roofline"]["launch_ms"], d["roofline"]["frac"], d["value"], d["result"]["logl"], d["result"]["df"], d["clocks"]))
except Exception as ex:
    print("variant $v failed", ex, open("gpurun_out/s8/n4_v$v.err").read()[-500:])
PY
done
