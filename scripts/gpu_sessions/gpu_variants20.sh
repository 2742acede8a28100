#!/bin/bash
# A/B of the 20-state walker variants on config 3 (and a C60 slice): kernel ms, logL
mkdir -p gpurun_out/s8
for v in ${VARIANTS:-0 5 6 7 8 9}; do
  timeout 120 python bench.py --workload aa_200x100k_lg_g4 --steps 10 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/s8/v20_$v.json 2> gpurun_out/s8/v20_$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/v20_$v.json"))
    print("variant $v cfg3: launch_ms %.3f frac %.3f evals/s %.1f logl %.6f df %.6f" % (d["roofline"]["launch_ms"], d["roofline"]["frac"], d["value"], d["result"]["logl"], d["result"]["df"]))
except Exception as ex:
    print("variant $v failed", ex, open("gpurun_out/s8/v20_$v.err").read()[-500:])
PY
done
