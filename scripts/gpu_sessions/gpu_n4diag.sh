#!/bin/bash
mkdir -p gpurun_out/s8
for v in ${VARIANTS:-0 6 7 8}; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --variant $v --patterns ${PATTERNS:-524288} > gpurun_out/s8/n4_v$v.json 2> gpurun_out/s8/n4_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/n4_v$v.json"))
    print("n4 variant $v: launch_ms %.3f frac %.3f logl %.6f clocks %s" % (d["roofline"]["launch_ms"], d["roofline"]["frac"], d["result"]["logl"], d["clocks"]))
except Exception as ex:
    print("variant $v failed", ex, open("gpurun_out/s8/n4_v$v.err").read()[-500:])
PY
done
