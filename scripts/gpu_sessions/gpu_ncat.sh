#!/bin/bash
mkdir -p gpurun_out/s8
for nc in 1 2 4 8 16; do
  n=$((454656/nc))   # 454656 = 8 x 56832: whole waves at every ncat
  timeout 300 python bench.py --workload aa_200x100k_lg_g4 --ncat $nc --patterns $n --steps 5 --warmup 3 --no-cpu-baseline --variant ${V:-0} > gpurun_out/s8/nc_$nc.json 2> gpurun_out/s8/nc_$nc.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/nc_$nc.json"))
    ms=d["roofline"]["launch_ms"]
    print("ncat $nc patterns $n: launch_ms %.3f  ns/(pattern*cat) %.2f  frac %.3f" % (ms, ms*1e6/($n*$nc), d["roofline"]["frac"]))
except Exception as ex:
    print("ncat $nc failed", ex, open("gpurun_out/s8/nc_$nc.err").read()[-400:])
PY
done
