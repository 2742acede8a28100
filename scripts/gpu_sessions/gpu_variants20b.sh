#!/bin/bash
# steady-state cost per pattern of each variant at pattern counts that fill whole waves
mkdir -p gpurun_out/s8
run() { v=$1; n=$2;
  timeout 300 python bench.py --workload aa_200x100k_lg_g4 --patterns $n --steps 10 --warmup 3 --no-cpu-baseline --variant $v > gpurun_out/s8/w20_${v}_$n.json 2> gpurun_out/s8/w20_${v}_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/w20_${v}_$n.json"))
    ms=d["roofline"]["launch_ms"]
    print("variant $v patterns $n: launch_ms %.3f  ns/pattern %.2f  frac %.3f" % (ms, ms*1e6/$n, d["roofline"]["frac"]))
except Exception as ex:
    print("variant $v failed", ex, open("gpurun_out/s8/w20_${v}_$n.err").read()[-500:])
PY
}
run 0 56832; run 0 113664; run 0 170496
run 5 75776; run 5 151552
run 6 75776; run 6 151552
run 8 75776
run 7 85248
run 9 94720
