#!/bin/bash
# round 2, session 25 (second run): k_walk_n4c4 L2 bulk prefetches 4 ops ahead: variants 20 none, 21 tip rows, 22 sibling CLVs, 23 both
mkdir -p gpurun_out/r2_25
O=gpurun_out/r2_25
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "parity_small or deep" 2>&1 | tail -3
for v in 0 27 0 27; do
timeout 600 python bench.py --steps 10 --warmup 3 --variant $v --no-cpu-baseline --no-secondary > $O/cfg2_v$v.json 2> $O/cfg2_v$v.err; echo "v$v rc=$?"; tail -c 300 $O/cfg2_v$v.err
python - <<PY
import json
d=json.load(open("$O/cfg2_v$v.json"))
print("cfg2 v$v value %.3f e2e %.3f ms %.3f launch_ms %.3f clocks %s logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["clocks"]["sm_mhz"], d["result"]["logl"]))
PY
done
