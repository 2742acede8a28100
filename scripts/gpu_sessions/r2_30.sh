#!/bin/bash
# round 2, session 30: k_walk_n4c4 with per-op byte offsets precomputed by the info kernel: parity, bench, instruction count (ncu)
mkdir -p gpurun_out/r2_30
O=gpurun_out/r2_30
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_phylotree.py -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $O/cfg2_$i.json 2> $O/cfg2_$i.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("$O/cfg2_$i.json"))
print("cfg2 value %.3f e2e %.3f ms %.3f launch_ms %.3f clocks %s logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["clocks"]["sm_mhz"], d["result"]["logl"]))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_walk_n4c4 -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary 2>&1 | grep -E "gpu__time|inst_executed|issue_active"
