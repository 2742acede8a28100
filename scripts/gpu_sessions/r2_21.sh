#!/bin/bash
# round 2, session 21: k_site_walk20 with 16 warps per CTA (default) against 8 (variant 6); ncu of the kernel
mkdir -p gpurun_out/r2_21  # (second run of this script: offsets hoisted out of the op loop)
O=gpurun_out/r2_21
timeout 900 python -m pytest tests/test_gpu_site.py -q -x 2>&1 | tail -5
for v in 0 6; do
timeout 600 python bench.py --workload aa_100x20k_pmsf_g4 --steps 5 --warmup 3 --variant $v --no-cpu-baseline > $O/site_v$v.json 2> $O/site_v$v.err; echo "site v$v rc=$?"; tail -c 300 $O/site_v$v.err
python - <<PY
import json
d=json.load(open("$O/site_v$v.json"))
print("site v$v value %.3f e2e %.3f ms %.3f launch_ms %.3f logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["result"]["logl"]))
PY
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_site_walk20 -s 3 -c 1 -f -o $O/site_walk20_b python bench.py --workload aa_100x20k_pmsf_g4 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O
