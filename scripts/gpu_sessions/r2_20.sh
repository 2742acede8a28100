#!/bin/bash
# round 2, session 20: k_site_walk20 (site-specific 20-state partials on the FP64-MMA pipe, matrices resident in shared memory)
mkdir -p gpurun_out/r2_20
O=gpurun_out/r2_20
timeout 900 python -m pytest tests/test_gpu_site.py -q -x 2>&1 | tail -25
timeout 600 python bench.py --workload aa_100x20k_pmsf_g4 --steps 5 --warmup 3 > $O/site.json 2> $O/site.err; echo "site rc=$?"; tail -c 800 $O/site.err
python - <<PY
import json
d=json.load(open("$O/site.json"))
print("site value %.3f e2e %.3f ms %.3f logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["result"]["logl"]), json.dumps(d.get("roofline",{}))[:600])
print(d.get("cpu_baseline"))
PY
