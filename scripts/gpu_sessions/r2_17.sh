#!/bin/bash
# round 2, session 17: site-specific bench line (PMSF-like 20 states + G4) against the reference's SITE_MODEL kernel; launch list
mkdir -p gpurun_out/r2_17
O=gpurun_out/r2_17
timeout 600 python bench.py --workload aa_100x20k_pmsf_g4 --steps 5 --warmup 3 > $O/site.json 2> $O/site.err; echo "site rc=$?"; tail -c 800 $O/site.err
python - <<PY
import json
d=json.load(open("$O/site.json"))
print("site value %.3f e2e %.3f ms %.3f logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["result"]["logl"]), json.dumps(d.get("roofline",{}))[:900])
print(d.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 16 --csv --log-file $O/launches_site.csv python bench.py --workload aa_100x20k_pmsf_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_site.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-70s %s %s" % (r[ki][:70], r[vi], r[ui]))
PY
