#!/bin/bash
# round 2, session 12: codon on the general-n walker (ring depth 2 for 8 state tiles): parity, bench, launch list, ncu
mkdir -p gpurun_out/r2_12
O=gpurun_out/r2_12
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "codon or n36 or n23 or n12" 2>&1 | tail -5
timeout 600 python bench.py --workload codon_100x20k_rev61_g4 --steps 5 --warmup 3 > $O/codon.json 2> $O/codon.err; echo "codon rc=$?"; tail -c 600 $O/codon.err
python - <<PY
import json
d=json.load(open("$O/codon.json"))
print("codon value %.3f e2e %.3f ms %.3f logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["result"]["logl"]), json.dumps(d.get("roofline",{}))[:1200])
print(d["cpu_baseline"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 20 --csv --log-file $O/launches_codon.csv python bench.py --workload codon_100x20k_rev61_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_codon.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-70s %s %s" % (r[ki][:70], r[vi], r[ui]))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_walk_gen -s 3 -c 1 -o $O/walkgen python bench.py --workload codon_100x20k_rev61_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.log
