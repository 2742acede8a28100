#!/bin/bash
# round 2, session 11: why is k_walk_gen<8> slow on the codon workload — launch list + full ncu capture
mkdir -p gpurun_out/r2_11
O=gpurun_out/r2_11
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/launches_codon.csv python bench.py --workload codon_100x20k_rev61_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_codon.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-70s %s %s" % (r[ki][:70], r[vi], r[ui]))
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_walk_gen -s 3 -c 1 -o $O/walkgen python bench.py --workload codon_100x20k_rev61_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -5 $O/ncu.log
ls -la $O
