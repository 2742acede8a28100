#!/bin/bash
# round 2, session 8: config 5 steady-state breakdown, RELL tests
mkdir -p gpurun_out/r2_08
O=gpurun_out/r2_08
timeout 600 python -m pytest tests/test_gpu_supertree.py tests/test_gpu_rell.py -q 2>&1 | tail -8
B200ST_TIMING=1 timeout 600 python bench.py --workload super_500part_100taxa --steps 20 --warmup 3 > $O/cfg5.json 2> $O/cfg5.err; grep b200st $O/cfg5.err
python -c "
import json; d=json.load(open('$O/cfg5.json')); print('cfg5: evals/s %.1f ms %.3f launches/step %.0f' % (d['value'], d['ms_per_step'], d['gpu_launches']/d['steps']))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 44 --csv --log-file $O/launches_cfg5.csv python bench.py --workload super_500part_100taxa --steps 3 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_cfg5.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-60s %s %s" % (r[ki][:60], r[vi], r[ui]))
PY
