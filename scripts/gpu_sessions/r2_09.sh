#!/bin/bash
# round 2, session 9: general-n MMA walker parity; config 5 per-kernel times
mkdir -p gpurun_out/r2_09
O=gpurun_out/r2_09
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference.py tests/test_gpu_rell.py -q -x 2>&1 | tail -30 > $O/pytest_a.log
tail -25 $O/pytest_a.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_group -s 22 -c 33 --csv --log-file $O/launches_cfg5.csv python bench.py --workload super_500part_100taxa --steps 3 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_cfg5.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-60s %s %s" % (r[ki][:60], r[vi], r[ui]))
PY
