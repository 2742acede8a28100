#!/bin/bash
# round 2, session 7: config 5 breakdown (host phases, per-kernel device time), single-wait bench path
mkdir -p gpurun_out/r2_07
O=gpurun_out/r2_07
timeout 300 python -m pytest tests/test_gpu_supertree.py -q -x 2>&1 | tail -5
B200ST_TIMING=1 timeout 600 python bench.py --workload super_500part_100taxa --steps 10 --warmup 3 > $O/cfg5.json 2> $O/cfg5.err; grep b200st $O/cfg5.err
python -c "
import json; d=json.load(open('$O/cfg5.json')); print('cfg5: evals/s %.1f ms %.3f launches/step %.0f' % (d['value'], d['ms_per_step'], d['gpu_launches']/d['steps']))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file $O/launches_cfg5.csv python bench.py --workload super_500part_100taxa --steps 3 --warmup 3 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_cfg5.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-70s %s %s" % (r[ki][:70], r[vi], r[ui]))
PY
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 400 $O/bench_default.err
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("default: value %.2f e2e %.2f launch_ms %.3f logl %.7f" % (d["value"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["result"]["logl"]))
s=d["secondary"]; print("secondary: value %.2f e2e %.2f launch_ms %.3f logl %.7f" % (s["value"], s["e2e"]["value"], s["roofline"]["launch_ms"], s["result"]["logl"]))
PY
