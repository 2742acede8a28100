#!/bin/bash
# end-of-session evidence: suites, smoke, bench lines (both arms), launch lists of the bench command, ncu of both walkers
O=gpurun_out/s8h; mkdir -p $O
( time python -m pytest tests -m gpu -q 2>&1 | tail -3 ) > $O/pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_cfg2.json 2> $O/bench_ref_cfg2.err
timeout 600 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 600 python bench.py --workload aa_200x100k_lg_g4 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 900 python bench.py --workload aa_100x50k_lg_c60_g4 --steps 5 --no-cpu-baseline > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 600 python bench.py --workload super_500part_100taxa --host-threads 8 > $O/bench_cfg5.json 2> $O/bench_cfg5.err
timeout 600 python bench.py --workload super_500part_100taxa --impl reference --steps 2 --warmup 1 > $O/bench_ref_cfg5.json 2> $O/bench_ref_cfg5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_cfg2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload aa_200x100k_lg_g4 > $O/launches_cfg3.log 2>&1
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_walk_n4c4 -s 2 -c 1 --csv --log-file $O/traffic_cfg2.csv $B > $O/traffic_cfg2.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_walk_n20_mma -s 2 -c 1 --csv --log-file $O/traffic_cfg3.csv $B --workload aa_200x100k_lg_g4 > $O/traffic_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20_mma -s 2 -c 1 -o $O/n20_full $B --workload aa_200x100k_lg_g4 > $O/n20_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n4c4 -s 2 -c 1 -o $O/n4_full $B --patterns 262144 > $O/n4_full.log 2>&1
cat $O/pytest_gpu.log $O/smoke.log
for f in bench_ref_cfg2 bench_cfg2 bench_cfg3 bench_cfg4 bench_cfg5 bench_ref_cfg5; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json"))
    print("$f", "value %.3f" % d["value"], "e2e %.3f" % d["e2e"]["value"], "frac", d.get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
except Exception as ex:
    print("$f failed", ex, open("$O/$f.err").read()[-300:])
PY
done
