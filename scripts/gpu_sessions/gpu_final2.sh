#!/bin/bash
# last evidence refresh of the session: bench lines of the final code (configs 2, 3, 5; reference arm of config 3), traffic + full ncu of the 20-state walker
O=gpurun_out/s8i; mkdir -p $O
timeout 300 python bench.py > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --workload aa_200x100k_lg_g4 > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 300 python bench.py --workload super_500part_100taxa --host-threads 8 > $O/bench_cfg5.json 2> $O/bench_cfg5.err
timeout 300 python bench.py --workload aa_200x100k_lg_g4 --impl reference --steps 3 --warmup 1 > $O/bench_ref_cfg3.json 2> $O/bench_ref_cfg3.err
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload aa_200x100k_lg_g4"
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_walk_n20_mma -s 2 -c 1 --csv --log-file $O/traffic_cfg3.csv $B > $O/traffic_cfg3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20_mma -s 2 -c 1 -o $O/n20_full $B > $O/n20_full.log 2>&1
for f in bench_cfg2 bench_cfg3 bench_cfg5 bench_ref_cfg3; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json"))
    print("$f", "value %.3f" % d["value"], "e2e %.3f" % d["e2e"]["value"], "frac", d.get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex:
    print("$f failed", ex, open("$O/$f.err").read()[-300:])
PY
done
