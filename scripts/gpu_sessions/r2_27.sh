#!/bin/bash
# round 2, session 27: tip states through cp.async in the any-C 4-state walker (k_walk_n4<C>, k_group_walk_n4<C>): parity + G1 / G8 / config 5
mkdir -p gpurun_out/r2_27
O=gpurun_out/r2_27
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_supertree.py tests/test_gpu_mixlen.py -q -x 2>&1 | tail -3
for c in 1 8; do
timeout 600 python bench.py --patterns 262144 --ncat $c --steps 10 --warmup 3 --no-cpu-baseline > $O/dna_c$c.json 2> $O/dna_c$c.err; echo "c$c rc=$?"; tail -c 300 $O/dna_c$c.err
done
timeout 600 python bench.py --workload super_500part_100taxa --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg5.json 2> $O/cfg5.err; echo "cfg5 rc=$?"; tail -c 300 $O/cfg5.err
python - <<PY
import json
for f in ("dna_c1","dna_c8","cfg5"):
    d=json.loads([l for l in open("$O/%s.json"%f) if l.startswith("{")][-1])
    print(f, "value %.2f ms %.3f launch_ms %.3f" % (d["value"], d["ms_per_step"], d.get("roofline",{}).get("launch_ms",0) or 0))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $O/cfg2.json 2> $O/cfg2.err; echo "cfg2 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("$O/cfg2.json") if l.startswith("{")][-1])
print("cfg2 value %.2f ms %.3f launch_ms %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["launch_ms"]))
PY
