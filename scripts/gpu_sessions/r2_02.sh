#!/bin/bash
# round 2, session 2: state-space walker variants + ncu, new tests (config 4/5 full size, NaN repair), new bench.py end to end
mkdir -p gpurun_out/r2_02
O=gpurun_out/r2_02
for v in 0 1 2 3; do
  timeout 300 python bench.py --workload aa_200x100k_lg_g4 --steps 10 --warmup 3 --no-cpu-baseline --variant $v > $O/cfg3_v$v.json 2> $O/cfg3_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/cfg3_v$v.json"))
    print("variant $v: evals/s %.1f launch_ms %.3f" % (d["value"], d["roofline"]["launch_ms"]))
except Exception as ex:
    print("variant $v failed", ex, open("$O/cfg3_v$v.err").read()[-600:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20ss -s 3 -c 1 -o $O/walk20ss python bench.py --workload aa_200x100k_lg_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1
echo "ncu rc=$?"
timeout 1500 python -m pytest tests/test_gpu_underflow.py tests/test_gpu_fullsize.py -q 2>&1 | tail -30 > $O/pytest_new.log
tail -12 $O/pytest_new.log
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 600 $O/bench_default.err
python - <<PY
import json
try:
    d=json.load(open("$O/bench_default.json"))
    print("default: value %.2f e2e %.2f frac %.3f launch_ms %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["launch_ms"]))
    s=d["secondary"]; print("secondary: value %.2f e2e %.2f roofline %s" % (s["value"], s["e2e"]["value"], json.dumps(s["roofline"])[:600]))
except Exception as ex:
    print("default bench parse failed", ex)
PY
for c in 1 8; do
  timeout 300 python bench.py --workload dna_1000x1M_gtr_g4 --patterns 262144 --ncat $c --steps 5 --warmup 3 --no-cpu-baseline > $O/dna_c$c.json 2> $O/dna_c$c.err
  python - <<PY
import json
try:
    d=json.load(open("$O/dna_c$c.json"))
    print("dna ncat $c: evals/s %.1f launch_ms %.3f kernel %s alg GB/s %.0f" % (d["value"], d["roofline"]["launch_ms"], d["roofline"]["kernel"], d["roofline"]["algorithmic_equivalent_GBps"]))
except Exception as ex:
    print("dna ncat $c failed", ex, open("$O/dna_c$c.err").read()[-600:])
PY
done
(time timeout 900 python bench.py --impl reference --steps 3 --warmup 1) > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 400 $O/bench_ref.err; head -c 1500 $O/bench_ref.json
