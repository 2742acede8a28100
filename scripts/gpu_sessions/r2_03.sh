#!/bin/bash
# round 2, session 3: second-generation 20-state walker (TMA rings): parity under a watchdog, then timing and ncu
mkdir -p gpurun_out/r2_03
O=gpurun_out/r2_03
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mem_save.py -q -x 2>&1 | tail -25 > $O/pytest_parity.log
tail -12 $O/pytest_parity.log
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -25 > $O/pytest_all.log
tail -8 $O/pytest_all.log
for v in 0 1 3 5; do
  timeout 300 python bench.py --workload aa_200x100k_lg_g4 --steps 10 --warmup 3 --no-cpu-baseline --variant $v > $O/cfg3_v$v.json 2> $O/cfg3_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/cfg3_v$v.json"))
    print("variant $v: evals/s %.1f launch_ms %.3f logl %.6f" % (d["value"], d["roofline"]["launch_ms"], d["result"]["logl"]))
except Exception as ex:
    print("variant $v failed", ex, open("$O/cfg3_v$v.err").read()[-600:])
PY
done
timeout 300 python bench.py --workload aa_200x100k_lg_g4 --patterns 12544 --steps 10 --warmup 3 --no-cpu-baseline > $O/cfg3_small.json 2> $O/cfg3_small.err
python -c "
import json; d=json.load(open('$O/cfg3_small.json')); print('12.5k patterns (8-GPU shard): evals/s %.1f launch_ms %.3f' % (d['value'], d['roofline']['launch_ms']))"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20v2 -s 3 -c 1 -o $O/walk20v2 python bench.py --workload aa_200x100k_lg_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu.log 2>&1
echo "ncu rc=$?"
timeout 900 python bench.py --workload aa_100x50k_lg_c60_g4 --steps 3 --warmup 3 --no-cpu-baseline > $O/cfg4.json 2> $O/cfg4.err
python -c "
import json; d=json.load(open('$O/cfg4.json')); print('cfg4: evals/s %.2f launch_ms %.2f logl %.6f' % (d['value'], d['roofline']['launch_ms'], d['result']['logl']))" || tail -c 500 $O/cfg4.err
timeout 1200 python -m pytest tests/test_gpu_fullsize.py -q 2>&1 | tail -15
