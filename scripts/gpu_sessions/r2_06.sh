#!/bin/bash
# round 2, session 6: grouped launches (config 5), 3-deep staging, category-split race fix
mkdir -p gpurun_out/r2_06
O=gpurun_out/r2_06
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_supertree.py tests/test_gpu_supertree_nni.py tests/test_gpu_supertree_unlinked.py tests/test_gpu_mem_save.py -q -x 2>&1 | tail -25 > $O/pytest_a.log
tail -6 $O/pytest_a.log
for v in 0 1 2 3 4; do
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
import json; d=json.load(open('$O/cfg3_small.json')); print('12.5k patterns (8-GPU shard): evals/s %.1f launch_ms %.3f logl %.6f' % (d['value'], d['roofline']['launch_ms'], d['result']['logl']))"
for g in 0 1; do
  if [ $g = 1 ]; then export B200ST_NO_GROUP=1; fi
  timeout 600 python bench.py --workload super_500part_100taxa --steps 10 --warmup 3 > $O/cfg5_nogroup$g.json 2> $O/cfg5_nogroup$g.err
  python - <<PY
import json
try:
    d=json.load(open("$O/cfg5_nogroup$g.json"))
    print("cfg5 NO_GROUP=$g: evals/s %.1f ms %.3f launches/step %.0f logl %.6f df %.6e" % (d["value"], d["ms_per_step"], d["gpu_launches"]/d["steps"], d["result"]["logl"], d["result"]["df"]))
except Exception as ex:
    print("cfg5 failed", ex, open("$O/cfg5_nogroup$g.err").read()[-800:])
PY
done
unset B200ST_NO_GROUP
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
