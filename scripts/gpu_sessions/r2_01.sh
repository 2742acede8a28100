#!/bin/bash
# round 2, session 1: state-space 20-state walker — parity suite, then config 3 with both representations
mkdir -p gpurun_out/r2_01
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py 2>&1 | tail -40 > gpurun_out/r2_01/pytest.log
tail -15 gpurun_out/r2_01/pytest.log
for mode in 0 1; do
  B200LK_EIGEN20=$mode timeout 300 python bench.py --workload aa_200x100k_lg_g4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_01/cfg3_eigen$mode.json 2> gpurun_out/r2_01/cfg3_eigen$mode.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_01/cfg3_eigen$mode.json"))
    print("EIGEN20=$mode: evals/s %.1f launch_ms %.3f logl %.6f df %.6e" % (d["value"], d["roofline"]["launch_ms"], d["result"]["logl"], d["result"]["df"]))
except Exception as ex:
    print("EIGEN20=$mode failed", ex, open("gpurun_out/r2_01/cfg3_eigen$mode.err").read()[-800:])
PY
done
