#!/bin/bash
mkdir -p gpurun_out/s8
for ht in ${HT:-0 1 4}; do
  timeout 600 python bench.py --workload super_500part_100taxa --steps 10 --warmup 3 --host-threads $ht > gpurun_out/s8/super_ht$ht.json 2> gpurun_out/s8/super_ht$ht.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/s8/super_ht$ht.json"))
    print("host_threads $ht: ms/step %.2f evals/s %.2f launches/step %d achieved %.0f GB/s logl %.4f df %.4f" % (d["ms_per_step"], d["value"], d["gpu_launches"]/d["steps"], d["roofline"]["achieved"], d["result"]["logl"], d["result"]["df"]))
except Exception as ex:
    print("ht $ht failed", ex, open("gpurun_out/s8/super_ht$ht.err").read()[-800:])
PY
done
