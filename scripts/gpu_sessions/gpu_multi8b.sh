#!/bin/bash
N=${N:-8}
mkdir -p gpurun_out/s8
for wl in aa_200x100k_lg_g4 super_500part_100taxa; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 --workload $wl > gpurun_out/s8/bench2_${wl}_n$N.json 2> gpurun_out/s8/bench2_${wl}_n$N.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s8/bench2_${wl}_n$N.json") if l.startswith("{")][-1])
    print("$wl N=$N: evals/s %.2f ms/step %.3f e2e %.2f logl %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["result"]["logl"]))
except Exception as ex:
    print("$wl failed", ex, open("gpurun_out/s8/bench2_${wl}_n$N.err").read()[-600:])
PY
done
