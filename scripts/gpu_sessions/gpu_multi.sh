#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out/s8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/s8/bench_cfg2_n$N.json 2> gpurun_out/s8/bench_cfg2_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --workload super_500part_100taxa > gpurun_out/s8/bench_cfg5_n$N.json 2> gpurun_out/s8/bench_cfg5_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --workload aa_200x100k_lg_g4 > gpurun_out/s8/bench_cfg3_n$N.json 2> gpurun_out/s8/bench_cfg3_n$N.err
for f in cfg2 cfg5 cfg3; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/s8/bench_${f}_n$N.json") if l.startswith("{")][-1])
    print("$f N=$N: evals/s %.2f ms/step %.3f e2e %.2f logl %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["result"]["logl"]))
except Exception as ex:
    print("$f failed", ex, open("gpurun_out/s8/bench_${f}_n$N.err").read()[-600:])
PY
done
