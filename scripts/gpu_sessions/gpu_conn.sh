#!/bin/bash
mkdir -p gpurun_out/s8
for c in 8 32; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c B200_TRACE=1 python bench.py --workload super_500part_100taxa --steps 5 --warmup 3 --host-threads 8 > gpurun_out/s8/conn$c.json 2> gpurun_out/s8/conn$c.err
  python -c "import json; d=json.load(open('gpurun_out/s8/conn$c.json')); print('connections $c: ms/step', d['ms_per_step'], 'evals/s', d['value'])"
  grep computeFunction gpurun_out/s8/conn$c.err | tail -2
done
