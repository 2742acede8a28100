#!/bin/bash
# NNI evaluation (getBestNNIForBran over every inner branch) against the compiled reference on the box's host threads
mkdir -p gpurun_out/s8
timeout 250 python scripts/bench_nni.py --taxa 200 --patterns 100000 --kind dna --cpu-patterns 10000 > gpurun_out/s8/nni_dna.json 2> gpurun_out/s8/nni_dna.err; tail -c 1500 gpurun_out/s8/nni_dna.json
timeout 250 python scripts/bench_nni.py --taxa 100 --patterns 20000 --kind aa --cpu-patterns 4000 > gpurun_out/s8/nni_aa.json 2> gpurun_out/s8/nni_aa.err; tail -c 1500 gpurun_out/s8/nni_aa.json
timeout 120 python scripts/bench_nni.py --taxa 60 --patterns 4000 --kind dna > gpurun_out/s8/nni_small.json 2> gpurun_out/s8/nni_small.err; tail -c 1500 gpurun_out/s8/nni_small.json
for f in gpurun_out/s8/nni_*.err; do tail -n 5 $f; done
