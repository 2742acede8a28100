#!/bin/bash
mkdir -p gpurun_out/s8
python scripts/bench_sweep.py --taxa 200 --patterns 100000 --kind dna --cpu-patterns 20000 > gpurun_out/s8/sweep_dna_200x100k.json 2> gpurun_out/s8/sweep1.err; cat gpurun_out/s8/sweep_dna_200x100k.json; tail -2 gpurun_out/s8/sweep1.err
python scripts/bench_sweep.py --taxa 100 --patterns 20000 --kind aa --cpu-patterns 5000 > gpurun_out/s8/sweep_aa_100x20k.json 2> gpurun_out/s8/sweep2.err; cat gpurun_out/s8/sweep_aa_100x20k.json; tail -2 gpurun_out/s8/sweep2.err
python scripts/bench_sweep.py --taxa 60 --patterns 4000 --kind dna > gpurun_out/s8/sweep_dna_60x4k.json 2> gpurun_out/s8/sweep3.err; cat gpurun_out/s8/sweep_dna_60x4k.json; tail -2 gpurun_out/s8/sweep3.err
