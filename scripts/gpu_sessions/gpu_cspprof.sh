#!/bin/bash
mkdir -p gpurun_out/s8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20_mma -s 2 -c 1 -o gpurun_out/s8/n20csp_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload aa_200x100k_lg_g4 --patterns 4096 > gpurun_out/s8/n20csp_full.log 2>&1
ls -la gpurun_out/s8/n20csp_full.ncu-rep
