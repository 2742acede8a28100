#!/bin/bash
mkdir -p gpurun_out/s8
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --workload aa_200x100k_lg_g4"
for v in 0 8 9; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20_mma -s 2 -c 1 -o gpurun_out/s8/n20v${v}_full $B --variant $v > gpurun_out/s8/n20v${v}_full.log 2>&1
done
ls -la gpurun_out/s8/*.ncu-rep
