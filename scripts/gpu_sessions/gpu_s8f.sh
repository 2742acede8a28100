#!/bin/bash
mkdir -p gpurun_out/s8
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --patterns 262144"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n4c4 -s 2 -c 1 -o gpurun_out/s8/n4pf_full $B > gpurun_out/s8/n4pf_full.log 2>&1
ls -la gpurun_out/s8/n4pf_full.ncu-rep
