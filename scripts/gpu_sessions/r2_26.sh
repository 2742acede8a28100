#!/bin/bash
# round 2, session 26: ncu --set full of k_walk_n4c4 variant 25 (tip states through shared memory) at the full config-2 size
mkdir -p gpurun_out/r2_26
O=gpurun_out/r2_26
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_walk_n4c4 -s 3 -c 1 -f -o $O/walk_n4c4_v25 python bench.py --steps 1 --warmup 3 --variant 25 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1; echo "ncu rc=$?"
