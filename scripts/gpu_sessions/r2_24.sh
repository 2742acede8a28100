#!/bin/bash
# round 2, session 24: fresh ncu --set full capture of k_walk_n4c4 at the full config-2 size (source-level stall attribution)
mkdir -p gpurun_out/r2_24
O=gpurun_out/r2_24
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_walk_n4c4 -s 3 -c 1 -f -o $O/walk_n4c4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1; echo "ncu rc=$?"
ls -la $O
