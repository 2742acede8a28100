#!/bin/bash
mkdir -p gpurun_out/s8
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_walk_n4c4 -s 2 -c 1 --csv --log-file gpurun_out/s8/traffic_cfg2.csv $B > gpurun_out/s8/traffic_cfg2.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_walk_n20_mma -s 2 -c 1 --csv --log-file gpurun_out/s8/traffic_cfg3.csv $B --workload aa_200x100k_lg_g4 > gpurun_out/s8/traffic_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n20_mma -s 2 -c 1 -o gpurun_out/s8/n20_full $B --workload aa_200x100k_lg_g4 > gpurun_out/s8/n20_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk_n4c4 -s 2 -c 1 -o gpurun_out/s8/n4_full $B --patterns 262144 > gpurun_out/s8/n4_full.log 2>&1
grep -h k_walk gpurun_out/s8/traffic_cfg2.csv gpurun_out/s8/traffic_cfg3.csv | cut -d, -f5,13-15
ls -la gpurun_out/s8
