#!/bin/bash
# session 8, call A: restored state check + traffic capture
mkdir -p gpurun_out/s8
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/s8/gpu.txt
nproc >> gpurun_out/s8/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s8/pytest_gpu.log 2>&1
tail -3 gpurun_out/s8/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/s8/bench_cfg2.json 2> gpurun_out/s8/bench_cfg2.err
timeout 600 python bench.py --workload aa_200x100k_lg_g4 > gpurun_out/s8/bench_cfg3.json 2> gpurun_out/s8/bench_cfg3.err
# DRAM traffic of the walker at full size: single-pass metrics, one launch after warm-up
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_walk -s 2 -c 1 --csv --log-file gpurun_out/s8/traffic_cfg2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s8/traffic_cfg2.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_walk -s 2 -c 1 --csv --log-file gpurun_out/s8/traffic_cfg3.csv python bench.py --workload aa_200x100k_lg_g4 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/s8/traffic_cfg3.log 2>&1
cat gpurun_out/s8/bench_cfg2.json | cut -c1-400
cat gpurun_out/s8/bench_cfg3.json | cut -c1-400
tail -3 gpurun_out/s8/traffic_cfg2.csv gpurun_out/s8/traffic_cfg3.csv
