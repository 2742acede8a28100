#!/bin/bash
mkdir -p gpurun_out/s8
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 
VARIANTS="0 1 2 3 4 5" bash scripts/gpu_sessions/gpu_variants20.sh
