#!/bin/bash
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "kernel_variants and 6" 2>&1 | tail -4
VARIANTS='0 6 0 6' bash scripts/gpu_sessions/gpu_variants20.sh
