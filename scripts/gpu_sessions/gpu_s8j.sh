#!/bin/bash
# A/B of the 4-state walker's per-child-kind specialisation (variant 0) against the single category-loop body (variant 3)
VARIANTS="0 3 0 3" PATTERNS=524288 bash scripts/gpu_sessions/gpu_n4diag.sh
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
