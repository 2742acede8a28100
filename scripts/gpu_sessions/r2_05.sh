#!/bin/bash
mkdir -p gpurun_out/r2_05
O=gpurun_out/r2_05
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "multi_block" 2>&1 | tail -60 > $O/pytest_mb.log
grep -E "^FAILED|passed|failed|AssertionError|assert |differs|Mismatch|Max abs|slot" $O/pytest_mb.log | head -40
