#!/bin/bash
# round 2, session 14: non-reversible kernels — parity against the restatement and the live reference
mkdir -p gpurun_out/r2_14
O=gpurun_out/r2_14
timeout 900 python -m pytest tests/test_gpu_nonrev.py -q -x 2>&1 | tail -40 > $O/pytest_nonrev.log; tail -40 $O/pytest_nonrev.log
