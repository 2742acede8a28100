#!/bin/bash
# round 2, session 16: site-specific models (ModelSet / PMSF) on the device
mkdir -p gpurun_out/r2_16
timeout 900 python -m pytest tests/test_gpu_site.py -q -x 2>&1 | tail -25
