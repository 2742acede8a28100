#!/bin/bash
# round 2, session 18: mixed branch lengths (heterotachy) on the device
mkdir -p gpurun_out/r2_18
timeout 900 python -m pytest tests/test_gpu_mixlen.py tests/test_gpu_site.py -q -x 2>&1 | tail -25
