#!/bin/bash
# round 2, session 19: the reference's PhyloTree with a site-specific model / mixed branch lengths over the CUDA kernels
mkdir -p gpurun_out/r2_19
timeout 900 python -m pytest tests/test_gpu_reference_shim.py -q -x 2>&1 | tail -25
