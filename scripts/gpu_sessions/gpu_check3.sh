#!/bin/bash
O=gpurun_out/s8san2; mkdir -p $O
timeout 420 compute-sanitizer --tool racecheck --error-exitcode 77 --log-file $O/racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_parity_small and (aa or dna) and not dna_" > $O/racecheck.pytest 2>&1
echo "racecheck rc=$? $(tail -1 $O/racecheck.pytest)"; grep -E "RACECHECK SUMMARY" $O/racecheck.log | tail -1
grep -E "^========= (Error|Warning)" $O/racecheck.log | cut -c1-200 | sort | uniq -c | sort -rn | head -8
