#!/bin/bash
O=gpurun_out/s8san2; mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file $O/memcheck.log python -m pytest tests -m gpu -q -x -k "test_asc_lewis or test_parity_small or test_multifurcating or test_compress_matches" > $O/memcheck.pytest 2>&1
echo "memcheck rc=$? $(tail -1 $O/memcheck.pytest)"; grep -E "ERROR SUMMARY" $O/memcheck.log | tail -1; grep -E "Invalid|at .*b200lk" $O/memcheck.log | head -6
