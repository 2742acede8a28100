#!/bin/bash
# compute-sanitizer over the small parity cases: memcheck (out-of-bounds / misaligned), racecheck (shared-memory hazards in the
# walkers' staging rings, category-split maxima, block reductions), synccheck
O=gpurun_out/s8san; mkdir -p $O
SEL="test_parity_small or test_parity_kernel_variants or test_edge_cases or test_multifurcating or test_asc_lewis or test_rell or test_supertree_logl"
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 77 --log-file $O/$tool.log python -m pytest tests -m gpu -q -x -k "$SEL" > $O/$tool.pytest 2>&1
  echo "$tool rc=$? $(tail -1 $O/$tool.pytest)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/$tool.log | tail -3
done
