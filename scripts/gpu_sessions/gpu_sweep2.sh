#!/bin/bash
python -m pytest tests/test_gpu_phylotree.py tests/test_gpu_phylotree_multifurc_asc.py tests/test_gpu_supertree.py -m gpu -q -x 2>&1 | tail -8
for args in "--taxa 200 --patterns 100000 --kind dna" "--taxa 100 --patterns 20000 --kind aa" "--taxa 60 --patterns 4000 --kind dna" "--taxa 1000 --patterns 20000 --kind dna"; do
  for p in "" "--pipelined"; do
    python scripts/bench_sweep.py $args --no-reference $p 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$args $p', 'ms %.2f' % (d['gpu_sec']*1e3), 'logl_after %.4f' % d['logl_after'])"
  done
done
