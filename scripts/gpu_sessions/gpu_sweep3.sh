#!/bin/bash
for i in 1 2 3; do for p in "" "--pipelined"; do
    python scripts/bench_sweep.py --taxa 100 --patterns 20000 --kind aa --no-reference $p 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('aa 100x20000 $p', 'ms %.2f' % (d['gpu_sec']*1e3))"
done; done
