#!/bin/bash
# round 2, session 22: whole GPU suite after the site-specific / mixlen work; smoke; default bench line
mkdir -p gpurun_out/r2_22
O=gpurun_out/r2_22
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $O/pytest_gpu_tail.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 400 $O/bench_default.err
python - <<PY
import json
d=json.load(open("$O/bench_default.json"))
print("default value %.3f e2e %.3f ms %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"]))
s=d.get("secondary")
if s: print("secondary value %.3f e2e %.3f launch_ms %.3f" % (s["value"], s["e2e"]["value"], s["roofline"]["launch_ms"]))
PY
