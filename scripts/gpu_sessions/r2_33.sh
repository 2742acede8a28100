#!/bin/bash
# round 2, session 33: the final binary — whole GPU suite, smoke, default bench line
mkdir -p gpurun_out/r2_33
O=gpurun_out/r2_33
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $O/pytest_gpu_tail.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 300 $O/bench_default.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench_default.json") if l.startswith("{")][-1])
r=d["roofline"]
print("default value %.3f e2e %.3f ms %.3f launch_ms %.3f frac %.3f launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["launch_ms"], r["frac"], d.get("gpu_launches")))
s=d["secondary"]; print("secondary value %.3f e2e %.3f" % (s["value"], s["e2e"]["value"]))
print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"]["kind"])
PY
