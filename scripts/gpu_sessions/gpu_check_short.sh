#!/bin/bash
# short end-of-session check (the session-8 re-entry): the whole GPU suite, smoke(), one short bench line for configs 2 and 3
mkdir -p gpurun_out/s8
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py --steps 5 --warmup 3 --workload aa_200x100k_lg_g4 --no-cpu-baseline > gpurun_out/s8/final_cfg3.json 2> gpurun_out/s8/final_cfg3.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s8/final_cfg2.json 2> gpurun_out/s8/final_cfg2.err
python - <<'PY'
import json
for c in ("cfg2", "cfg3"):
    try:
        d = json.load(open("gpurun_out/s8/final_%s.json" % c))
        print(c, "value %.2f e2e %.2f ms/launch %.3f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"]))
    except Exception as ex:
        print(c, "failed", ex, open("gpurun_out/s8/final_%s.err" % c).read()[-400:])
PY
