#!/bin/bash
# round 2, session 10: the whole GPU suite as the driver runs it, smoke, default bench (both arms), codon bench + ncu of k_walk_gen
mkdir -p gpurun_out/r2_10
O=gpurun_out/r2_10
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 300 $O/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 300 $O/bench_ref.err
timeout 600 python bench.py --workload codon_100x20k_rev61_g4 --steps 5 --warmup 3 > $O/codon.json 2> $O/codon.err; echo "codon rc=$?"; tail -c 600 $O/codon.err
timeout 600 python bench.py --workload codon_100x20k_rev61_g4 --steps 5 --warmup 3 --schedule 1 --no-cpu-baseline > $O/codon_levelled.json 2> $O/codon_levelled.err; echo "codon levelled rc=$?"
python - <<PY
import json
for f in ("bench_default","bench_ref","codon","codon_levelled"):
    try:
        d=json.load(open("$O/%s.json"%f))
        print(f, "value %.3f e2e %.3f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), json.dumps(d.get("roofline",{}))[:700])
        if "secondary" in d: s=d["secondary"]; print("  secondary value %.3f e2e %.3f ms %.3f" % (s["value"], s["e2e"]["value"], s["ms_per_step"]))
    except Exception as ex: print(f, "failed", ex)
PY
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_walk_gen -s 3 -c 1 -o $O/walkgen python bench.py --workload codon_100x20k_rev61_g4 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/ncu.log; tail -3 $O/ncu.log
