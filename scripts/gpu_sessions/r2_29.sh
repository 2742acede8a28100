#!/bin/bash
# round 2, session 29: final state — whole GPU suite, smoke, default bench line + reference arm, launch list and ncu --set full of the
# dominant kernel, the secondary workloads' lines
mkdir -p gpurun_out/r2_29
O=gpurun_out/r2_29
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $O/pytest_gpu_tail.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -c 300 $O/bench_default.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 300 $O/bench_ref.err
for w in aa_100x50k_lg_c60_g4 codon_100x20k_rev61_g4 dna_500x200k_unrest_g4 aa_100x20k_pmsf_g4 super_500part_100taxa; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $O/$w.json 2> $O/$w.err; echo "$w rc=$?"; tail -c 200 $O/$w.err
done
python - <<PY
import json
def last(f):
    return json.loads([l for l in open(f) if l.startswith("{")][-1])
for f in ("bench_default","bench_ref","aa_100x50k_lg_c60_g4","codon_100x20k_rev61_g4","dna_500x200k_unrest_g4","aa_100x20k_pmsf_g4","super_500part_100taxa"):
    try:
        d=last("$O/%s.json"%f)
        r=d.get("roofline") or {}
        print(f, "value %.3f e2e %.3f ms %.3f launch_ms %s frac %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r.get("launch_ms"), r.get("frac")))
        s=d.get("secondary")
        if s: print("   secondary value %.3f e2e %.3f launch_ms %s" % (s["value"], s["e2e"]["value"], (s.get("roofline") or {}).get("launch_ms")))
    except Exception as ex:
        print(f, "failed", ex)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $O/launch_run.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_walk_n4c4 -s 3 -c 1 -f -o $O/walk_n4c4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu.log 2>&1; echo "ncu rc=$?"
