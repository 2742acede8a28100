#!/bin/bash
# round 2, session 15: the reference's rooted PhyloTree + non-reversible model over the CUDA kernels; non-reversible bench line; whole suite
mkdir -p gpurun_out/r2_15
O=gpurun_out/r2_15
timeout 600 python -m pytest tests/test_gpu_reference_shim.py -q -x -k nonrev 2>&1 | tail -15
timeout 600 python bench.py --workload dna_500x200k_unrest_g4 --steps 5 --warmup 3 > $O/nonrev.json 2> $O/nonrev.err; echo "nonrev rc=$?"; tail -c 600 $O/nonrev.err
python - <<PY
import json
d=json.load(open("$O/nonrev.json"))
print("nonrev value %.3f e2e %.3f ms %.3f logl %.6f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["result"]["logl"]), json.dumps(d.get("roofline",{}))[:900])
print(d.get("cpu_baseline"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 16 --csv --log-file $O/launches_nonrev.csv python bench.py --workload dna_500x200k_unrest_g4 --steps 2 --warmup 3 --no-cpu-baseline > $O/launch_run.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_nonrev.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print("%-70s %s %s" % (r[ki][:70], r[vi], r[ui]))
PY
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
