#!/bin/bash
# round 2, session 34: ncu summary of k_nonrev_walk4 (rooted DNA 500 x 200k, 12-parameter model + G4) and of k_site_walk20 (final form);
# the reports are read on the box (raw page -> CSV), only the CSVs travel back
mkdir -p gpurun_out/r2_34
O=gpurun_out/r2_34
T=/tmp/ncu34; mkdir -p $T
timeout 900 ncu --set full --clock-control none -k regex:k_nonrev_walk4 -s 3 -c 1 -f -o $T/nonrev_walk4 python bench.py --workload dna_500x200k_unrest_g4 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_nonrev.log 2>&1; echo "ncu nonrev rc=$?"
ncu -i $T/nonrev_walk4.ncu-rep --page raw --csv > $O/nonrev_walk4_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:k_site_walk20 -s 3 -c 1 -f -o $T/site_walk20 python bench.py --workload aa_100x20k_pmsf_g4 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_site.log 2>&1; echo "ncu site rc=$?"
ncu -i $T/site_walk20.ncu-rep --page raw --csv > $O/site_walk20_raw.csv 2>/dev/null
timeout 600 python bench.py --workload dna_500x200k_unrest_g4 --steps 8 --warmup 3 > $O/nonrev.json 2> $O/nonrev.err; echo "nonrev rc=$?"
timeout 600 python bench.py --workload aa_100x20k_pmsf_g4 --steps 8 --warmup 3 > $O/site.json 2> $O/site.err; echo "site rc=$?"
ls -la $O
