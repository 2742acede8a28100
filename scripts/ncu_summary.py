#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: ncu_summary.py file.ncu-rep [row]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2 + row]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active"]
for k in keys:
    if k in d: print(f"{k:80s} {d[k][0]} {d[k][1]}")
st = sorted(((float(v[0]), h) for h, v in d.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and v[0]), reverse=True)
print("stall cycles per issued instruction (top):")
for v, h in st[:9]: print(f"   {v:6.2f}  {h.split('stalled_')[1].split('_per_issue')[0]}")
