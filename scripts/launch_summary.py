#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total and share of time."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum": continue
    name = r[ix["Kernel Name"]].split("(")[0]
    v = float(r[ix["Metric Value"]].replace(",", "")); u = r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:8d} {t:10.3f} {100*t/tot:6.2f}%")
print(f"{'all':60s} {sum(a[0] for a in agg.values()):8d} {tot:10.3f}")
