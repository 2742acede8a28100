#!/usr/bin/env python3
"""The BASELINE metric (full-tree logL + derivative evaluations/s: clearAllPartialLH -> computeLikelihoodBranch ->
computeLikelihoodDerv) measured through the reference's own PhyloTree object: over its AVX/FMA kernel on the host threads, and with
its kernel pointers bound to libb200lk.so (oracle/ref_b200_shim.cpp).  The bound arm includes everything a host pays: the reference's
traversal planning, marshalling, the C ABI calls, _pattern_lh copied back to the host after every evaluation.
usage: bench_shim_eval.py [--taxa 200] [--patterns 100000] [--kind aa|dna] [--cpu-patterns 5000] [--reps 10]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from iqtree2_b200 import models, treesim
from oracle import ref_engine

ap = argparse.ArgumentParser()
ap.add_argument("--taxa", type=int, default=200)
ap.add_argument("--patterns", type=int, default=100000)
ap.add_argument("--kind", default="aa")
ap.add_argument("--cpu-patterns", type=int, default=5000)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
model = models.make_gtr_gamma() if a.kind == "dna" else models.make_lg_gamma()
tree = treesim.yule_tree(a.taxa, seed=1)
pat = treesim.evolve_patterns_fast(tree, model.nstates, a.patterns, seed=2)
freq = np.ones(a.patterns)
nt = min(os.cpu_count(), ref_engine.max_threads())
g = ref_engine.RefTree(tree, model, pat, freq, nthreads=1, b200=True)
g.time_eval(0, 3)
rg = g.time_eval(0, a.reps)
g.close()
n = min(a.cpu_patterns, a.patterns)
c = ref_engine.RefTree(tree, model, np.ascontiguousarray(pat[:, :n]), freq[:n], nthreads=nt)
c.time_eval(0, 1)
rc = c.time_eval(0, 3)
k = a.patterns / n
print(json.dumps({"what": "BASELINE metric through the reference's own PhyloTree object", "taxa": a.taxa, "patterns": a.patterns, "kind": a.kind,
                  "bound_to_libb200lk_evals_per_s": 1.0 / rg["mean_sec"], "bound_ms_per_eval": 1e3 * rg["mean_sec"], "bound_best_ms": 1e3 * rg["best_sec"],
                  "reference_avx_fma_evals_per_s": 1.0 / (rc["mean_sec"] * k), "reference_threads": nt, "reference_sample_patterns": n,
                  "ratio": rc["mean_sec"] * k / rg["mean_sec"], "logl_bound": rg["logl"], "df_bound": rg["df"]}))
