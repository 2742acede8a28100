#!/usr/bin/env python3
"""Secondary measurement: the drop-in itself.  The UNMODIFIED reference's PhyloTree runs its own optimizeAllBranches (one sweep)
and getBestNNIForBran (every inner branch) twice: over its AVX/FMA kernel on the host threads, and with its kernel pointers bound to
libb200lk.so (oracle/ref_b200_shim.cpp).  Same tree, same data; the CPU arm on a pattern sample, scaled.
usage: bench_shim.py [--taxa 200] [--patterns 100000] [--kind dna|aa] [--cpu-patterns 10000]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from iqtree2_b200 import models, treesim
from oracle import ref_engine
from refmap import branch_splits

ap = argparse.ArgumentParser()
ap.add_argument("--taxa", type=int, default=200)
ap.add_argument("--patterns", type=int, default=100000)
ap.add_argument("--kind", default="dna")
ap.add_argument("--cpu-patterns", type=int, default=10000)
a = ap.parse_args()
model = models.make_gtr_gamma() if a.kind == "dna" else models.make_lg_gamma()
tree = treesim.yule_tree(a.taxa, seed=1)
pat = treesim.evolve_patterns_fast(tree, model.nstates, a.patterns, seed=2)
freq = np.ones(a.patterns)
inner = [s for s, (u, v) in zip(branch_splits(tree), tree.edges) if u >= a.taxa and v >= a.taxa]


def run(r):
    r.compute_likelihood()
    t0 = time.perf_counter(); ll = r.optimize_all_branches(1, 0.001, 100); t_sweep = time.perf_counter() - t0
    t0 = time.perf_counter(); best = max(max(m["score"] for m in r.best_nni(s)) for s in inner); t_nni = time.perf_counter() - t0
    return ll, best, t_sweep, t_nni


nt = min(os.cpu_count(), ref_engine.max_threads())   # before any tree is made: the harness sets the process-wide OpenMP thread count
g = ref_engine.RefTree(tree, model, pat, freq, nthreads=1, b200=True)
run(g)                                   # warm (first launches, allocations)
g.close()
g = ref_engine.RefTree(tree, model, pat, freq, nthreads=1, b200=True)
ll_g, best_g, sweep_g, nni_g = run(g)
calls, launches = g.b200_kernel_calls()
g.close()
f = ref_engine.RefTree(tree, model, pat, freq, nthreads=1, b200=True)
f.b200_set_fused_newton(True)            # the optional hook: optimizeOneBranch -> b200lk_newton_branch
ll_f, best_f, sweep_f, nni_f = run(f)
f.close()
n = min(a.cpu_patterns, a.patterns)
c = ref_engine.RefTree(tree, model, np.ascontiguousarray(pat[:, :n]), freq[:n], nthreads=nt)
ll_c, best_c, sweep_c, nni_c = run(c)
k = a.patterns / n
print(json.dumps({"what": "the reference's own optimizeAllBranches (1 sweep) and getBestNNIForBran (all inner branches): AVX/FMA kernel vs kernel pointers bound to libb200lk.so",
                  "taxa": a.taxa, "patterns": a.patterns, "kind": a.kind, "inner_branches": len(inner),
                  "b200_sweep_sec": sweep_g, "b200_nni_sec": nni_g, "b200_kernel_calls": calls, "b200_launches": launches,
                  "b200_fused_newton_sweep_sec": sweep_f, "b200_fused_newton_nni_sec": nni_f, "logl_b200_fused_newton": ll_f,
                  "cpu_threads": nt, "cpu_sample_patterns": n, "cpu_sweep_sec_scaled": sweep_c * k, "cpu_nni_sec_scaled": nni_c * k,
                  "speedup_sweep": sweep_c * k / sweep_g, "speedup_nni": nni_c * k / nni_g, "logl_b200": ll_g, "logl_cpu_sample": ll_c}))
