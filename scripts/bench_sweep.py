#!/usr/bin/env python3
"""Secondary measurement (not the BASELINE metric): one optimizeAllBranches sweep — the inner loop of tree search and of
ModelFinder (tree/phylotree.cpp:2698-2774: per branch a partial traversal of the stale CLVs, the theta buffer, a
Newton-Raphson iteration on it, computeLikelihoodFromBuffer) — through the host mirror over the CUDA engine (CLVs resident,
fused Newton) and through the compiled reference on the host cores, same tree, same data, same tolerances.
usage: bench_sweep.py [--taxa 200] [--patterns 100000] [--kind dna|aa] [--iterations 1]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from iqtree2_b200 import models, treesim
from iqtree2_b200.phylotree import PhyloTree

ap = argparse.ArgumentParser()
ap.add_argument("--taxa", type=int, default=200)
ap.add_argument("--patterns", type=int, default=100000)
ap.add_argument("--kind", default="dna")
ap.add_argument("--iterations", type=int, default=1)
ap.add_argument("--cpu-patterns", type=int, default=0, help="pattern sample for the reference arm (0 = same as --patterns)")
ap.add_argument("--no-reference", action="store_true")
ap.add_argument("--pipelined", action="store_true", help="queue every sweep as a whole on the device")
a = ap.parse_args()
model = models.make_gtr_gamma() if a.kind == "dna" else models.make_lg_gamma()
tree = treesim.yule_tree(a.taxa, seed=1)
pat = treesim.evolve_patterns_fast(tree, model.nstates, a.patterns, seed=2)
freq = np.ones(a.patterns)
nwk = tree.to_newick(fmt="%.17g")
g = PhyloTree()
g.set_alignment(tree.names, pat, freq, model); g.set_model(model); g.read_tree(nwk)
if a.pipelined:
    g.set_pipelined_sweep(True)
ll0 = g.compute_likelihood()
g.read_tree(nwk); g.clear_all_partial_lh(); g.compute_likelihood()      # warm (allocations, first launches)
import torch
torch.cuda.synchronize()
t0 = time.perf_counter()
ll1 = g.optimize_all_branches(iterations=a.iterations, tolerance=0.001)
torch.cuda.synchronize()
tg = time.perf_counter() - t0
out = {"what": "one optimizeAllBranches call", "taxa": a.taxa, "patterns": a.patterns, "kind": a.kind, "iterations": a.iterations, "pipelined": a.pipelined,
       "gpu_sec": tg, "logl_before": ll0, "logl_after": ll1, "partial_lh_updates": g.num_partial_lh_computations,
       "branches": int(len(g.branches()[2]))}
if not a.no_reference:
    from oracle import ref_engine
    n = a.cpu_patterns or a.patterns
    r = ref_engine.RefTree(tree, model, np.ascontiguousarray(pat[:, :n]), freq[:n], nthreads=min(os.cpu_count(), ref_engine.max_threads()))
    r.compute_likelihood()
    t0 = time.perf_counter()
    llr = r.optimize_all_branches(iterations=a.iterations, tolerance=0.001)
    tr = time.perf_counter() - t0
    out.update({"reference_sec_on_sample": tr, "reference_sample_patterns": n, "reference_threads": r.info()["num_threads"],
                "reference_sec_scaled": tr * a.patterns / n, "speedup": tr * a.patterns / n / tg, "reference_logl_after_on_sample": llr})
    if n == a.patterns:
        _, _, ln = g.branches()
        out["max_abs_sorted_branch_length_difference"] = float(np.abs(np.sort(ln) - np.sort(np.array(list(r.branches().values())))).max())
print(json.dumps(out))
