#!/usr/bin/env python3
"""Secondary measurement (not the BASELINE metric): NNI evaluation — PhyloTree::getBestNNIForBran (tree/phylotree.cpp:4069-4322;
per inner branch two alternative topologies, each with five Newton-Raphson branch optimisations on the CLVs around the branch)
over every inner branch of the tree: the host mirror over the CUDA engine (CLVs resident, the two spare slots, fused Newton)
against the compiled reference on the host cores, same tree, same data.  This is the loop that dominates IQ-TREE's tree search
(IQTree::evaluateNNIs, tree/iqtree.cpp:2996-3167).
usage: bench_nni.py [--taxa 200] [--patterns 100000] [--kind dna|aa] [--cpu-patterns N] [--branches K]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from iqtree2_b200 import models, treesim
from iqtree2_b200.phylotree import PhyloTree

ap = argparse.ArgumentParser()
ap.add_argument("--taxa", type=int, default=200)
ap.add_argument("--patterns", type=int, default=100000)
ap.add_argument("--kind", default="dna")
ap.add_argument("--branches", type=int, default=0, help="evaluate only the first K inner branches (0 = all)")
ap.add_argument("--cpu-patterns", type=int, default=0, help="pattern sample for the reference arm (0 = same as --patterns)")
ap.add_argument("--no-reference", action="store_true")
a = ap.parse_args()
model = models.make_gtr_gamma() if a.kind == "dna" else models.make_lg_gamma()
tree = treesim.yule_tree(a.taxa, seed=1)
pat = treesim.evolve_patterns_fast(tree, model.nstates, a.patterns, seed=2)
freq = np.ones(a.patterns)
nwk = tree.to_newick(fmt="%.17g")
g = PhyloTree()
g.set_alignment(tree.names, pat, freq, model); g.set_model(model); g.read_tree(nwk)
ll0 = g.compute_likelihood()
na, nb, _ = g.branches()
inner = [b for b in range(len(na)) if na[b] >= a.taxa and nb[b] >= a.taxa]
if a.branches:
    inner = inner[:a.branches]
g.get_best_nni_for_bran(inner[0])      # warm (first launches)
import torch
torch.cuda.synchronize()
n0 = g.num_partial_lh_computations
t0 = time.perf_counter()
scores = [max(m["score"] for m in g.get_best_nni_for_bran(b)) for b in inner]
torch.cuda.synchronize()
tg = time.perf_counter() - t0
out = {"what": "getBestNNIForBran (nni5) over inner branches", "taxa": a.taxa, "patterns": a.patterns, "kind": a.kind,
       "inner_branches": len(inner), "gpu_sec": tg, "gpu_ms_per_branch": 1e3 * tg / len(inner), "logl": ll0,
       "best_nni_score": max(scores), "partial_lh_updates": int(g.num_partial_lh_computations - n0),
       "logl_after": g.compute_likelihood()}
if not a.no_reference:
    from oracle import ref_engine
    from refmap import branch_splits
    n = a.cpu_patterns or a.patterns
    r = ref_engine.RefTree(tree, model, np.ascontiguousarray(pat[:, :n]), freq[:n], nthreads=min(os.cpu_count(), ref_engine.max_threads()))
    r.compute_likelihood()
    spl = [s for s, (u, v) in zip(branch_splits(tree), tree.edges) if u >= a.taxa and v >= a.taxa]
    if a.branches:
        spl = spl[:a.branches]
    r.best_nni(spl[0])
    t0 = time.perf_counter()
    rs = [max(m["score"] for m in r.best_nni(s)) for s in spl]
    tr = time.perf_counter() - t0
    out.update({"reference_sec_on_sample": tr, "reference_sample_patterns": n, "reference_threads": r.info()["num_threads"],
                "reference_inner_branches": len(spl), "reference_sec_scaled": tr * a.patterns / n * len(inner) / len(spl),
                "speedup": tr * a.patterns / n * len(inner) / len(spl) / tg})
    if n == a.patterns and len(spl) == len(inner):
        out["max_abs_difference_of_best_scores"] = float(abs(max(rs) - max(scores)))
        out["max_abs_difference_of_sorted_scores"] = float(np.abs(np.sort(rs) - np.sort(scores)).max())
print(json.dumps(out))
