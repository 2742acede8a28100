#!/usr/bin/env python3
"""Secondary measurement: the E-step of ModelMixture::optimizeWithEM / RateFree::optimizeWithEM on the device
(b200lk_em_posterior) against the reference's data flow through the C ABI (read _pattern_lh_cat back, loop on the host).
usage: bench_em.py [--taxa 100] [--patterns 20000] [--model c60|lg|gtr]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from iqtree2_b200 import models, plan, treesim
from iqtree2_b200.engine import LikelihoodEngine

ap = argparse.ArgumentParser()
ap.add_argument("--taxa", type=int, default=100)
ap.add_argument("--patterns", type=int, default=20000)
ap.add_argument("--model", default="c60")
a = ap.parse_args()
model = {"c60": models.make_lg_c60_gamma, "lg": models.make_lg_gamma, "gtr": models.make_gtr_gamma}[a.model]()
wsl, add_invar = (3, False) if model.nmix > 1 else (2, True)
tree = treesim.yule_tree(a.taxa, seed=1)
pat = treesim.evolve_patterns_fast(tree, model.nstates, a.patterns, seed=2)
freq = np.ones(a.patterns)
g = LikelihoodEngine(model, tree.ntaxa, a.patterns)
g.set_tip_states(pat); g.set_pattern_freq(freq); g.set_pattern_invar(np.zeros(a.patterns))
ops, ea, eb, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
ll = g.branch_logl(ops, ea, eb, L)
g.em_posterior(wsl, add_invar)
reps = 10
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(reps):
    prop = g.em_posterior(wsl, add_invar)
torch.cuda.synchronize(); t_dev = (time.perf_counter() - t0) / reps
ncat, nmix = model.ncat, model.nmix
t0 = time.perf_counter()
for _ in range(3):
    cat = g.get_pattern_lh_cat().reshape(a.patterns, -1)
    if model.nmix > 1:
        cat = cat.reshape(a.patterns, nmix, ncat).sum(axis=2)
    lk = cat.sum(axis=1)
    prop_h = (cat * (freq / lk)[:, None]).sum(axis=0)
t_host = (time.perf_counter() - t0) / 3
print(json.dumps({"what": "EM E-step", "model": a.model, "taxa": a.taxa, "patterns": a.patterns, "categories": int(g.Cmix), "K": int(len(prop)),
                  "device_ms": 1e3 * t_dev, "readback_and_host_loop_ms": 1e3 * t_host, "pattern_lh_cat_MB": a.patterns * int(g.Cmix) * 8 / 1e6,
                  "max_rel_diff": float(np.abs(prop / prop_h - 1).max()), "logl": ll}))
