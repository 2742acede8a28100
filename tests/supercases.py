"""Shared partition-model test case: a Yule tree, a mix of DNA (GTR+G4) and protein (LG+G4) partitions with their own
rates, and taxa missing per partition (rows of STATE_UNKNOWN)."""
from __future__ import annotations

import numpy as np

from iqtree2_b200 import models, treesim
from cases import ptn_invar_for


def make_supercase(ntaxa=14, nparts=6, seed=4, missing=0.15, sites=(120, 400), pinvar=0.0):
    rng = np.random.default_rng(seed)
    tree = treesim.yule_tree(ntaxa, seed=seed, mean_len=0.12)
    parts = []
    for p in range(nparts):
        aa = (p % 3 == 2)
        if aa:
            model = models.make_lg_gamma(alpha=0.6 + 0.2 * rng.random(), pinvar=pinvar)
        else:
            r = 0.5 + 3.0 * rng.random(6)
            pi = rng.dirichlet([8, 8, 8, 8])
            model = models.make_gtr_gamma(rates6=tuple(r), pi=tuple(pi), alpha=0.4 + rng.random(), pinvar=pinvar)
        rate = float(rng.gamma(2.0, 0.5)) + 0.05
        nsites = int(rng.integers(sites[0], sites[1]))
        scaled = treesim.Tree(tree.ntaxa, tree.edges.copy(), tree.lengths * rate, tree.names)
        aln = treesim.simulate_alignment(scaled, model, nsites, seed=seed * 100 + p)
        absent = rng.random(ntaxa) < missing
        if absent.sum() > ntaxa - 4:
            absent[:] = False
        aln = aln.copy()
        aln[absent, :] = model.state_unknown
        pat, freq = treesim.compress_patterns(aln)
        inv = ptn_invar_for(model, pat)
        parts.append(dict(model=model, rate=rate, pat=pat, freq=freq, inv=inv, present=~absent))
    return tree, parts


def induced_subtree(tree, keep):
    """The subtree of `tree` induced by the leaves keep[i] == True: degree-2 nodes suppressed, their branch lengths summed
    (what PhyloSuperTree::linkTrees / mapTrees give a partition tree).  Returns (treesim.Tree over the kept leaves in their
    original order, list of kept leaf ids)."""
    adj = tree.adjacency()
    kept = [i for i in range(tree.ntaxa) if keep[i]]
    alive = [True] * tree.nnodes
    deg = [len(a) for a in adj]
    nbrs = [dict((v, float(tree.lengths[b])) for v, b in a) for a in adj]
    # remove dropped leaves, then repeatedly remove degree-1 internal nodes and splice degree-2 nodes
    stack = [i for i in range(tree.ntaxa) if not keep[i]]
    while stack:
        u = stack.pop()
        if not alive[u]:
            continue
        alive[u] = False
        for v in list(nbrs[u]):
            del nbrs[v][u]
            if v >= tree.ntaxa and len(nbrs[v]) == 1:
                stack.append(v)
        nbrs[u] = {}
    for u in range(tree.ntaxa, tree.nnodes):
        if alive[u] and len(nbrs[u]) == 2:
            (a, la), (b, lb) = nbrs[u].items()
            del nbrs[a][u]; del nbrs[b][u]
            nbrs[a][b] = nbrs[b][a] = la + lb
            nbrs[u] = {}
            alive[u] = False
    new_id = {}
    for k, i in enumerate(kept):
        new_id[i] = k
    nxt = len(kept)
    for u in range(tree.ntaxa, tree.nnodes):
        if alive[u]:
            new_id[u] = nxt; nxt += 1
    edges, lens = [], []
    for u in new_id:
        for v, l in nbrs[u].items():
            if u < v:
                edges.append([new_id[u], new_id[v]]); lens.append(l)
    assert nxt == 2 * len(kept) - 2 and len(edges) == 2 * len(kept) - 3
    return treesim.Tree(len(kept), np.array(edges), np.array(lens), [tree.names[i] for i in kept]), kept
