"""Synthetic workload generators (host side, numpy): Yule-Harding trees and AliSim-style alignments.

SURVEY §8(d) recipe: Yule-Harding topology (the family the reference uses in testNumThreads,
tree/iqtree.cpp:4594-4598), branch lengths Exp(mean 0.1) clipped to [1e-6, 10]
(utils/tools.cpp:1235-1238), sequences evolved down the tree under the named model with
per-site Gamma category and (for mixtures) per-site class.  Seeds: tree 1, sequences 2, parameters 3.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Tree:
    """Unrooted binary tree in array form. Nodes 0..ntaxa-1 are leaves (id == alignment row,
    as PhyloNode::id for leaves), ntaxa..2*ntaxa-3 are internal nodes (degree 3)."""
    ntaxa: int
    edges: np.ndarray     # [nbranch, 2] node ids
    lengths: np.ndarray   # [nbranch]
    names: list

    @property
    def nnodes(self):
        return 2 * self.ntaxa - 2

    def adjacency(self):
        adj = [[] for _ in range(self.nnodes)]
        for b, (u, v) in enumerate(self.edges):
            adj[int(u)].append((int(v), b))
            adj[int(v)].append((int(u), b))
        return adj

    def to_newick(self, root_leaf: int = 0, fmt="%.10g") -> str:
        adj = self.adjacency()
        inner = adj[root_leaf][0][0]

        # simple recursive version (trees here are <= a few thousand taxa; raise recursion limit)
        import sys
        sys.setrecursionlimit(max(10000, 4 * self.nnodes))

        def rec2(node, dad):
            if node < self.ntaxa:
                return self.names[node]
            parts = []
            for v, b in adj[node]:
                if v == dad:
                    continue
                parts.append(rec2(v, node) + ":" + (fmt % self.lengths[b]))
            return "(" + ",".join(parts) + ")"

        kids = []
        b0 = adj[root_leaf][0][1]
        kids.append(self.names[root_leaf] + ":" + (fmt % self.lengths[b0]))
        for v, b in adj[inner]:
            if v == root_leaf:
                continue
            kids.append(rec2(v, inner) + ":" + (fmt % self.lengths[b]))
        return "(" + ",".join(kids) + ");"


def yule_tree(ntaxa: int, seed: int = 1, mean_len: float = 0.1, min_len: float = 1e-6, max_len: float = 10.0) -> Tree:
    """Yule-Harding: grow by splitting a uniformly chosen pendant edge; unrooted result."""
    assert ntaxa >= 3
    rng = np.random.default_rng(seed)
    # start with a 3-taxon star around internal node `ntaxa`
    edges = [[0, ntaxa], [1, ntaxa], [2, ntaxa]]
    pendant = [0, 1, 2]  # edge index of the pendant edge of leaf i
    next_inner = ntaxa + 1
    for leaf in range(3, ntaxa):
        e = pendant[int(rng.integers(0, leaf))]
        u, v = edges[e]  # u is the leaf
        w = next_inner
        next_inner += 1
        edges[e] = [u, w]
        edges.append([w, v])
        edges.append([leaf, w])
        pendant.append(len(edges) - 1)
    edges = np.array(edges, dtype=np.int64)
    lengths = np.clip(rng.exponential(mean_len, size=len(edges)), min_len, max_len)
    # relabel leaves randomly so the topology is not correlated with row order
    perm = rng.permutation(ntaxa)
    mapping = np.arange(2 * ntaxa - 2)
    mapping[:ntaxa] = perm
    edges = mapping[edges]
    names = ["T%d" % i for i in range(ntaxa)]
    return Tree(ntaxa, edges, lengths, names)


def caterpillar_tree(ntaxa: int, seed: int = 1, mean_len: float = 0.1) -> Tree:
    """Maximally unbalanced tree (deepest traversal; stresses scaling counts and level batching)."""
    rng = np.random.default_rng(seed)
    edges = [[0, ntaxa]]
    for i in range(1, ntaxa - 1):
        inner = ntaxa + i - 1
        edges.append([i, inner])
        if i < ntaxa - 2:
            edges.append([inner, inner + 1])
    edges.append([ntaxa - 1, 2 * ntaxa - 3])
    edges = np.array(edges, dtype=np.int64)
    lengths = np.clip(rng.exponential(mean_len, size=len(edges)), 1e-6, 10.0)
    return Tree(ntaxa, edges, lengths, ["T%d" % i for i in range(ntaxa)])


def simulate_alignment(tree: Tree, model, nsites: int, seed: int = 2, batch: int = 1 << 18) -> np.ndarray:
    """Evolve `nsites` sites down the tree (AliSim-style). Returns uint8 states [ntaxa][nsites].

    Each site draws a Gamma category (prob cat_prop) and a mixture class (prob mix_weight); invariant sites
    (prob pinvar) copy the root state. P(t) = expm(Q r t) is built from the model's eigensystem.
    """
    rng = np.random.default_rng(seed)
    n, ncat, nmix = model.nstates, model.ncat, model.nmix
    ev = model.eval.reshape(nmix, n)
    U = model.evec.reshape(nmix, n, n)
    Ui = model.inv_evec.reshape(nmix, n, n)
    freq = model.state_freq.reshape(nmix, n)
    adj = tree.adjacency()
    root = tree.ntaxa  # an internal node
    # pre-order list of (node, dad, branch)
    order = []
    stack = [(root, -1, -1)]
    while stack:
        nd, dd, b = stack.pop()
        order.append((nd, dd, b))
        for v, bb in adj[nd]:
            if v != dd:
                stack.append((v, nd, bb))
    out = np.empty((tree.ntaxa, nsites), dtype=np.uint8)
    cat_p = model.cat_prop / model.cat_prop.sum()
    for s0 in range(0, nsites, batch):
        ns = min(batch, nsites - s0)
        cat = rng.choice(ncat, size=ns, p=cat_p)
        mix = rng.choice(nmix, size=ns, p=model.mix_weight / model.mix_weight.sum()) if nmix > 1 else np.zeros(ns, dtype=np.int64)
        invariant = rng.random(ns) < model.pinvar if model.pinvar > 0 else np.zeros(ns, dtype=bool)
        states = {}
        # root states from the class frequencies
        u = rng.random(ns)
        cdf = np.cumsum(freq, axis=1)
        cdf[:, -1] = 1.0
        rs = np.empty(ns, dtype=np.int64)
        for m in range(nmix):
            sel = mix == m
            rs[sel] = np.minimum((u[sel, None] > cdf[m][None, :]).sum(axis=1), n - 1)
        states[root] = rs
        key = mix * ncat + cat
        for nd, dd, b in order[1:]:
            t = tree.lengths[b]
            parent = states[dd]
            child = np.empty(ns, dtype=np.int64)
            u = rng.random(ns)
            for m in range(nmix):
                for c in range(ncat):
                    sel = np.nonzero(key == m * ncat + c)[0]
                    if sel.size == 0:
                        continue
                    P = (U[m] * np.exp(ev[m] * model.cat_rate[c] * t)[None, :]) @ Ui[m]
                    P = np.clip(P, 0.0, None)
                    cdfP = np.cumsum(P, axis=1)
                    cdfP /= cdfP[:, -1:]
                    rows = cdfP[parent[sel]]
                    child[sel] = np.minimum((u[sel, None] > rows).sum(axis=1), n - 1)
            child[invariant] = parent[invariant]
            states[nd] = child
            if nd < tree.ntaxa:
                out[nd, s0:s0 + ns] = child.astype(np.uint8)
            # free internal states no longer needed is unnecessary at these sizes
    return out


def compress_patterns(states: np.ndarray):
    """Site -> pattern compression (distinct columns, first-occurrence order; alignment.cpp:1092-1178).

    Returns (pattern_states uint8 [ntaxa][nptn], ptn_freq float64 [nptn]).
    """
    ntaxa, nsites = states.shape
    cols = np.ascontiguousarray(states.T)
    view = cols.view(np.dtype((np.void, cols.dtype.itemsize * ntaxa))).ravel()
    _, first, inverse, counts = np.unique(view, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")
    pat = cols[first[order]].T.copy()
    freq = counts[order].astype(np.float64)
    return pat, freq


def random_patterns(ntaxa: int, nptn: int, nstates: int, seed: int = 2, p_unknown: float = 0.0,
                    state_unknown: int | None = None, conserved: float = 0.7) -> np.ndarray:
    """Cheap synthetic tip states for throughput runs at sizes where tree simulation would dominate set-up:
    each pattern copies a random 'ancestral' state with prob `conserved`, else a uniform state. The kernels'
    work is independent of the state values, so this only affects how often scaling triggers."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, nstates, size=nptn, dtype=np.uint8)
    out = np.empty((ntaxa, nptn), dtype=np.uint8)
    for t in range(ntaxa):
        mut = rng.random(nptn) >= conserved
        row = base.copy()
        row[mut] = rng.integers(0, nstates, size=int(mut.sum()), dtype=np.uint8)
        if p_unknown > 0:
            unk = rng.random(nptn) < p_unknown
            row[unk] = state_unknown
        out[t] = row
    return out


def evolve_patterns_fast(tree: Tree, nstates: int, nptn: int, seed: int = 2, rate_spread: bool = True) -> np.ndarray:
    """Fast tree-structured synthetic tip states for bench-size inputs (1000 taxa x 1M patterns in a few seconds):
    states are evolved down the tree with an equal-rates (Jukes-Cantor-like) process, P(change on a branch of
    length t) = (1 - 1/n)(1 - exp(-n/(n-1) t r_site)), with 4 site-rate classes when `rate_spread`.  The likelihood
    kernels' work does not depend on the state values; tree-structured data only makes the amount of 2^-256
    rescaling realistic.  Returns uint8 [ntaxa][nptn]."""
    rng = np.random.default_rng(seed)
    adj = tree.adjacency()
    root = tree.ntaxa
    site_rate = np.ones(nptn, dtype=np.float32)
    if rate_spread:
        site_rate = np.array([0.03, 0.25, 0.82, 2.9], dtype=np.float32)[rng.integers(0, 4, size=nptn)]
    out = np.empty((tree.ntaxa, nptn), dtype=np.uint8)
    states = {root: rng.integers(0, nstates, size=nptn, dtype=np.uint8)}
    stack = [(root, -1)]
    k = nstates / (nstates - 1.0)
    while stack:
        nd, dd = stack.pop()
        cur = states.pop(nd)
        if nd < tree.ntaxa:
            out[nd] = cur
            continue
        for v, b in adj[nd]:
            if v == dd:
                continue
            pchange = (1.0 - 1.0 / nstates) * (1.0 - np.exp(-k * np.float32(tree.lengths[b]) * site_rate))
            mut = rng.random(nptn, dtype=np.float32) < pchange
            child = cur.copy()
            nmut = int(mut.sum())
            if nmut:
                # a change always lands on a different state
                child[mut] = (cur[mut] + rng.integers(1, nstates, size=nmut, dtype=np.uint8)) % nstates
            states[v] = child
            stack.append((v, nd))
    return out
