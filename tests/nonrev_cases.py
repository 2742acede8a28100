"""Shared cases for the non-reversible kernel (tree/phylokernelnonrev.h): a rooted tree = an unrooted treesim.Tree with ntaxa + 1
leaves whose LAST leaf is the virtual root leaf (MTree::readTree attaches it with id = leafNum - 1, tree/mtree.cpp:742-754), random
tip patterns for the ntaxa real rows, and an independent brute-force pruning (state space, P = expm(Q r t), log-scaled)."""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

from iqtree2_b200 import models, treesim
from iqtree2_b200.models import state_indicator


def make_nonrev_case(nstates=4, ncat=4, nmix=1, ntaxa=9, nptn=200, seed=3, tree_kind="yule", mean_len=0.1, p_unknown=0.05, pinvar=0.0,
                     state_unknown=None):
    """-> (tree with ntaxa + 1 leaves, model, pat [ntaxa][nptn], freq, invar)"""
    su = state_unknown
    if su is None:
        su = models.DNA_STATE_UNKNOWN if nstates == 4 else (models.PROTEIN_STATE_UNKNOWN if nstates == 20 else nstates)
    model = models.make_nonreversible(nstates, ncat=ncat, nmix=nmix, seed=seed, pinvar=pinvar, state_unknown=su, rooted=True)
    mk = treesim.yule_tree if tree_kind == "yule" else treesim.caterpillar_tree
    tree = mk(ntaxa + 1, seed=seed, mean_len=mean_len)
    rng = np.random.default_rng(seed + 11)
    # evolve down the rooted tree so that patterns carry signal: start at the root leaf with the stationary distribution of class 0
    n = nstates
    adj = tree.adjacency()
    root = ntaxa
    states = np.zeros((tree.nnodes, nptn), dtype=np.int64)
    pi = model.state_freq.reshape(nmix, n)[0]
    states[root] = rng.choice(n, size=nptn, p=pi / pi.sum())
    rate = model.cat_rate[rng.integers(0, ncat, size=nptn)]
    stack = [(root, -1)]
    while stack:
        u, dad = stack.pop()
        for v, b in adj[u]:
            if v == dad:
                continue
            if u == root:
                states[v] = states[u]        # the branch to the root leaf carries no evolution
            else:
                for r in np.unique(rate):
                    P = np.real(expm(model.Q[0] * (r * tree.lengths[b])))
                    P = np.clip(P, 0, None)
                    P /= P.sum(axis=1, keepdims=True)
                    idx = np.nonzero(rate == r)[0]
                    cum = np.cumsum(P[states[u][idx]], axis=1)
                    states[v][idx] = (rng.random(len(idx))[:, None] > cum).sum(axis=1).clip(0, n - 1)
            stack.append((v, u))
    pat = states[:ntaxa].astype(np.uint8)
    mask = rng.random(pat.shape) < p_unknown
    pat[mask] = su
    if nstates == 4 and su == models.DNA_STATE_UNKNOWN:   # a few IUPAC ambiguity codes (states 4..17)
        amb = rng.random(pat.shape) < 0.02
        pat[amb] = rng.integers(4, 18, size=int(amb.sum()))
    freq = rng.integers(1, 4, size=nptn).astype(np.float64)
    if pinvar > 0:
        from cases import ptn_invar_for
        invar = ptn_invar_for(model, pat)
    else:
        invar = np.zeros(nptn)
    return tree, model, pat, freq, invar


def brute_force_rooted(tree, model, pat, invar=None):
    """per-pattern log-likelihood of the ROOTED tree: sum_x pi_x L_top(x) with L pruned away from the root leaf, P(parent -> child)"""
    n, ncat, nmix = model.nstates, model.ncat, model.nmix
    nt = tree.ntaxa
    root = nt - 1
    nptn = pat.shape[1]
    adj = tree.adjacency()
    su = model.state_unknown
    ind = np.stack([state_indicator(s, n, su, model.seq_type) for s in range(su + 1)])
    invar = np.zeros(nptn) if invar is None else invar
    vals, logs = [], []
    for m in range(nmix):
        Q = model.Q[m]
        pi = model.state_freq.reshape(nmix, n)[m]
        for c in range(ncat):
            r = model.cat_rate[c]
            w = model.cat_prop[c] * model.mix_weight[m]

            def clv(node, dad):
                if node < nt:
                    return ind[pat[node]].astype(np.float64), np.zeros(nptn)
                L = np.ones((nptn, n))
                ls = np.zeros(nptn)
                for x, b in adj[node]:
                    if x == dad:
                        continue
                    Lk, lsk = clv(x, node)
                    P = np.real(expm(Q * (r * tree.lengths[b])))
                    L = L * (Lk @ P.T)
                    ls = ls + lsk
                mx = L.max(axis=1)
                mx[mx == 0] = 1.0
                return L / mx[:, None], ls + np.log(mx)

            top = adj[root][0][0]
            L, ls = clv(top, root)
            vals.append((L * pi[None, :]).sum(axis=1) * w)
            logs.append(ls)
    logs = np.array(logs)
    vals = np.array(vals)
    mx = logs.max(axis=0)
    lik = (vals * np.exp(logs - mx[None, :])).sum(axis=0)
    with np.errstate(divide="ignore"):
        return np.log(lik + invar * np.exp(-mx)) + mx
