"""Route C: an independent brute-force Felsenstein pruning (numpy), used to pin the oracle.

No eigen-space partials, no reference-style 2^-256 scaling, no shared code with oracle/ or the product:
conditional likelihoods live in state space, transition matrices are P(t) = expm(Q r t) (scipy), and
underflow is handled by per-(pattern, category) max-normalisation with a log accumulator.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm

from iqtree2_b200.models import state_indicator


def _tip_matrix(model):
    su = model.state_unknown
    return np.stack([state_indicator(s, model.nstates, su, model.seq_type) for s in range(su + 1)])  # [SU+1][n]


def pattern_loglik(tree, model, tips, branch, length=None, ptn_invar=None, derivatives=False):
    """Per-pattern log-likelihood evaluated across `branch` (optionally with a different length).

    Returns logl[nptn] (and dL/L, d2L/L per pattern if derivatives)."""
    n, ncat, nmix = model.nstates, model.ncat, model.nmix
    nptn = tips.shape[1]
    adj = tree.adjacency()
    nt = tree.ntaxa
    ind = _tip_matrix(model)
    u, v = int(tree.edges[branch][0]), int(tree.edges[branch][1])
    t_branch = tree.lengths[branch] if length is None else length
    ptn_invar = np.zeros(nptn) if ptn_invar is None else ptn_invar

    lik = np.zeros(nptn)
    d1 = np.zeros(nptn)
    d2 = np.zeros(nptn)
    # accumulate sum_c w_c * L_c with a common log-scale per pattern: first collect per-class values+logs
    vals, logs, vals1, vals2 = [], [], [], []
    for m in range(nmix):
        Q = model.Q[m]
        pi = model.state_freq.reshape(nmix, n)[m]
        for c in range(ncat):
            r = model.cat_rate[c]
            w = model.cat_prop[c] * model.mix_weight[m]

            def clv(node, dad):
                """returns (L[nptn][n], logscale[nptn]) of subtree `node` seen from `dad` (iterative)."""
                result = {}
                stack = [(node, dad, False)]
                while stack:
                    nd, dd, done = stack.pop()
                    if nd < nt:
                        result[(nd, dd)] = (ind[tips[nd]].astype(np.float64), np.zeros(nptn))
                        continue
                    kids = [(x, b) for x, b in adj[nd] if x != dd]
                    if not done:
                        stack.append((nd, dd, True))
                        for x, b in kids:
                            stack.append((x, nd, False))
                        continue
                    L = np.ones((nptn, n))
                    ls = np.zeros(nptn)
                    for x, b in kids:
                        Lk, lsk = result.pop((x, nd))
                        P = expm(Q * (r * tree.lengths[b]))
                        L = L * (Lk @ P.T)
                        ls = ls + lsk
                    mx = L.max(axis=1)
                    mx[mx == 0] = 1.0
                    L = L / mx[:, None]
                    ls = ls + np.log(mx)
                    result[(nd, dd)] = (L, ls)
                return result[(node, dad)]

            Lu, lsu = clv(u, v)
            Lv, lsv = clv(v, u)
            P = expm(Q * (r * t_branch))
            PLv = Lv @ P.T
            val = ((Lu * pi[None, :]) * PLv).sum(axis=1) * w
            vals.append(val)
            logs.append(lsu + lsv)
            if derivatives:
                QP = (Q * r) @ P
                Q2P = (Q * r) @ QP
                vals1.append(((Lu * pi[None, :]) * (Lv @ QP.T)).sum(axis=1) * w)
                vals2.append(((Lu * pi[None, :]) * (Lv @ Q2P.T)).sum(axis=1) * w)
    logs = np.array(logs)
    mxlog = logs.max(axis=0)
    scale = np.exp(logs - mxlog[None, :])
    lik = (np.array(vals) * scale).sum(axis=0)
    # invariant-site term lives on the absolute scale
    total = lik + ptn_invar * np.exp(-mxlog)
    logl = np.log(total) + mxlog
    if not derivatives:
        return logl
    d1 = (np.array(vals1) * scale).sum(axis=0)
    d2 = (np.array(vals2) * scale).sum(axis=0)
    return logl, d1 / total, d2 / total


def tree_loglik(tree, model, tips, ptn_freq, branch=0, length=None, ptn_invar=None):
    return float((pattern_loglik(tree, model, tips, branch, length, ptn_invar) * ptn_freq).sum())


def tree_derivatives(tree, model, tips, ptn_freq, branch, length=None, ptn_invar=None):
    """(logL, df, ddf) with df = sum f * L'/L and ddf = sum f * (L''/L - (L'/L)^2)  (phylokernelnew.h:2494-2505)."""
    logl, a, b = pattern_loglik(tree, model, tips, branch, length, ptn_invar, derivatives=True)
    return float((logl * ptn_freq).sum()), float((ptn_freq * a).sum()), float((ptn_freq * (b - a * a)).sum())
