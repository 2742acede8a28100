"""CPU tier: mixed branch lengths (heterotachy, `+H`; PhyloTree::isMixlen()) — the restatement (oracle/lk_oracle_mixlen.h + len_of() in
lk_oracle.c) against the compiled reference run as a mixlen tree (oracle/ref_harness.cpp MixlenTree: PhyloNodeMixlen /
PhyloNeighborMixlen nodes, the reference's own kernels), against a brute-force pruning with one tree per class, and against finite
differences of its own log-likelihood."""
import dataclasses

import numpy as np
import pytest

import bruteforce
from cases import make_case
from iqtree2_b200 import plan
from oracle.oracle_engine import OracleEngine

FACTORS = {1: [1.3], 2: [0.6, 1.7], 3: [0.4, 1.0, 2.2], 4: [0.35, 0.8, 1.4, 2.6]}


def make_mixlen_case(kind="dna", ncat=4, ntaxa=10, nsites=300, seed=3, **kw):
    """a reversible model whose ncat categories are heterotachy classes: proportions are free weights, rates as given (RateHeterotachy
    has rate 1 for every class; the kernels multiply by getRate(c) all the same, so the tests keep general rates)"""
    from iqtree2_b200 import models
    pinvar = kw.pop("pinvar", 0.0)
    if kind == "dna":
        m = models.make_gtr_gamma(ncat=ncat, pinvar=pinvar)
    elif kind == "aa":
        m = models.make_lg_gamma(ncat=ncat, pinvar=pinvar)
    elif kind == "aa_mix":
        m = models.make_lg_c60_gamma(ncat=ncat, nclasses=3)
    else:
        m = models.make_random_reversible(int(kind[1:]), ncat=ncat, seed=seed, pinvar=pinvar)
    tree, model, pat, freq, invar = make_case(kind, ntaxa=ntaxa, nsites=nsites, seed=seed, model=m, **kw)
    rng = np.random.default_rng(seed + 11)
    w = rng.dirichlet(np.full(model.ncat, 4.0)) * (1.0 - model.pinvar)
    model = dataclasses.replace(model, cat_prop=w, mixlen_factor=np.array(FACTORS[model.ncat]))
    return tree, model, pat, freq, invar


def mixlen_table(tree, model):
    return np.outer(np.asarray(tree.lengths, dtype=np.float64), model.mixlen_factor)


def mixlen_engine(cls, tree, model, pat, freq, invar, safe=False, **kw):
    e = cls(model, tree.ntaxa, pat.shape[1], safe_numeric=safe, **kw)
    e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(invar)
    e.set_mixlen(mixlen_table(tree, model))
    return e


def plan_rows(tree, br):
    """ops whose child lengths are branch ROWS of the mixlen table"""
    return plan.plan_branch(tree, br, lengths=np.arange(len(tree.lengths), dtype=np.float64))


def brute_force_mixlen(tree, model, pat, branch=0):
    """sum over classes of prop_c * L(pattern | tree with lengths * factor_c, rate_c): per-pattern log-likelihood"""
    tot = np.zeros(pat.shape[1])
    for c in range(model.ncat):
        t = dataclasses.replace(tree, lengths=np.asarray(tree.lengths) * model.mixlen_factor[c]) if dataclasses.is_dataclass(tree) else None
        if t is None:
            import copy
            t = copy.copy(tree)
            t.lengths = np.asarray(tree.lengths) * model.mixlen_factor[c]
        mc = dataclasses.replace(model, cat_rate=model.cat_rate[c:c + 1], cat_prop=np.array([1.0]), pinvar=0.0, mixlen_factor=None)
        tot += model.cat_prop[c] * np.exp(bruteforce.pattern_loglik(t, mc, pat, branch))
    return tot


@pytest.mark.parametrize("kind,ncat", [("dna", 4), ("dna", 2), ("aa", 3)])
@pytest.mark.parametrize("safe", [False, True])
def test_mixlen_oracle_matches_brute_force_and_finite_differences(kind, ncat, safe):
    tree, model, pat, freq, invar = make_mixlen_case(kind, ncat, ntaxa=7, nsites=40, seed=5)
    o = mixlen_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    want = np.log(brute_force_mixlen(tree, model, pat) + invar)
    for br in (0, 3, plan.pendant_branch_of(tree, 2)):
        ops, a, b, row = plan_rows(tree, br)
        ll = o.branch_logl(ops, a, b, row)
        np.testing.assert_allclose(o.get_pattern_lh(), want, rtol=1e-10)
        g, H, ll2 = o.branch_derv_mixlen(ops[:0], a, b, row)
        assert abs(ll2 - ll) < 1e-9 * abs(ll)
        # finite differences over the K lengths of the evaluation branch: the theta buffer serves any table row
        tab = mixlen_table(tree, model)
        K, h = model.ncat, 1e-5

        def at(delta):
            t = tab.copy()
            t[int(row)] += delta
            o.set_mixlen(t)
            return o.logl_from_buffer(row)
        for i in range(K):
            ei = np.zeros(K); ei[i] = h
            fd = (at(ei) - at(-ei)) / (2 * h)
            assert abs(fd - g[i]) < 1e-5 * max(1.0, abs(g[i])), (i, fd, g[i])
            for j in range(K):
                ej = np.zeros(K); ej[j] = h
                fdd = (at(ei + ej) - at(ei - ej) - at(-ei + ej) + at(-ei - ej)) / (4 * h * h)
                assert abs(fdd - H[i, j]) < 5e-3 * max(1.0, abs(H[i, j])), (i, j, fdd, H[i, j])
        o.set_mixlen(tab)
        o.theta_reset()


def _ref_available():
    from oracle import ref_engine
    return ref_engine.available()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind,ncat,safe,pinvar", [("dna", 4, False, 0.0), ("dna", 4, True, 0.1), ("aa", 4, False, 0.0), ("aa", 2, True, 0.0),
                                                   ("dna", 3, False, 0.0), ("n5", 3, True, 0.0)])
def test_mixlen_oracle_matches_the_reference_kernel(kind, ncat, safe, pinvar):
    from oracle import ref_engine
    tree, model, pat, freq, invar = make_mixlen_case(kind, ncat, ntaxa=12, nsites=260, seed=13, pinvar=pinvar)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=2)
    o = mixlen_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for leaf in (0, 4, 9):
        ops, a, b, row = plan_rows(tree, plan.pendant_branch_of(tree, leaf))
        ll_r, plh_r, cat_r = r.branch_logl(leaf, want_cat=True)
        ll_o = o.branch_logl(ops, a, b, row)
        assert abs(ll_r - ll_o) < 1e-9 * abs(ll_r), (leaf, ll_r, ll_o)
        np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12, atol=0)
        scale = np.abs(cat_r).max(axis=1, keepdims=True)
        np.testing.assert_allclose(o.get_pattern_lh_cat() / scale, cat_r / scale, rtol=0, atol=1e-12)
        g_r, H_r, l_r = r.branch_derv_mixlen(leaf)
        g_o, H_o, l_o = o.branch_derv_mixlen(ops[:0], a, b, row)
        np.testing.assert_allclose(g_o, g_r, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(H_o, H_r, rtol=1e-9, atol=1e-8)
        assert abs(l_o - l_r) < 1e-9 * abs(l_r)
        assert abs(o.logl_from_buffer(row) - r.logl_from_buffer()) < 1e-9 * abs(ll_r)
        for cur in range(model.ncat):
            np.testing.assert_allclose(o.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), r.branch_derv_mixlen_cur(leaf, cur), rtol=1e-9, atol=1e-9)
        o.theta_reset()
    r.close()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libiqref.so not built")
def test_mixlen_with_a_mixture_model_matches_the_reference():
    """classes x mixture components (ncat_mix = ncat * nmix): cat_id[c] = c % ncat picks the length"""
    from oracle import ref_engine
    tree, model, pat, freq, invar = make_mixlen_case("aa_mix", 2, ntaxa=9, nsites=120, seed=21)
    assert model.nmix > 1
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=True, nthreads=1)
    o = mixlen_engine(OracleEngine, tree, model, pat, freq, invar, True)
    ops, a, b, row = plan_rows(tree, plan.pendant_branch_of(tree, 3))
    ll_r, plh_r = r.branch_logl(3)
    assert abs(o.branch_logl(ops, a, b, row) - ll_r) < 1e-9 * abs(ll_r)
    np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12, atol=0)
    g_r, H_r, l_r = r.branch_derv_mixlen(3)
    g_o, H_o, l_o = o.branch_derv_mixlen(ops[:0], a, b, row)
    np.testing.assert_allclose(g_o, g_r, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(H_o, H_r, rtol=1e-9, atol=1e-8)
    for cur in range(model.ncat):
        np.testing.assert_allclose(o.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), r.branch_derv_mixlen_cur(3, cur), rtol=1e-9, atol=1e-9)
    r.close()
