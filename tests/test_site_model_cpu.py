"""CPU tier: site-specific models (ModelSet / PMSF; the SITE_MODEL template branch of tree/phylokernelnew.h) — the restatement
(oracle/lk_oracle_site.h) against the compiled reference (its own SITE_MODEL instantiations, Vec4d AVX+FMA) and against a brute-force
pruning with one rate matrix per pattern."""
import numpy as np
import pytest

import bruteforce
from iqtree2_b200 import models, plan, treesim
from oracle.oracle_engine import OracleEngine


def make_site_case(nstates=4, ncat=4, ntaxa=10, nptn=120, seed=3, pinvar=0.0, tree_kind="yule", mean_len=0.1, p_unknown=0.05):
    su = models.DNA_STATE_UNKNOWN if nstates == 4 else (models.PROTEIN_STATE_UNKNOWN if nstates == 20 else nstates)
    model = models.make_site_specific(nstates, nptn, ncat=ncat, seed=seed, pinvar=pinvar, state_unknown=su)
    tree = (treesim.yule_tree if tree_kind == "yule" else treesim.caterpillar_tree)(ntaxa, seed=seed, mean_len=mean_len)
    rng = np.random.default_rng(seed + 5)
    pat = treesim.evolve_patterns_fast(tree, nstates, nptn, seed=seed + 1).astype(np.uint8)
    pat[rng.random(pat.shape) < p_unknown] = su
    if nstates == 4:
        amb = rng.random(pat.shape) < 0.02
        pat[amb] = rng.integers(4, 18, size=int(amb.sum()))
    elif nstates == 20:
        amb = rng.random(pat.shape) < 0.01
        pat[amb] = rng.integers(20, 23, size=int(amb.sum()))
    freq = rng.integers(1, 4, size=nptn).astype(np.float64)
    if pinvar > 0:
        from cases import ptn_invar_for
        invar = ptn_invar_for(model, pat)
    else:
        invar = np.zeros(nptn)
    return tree, model, pat, freq, invar


def engine(cls, tree, model, pat, freq, invar, safe, **kw):
    e = cls(model, tree.ntaxa, pat.shape[1], safe_numeric=safe, **kw)
    e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(invar)
    return e


def brute_force_site(tree, model, pat, branch=0):
    """per-pattern logL with pattern p's own rate matrix (bruteforce.pattern_loglik on one column at a time)"""
    out = np.zeros(pat.shape[1])
    n = model.nstates
    for p in range(pat.shape[1]):
        mp = models._pack(n, [(model.site["eval"][p], model.site["evec"][p], model.site["inv_evec"][p], model.site["Q"][p], model.site["pi"][p])], [1.0],
                          model.cat_rate, model.cat_prop, model.pinvar, model.state_unknown, model.seq_type)
        out[p] = bruteforce.pattern_loglik(tree, mp, pat[:, p:p + 1], branch)[0]
    return out


@pytest.mark.parametrize("nstates,ncat", [(4, 4), (5, 2), (20, 1)])
@pytest.mark.parametrize("safe", [False, True])
def test_site_oracle_matches_brute_force(nstates, ncat, safe):
    tree, model, pat, freq, invar = make_site_case(nstates, ncat, ntaxa=7, nptn=24, seed=5)
    want = brute_force_site(tree, model, pat)
    o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for br in (0, 3, plan.pendant_branch_of(tree, 2)):
        ops, a, b, L = plan.plan_branch(tree, br)
        ll = o.branch_logl(ops, a, b, L)
        np.testing.assert_allclose(o.get_pattern_lh(), want, rtol=1e-10)
        assert abs(ll - float((want * freq).sum())) < 1e-8
        # derivatives of its own logL; the theta buffer serves another length
        df, ddf = o.branch_derv(ops[:0], a, b, L)
        h = 1e-5
        lp, lm, l0 = o.logl_from_buffer(L + h), o.logl_from_buffer(L - h), o.logl_from_buffer(L)
        assert abs(l0 - ll) < 1e-9 * abs(ll)
        assert abs(df - (lp - lm) / (2 * h)) < 1e-5 * max(1.0, abs(df))
        assert abs(ddf - (lp - 2 * l0 + lm) / (h * h)) < 2e-3 * max(1.0, abs(ddf))
        o.theta_reset()


def _ref_available():
    from oracle import ref_engine
    return ref_engine.available()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("nstates,ncat,safe,pinvar", [(4, 4, False, 0.0), (4, 4, True, 0.1), (20, 4, False, 0.0), (20, 2, True, 0.0), (5, 4, True, 0.0), (5, 3, False, 0.0)])
def test_site_oracle_matches_the_reference_kernel(nstates, ncat, safe, pinvar):
    from oracle import ref_engine
    tree, model, pat, freq, invar = make_site_case(nstates, ncat, ntaxa=12, nptn=203, seed=13, pinvar=pinvar)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=2)
    o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for leaf in (0, 4, 9):
        ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, leaf))
        ll_r, plh_r, cat_r = r.branch_logl(leaf, want_cat=True)
        ll_o = o.branch_logl(ops, a, b, L)
        assert abs(ll_r - ll_o) < 1e-9 * abs(ll_r), (leaf, ll_r, ll_o)
        np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12, atol=0)
        scale = np.abs(cat_r).max(axis=1, keepdims=True)
        np.testing.assert_allclose(o.get_pattern_lh_cat() / scale, cat_r / scale, rtol=0, atol=1e-12)
        np.testing.assert_allclose(o.branch_derv(ops[:0], a, b, L), r.branch_derv(leaf), rtol=1e-9)
        assert abs(o.logl_from_buffer(L) - r.logl_from_buffer()) < 1e-9 * abs(ll_r)
        o.theta_reset()
    r.close()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libiqref.so not built")
def test_site_oracle_matches_the_reference_on_a_deep_tree():
    from oracle import ref_engine
    tree, model, pat, freq, invar = make_site_case(4, 4, ntaxa=160, nptn=64, seed=7, tree_kind="caterpillar", mean_len=0.8, p_unknown=0.0)
    for safe in (False, True):
        r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=1)
        o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
        for leaf in (0, 80):
            ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, leaf))
            ll_r, plh_r = r.branch_logl(leaf)
            o.branch_logl(ops, a, b, L)
            assert max(o.get_scale_num(int(op["dst_slot"])).max() for op in ops) >= 1, "case does not rescale"
            np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-11, atol=0)
        r.close()
