"""CPU: +ASC (Lewis' ascertainment-bias correction; ASC_VARIANT, tree/phylokernelnew.h:3076-3091, 3198-3221, 2506-2535,
2618-2626, 3342-3356, 3412-3426) in the oracle against the compiled reference with ModelFactory::ASC_type = ASC_VARIANT and
the unobserved constant patterns supplied.  The harness alignment registers one site per pattern, so getNSite() = nptn there.
And +ASC_MIS (Holder's correction for missing data; ASC_VARIANT_MISSING, :3175-3197, 2580-2617, 3385-3410) with the
nstates x padded-nptn block of constant patterns of Alignment::getUnobservedConstPatterns (alignment/alignment.cpp:4158-4180)."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200 import plan
from oracle import ref_engine
from oracle.oracle_engine import OracleEngine


def asc_case(kind="bin", ntaxa=12, nsites=400, seed=3, mean_len=0.25):
    """variable sites only + the nstates unobserved constant patterns appended (Alignment::getUnobservedConstPatterns)"""
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=nsites, seed=seed, gaps=0.0, ambig=0.0, mean_len=mean_len)
    variable = ~np.all(pat == pat[0:1, :], axis=0)
    pat, freq = pat[:, variable], freq[variable]
    n = model.nstates
    unobs = np.repeat(np.arange(n, dtype=np.uint8)[:, None], ntaxa, axis=1)   # [n][ntaxa]
    pat_all = np.concatenate([pat, unobs.T], axis=1)
    freq_all = np.concatenate([freq, np.zeros(n)])
    return tree, model, pat, freq, unobs, np.ascontiguousarray(pat_all), freq_all


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind,safe", [("bin", True), ("dna", False), ("dna", True), ("n5mix", True)])
def test_oracle_asc_lewis_matches_reference(kind, safe):
    tree, model, pat, freq, unobs, pat_all, freq_all = asc_case(kind)
    nobs, U = pat.shape[1], unobs.shape[0]
    o = OracleEngine(model, tree.ntaxa, nobs + U, safe_numeric=safe)
    o.set_tip_states(pat_all); o.set_pattern_freq(freq_all); o.set_pattern_invar(np.zeros(nobs + U))
    o.set_ascertainment(1, U, nobs)           # nsite = what the harness alignment reports
    o.set_reference_quirks(1)                 # dotProductTriple's double-counted element for odd nstates
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll = o.branch_logl(ops, a, b, L)
    ref_engine.set_asc_next(1, unobs)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe)
    ll_r, plh_r = r.branch_logl(0)
    assert abs(ll - ll_r) < 1e-8 * max(1.0, abs(ll_r) * 1e-3), (ll, ll_r)
    np.testing.assert_allclose(o.get_pattern_lh()[:nobs], plh_r, rtol=1e-12, atol=0)
    np.testing.assert_allclose(o.branch_derv(ops[:0], a, b, L), r.branch_derv(0), rtol=1e-9, atol=1e-9)
    assert abs(o.logl_from_buffer(L) - r.logl_from_buffer()) < 1e-8 * max(1.0, abs(ll_r) * 1e-3)
    # the correction matters: without it the log-likelihood of the variable sites alone is different
    o2 = OracleEngine(model, tree.ntaxa, nobs, safe_numeric=safe)
    o2.set_tip_states(pat); o2.set_pattern_freq(freq); o2.set_pattern_invar(np.zeros(nobs))
    assert abs(o2.branch_logl(ops, a, b, L) - ll) > 1e-3
    r.close()


def holder_case(kind="dna", ntaxa=12, nsites=400, seed=3, mean_len=0.25, gaps=0.15, pad=4):
    """variable sites with missing data + for every state s and every (padded) pattern p the constant pattern that keeps p's gaps
    (alignment/alignment.cpp:4158-4180: the block is nstates x get_safe_upper_limit(nptn), padding patterns all unknown)"""
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=nsites, seed=seed, gaps=gaps, ambig=0.0, mean_len=mean_len)
    n, unk = model.nstates, model.state_unknown
    known = pat < n
    first = np.where(known, pat, 255).min(axis=0)
    last = np.where(known, pat, 0).max(axis=0)
    variable = (first != last) & (known.sum(axis=0) >= 2)
    pat, freq = pat[:, variable], freq[variable]
    nobs = pat.shape[1]
    step = (nobs + pad - 1) // pad * pad
    unobs = np.full((n, step, ntaxa), unk, dtype=np.uint8)
    for s in range(n):
        unobs[s, :nobs, :] = np.where(pat.T < n, s, unk)
    unobs = unobs.reshape(n * step, ntaxa)
    pat_all = np.ascontiguousarray(np.concatenate([pat, unobs.T], axis=1))
    freq_all = np.concatenate([freq, np.zeros(n * step)])
    return tree, model, pat, freq, unobs, pat_all, freq_all


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind,safe", [("bin", True), ("dna", False), ("dna", True), ("aa", False)])
def test_oracle_asc_holder_matches_reference(kind, safe):
    tree, model, pat, freq, unobs, pat_all, freq_all = holder_case(kind, nsites=300 if kind != "aa" else 120)
    nobs, U = pat.shape[1], unobs.shape[0]
    o = OracleEngine(model, tree.ntaxa, nobs + U, safe_numeric=safe)
    o.set_tip_states(pat_all); o.set_pattern_freq(freq_all); o.set_pattern_invar(np.zeros(nobs + U))
    o.set_ascertainment(2, U, nobs)
    o.set_reference_quirks(1)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll = o.branch_logl(ops, a, b, L)
    ref_engine.set_asc_next(2, unobs)         # ASC_VARIANT_MISSING
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe)
    ll_r, plh_r = r.branch_logl(0)
    assert abs(ll - ll_r) < 1e-8 * max(1.0, abs(ll_r) * 1e-3), (ll, ll_r)
    np.testing.assert_allclose(o.get_pattern_lh()[:nobs], plh_r, rtol=1e-12, atol=0)
    np.testing.assert_allclose(o.branch_derv(ops[:0], a, b, L), r.branch_derv(0), rtol=1e-9, atol=1e-9)
    assert abs(o.logl_from_buffer(L) - r.logl_from_buffer()) < 1e-8 * max(1.0, abs(ll_r) * 1e-3)
    # not Lewis' correction: one global term over the fully observed constant patterns gives another value here
    o2 = OracleEngine(model, tree.ntaxa, nobs, safe_numeric=safe)
    o2.set_tip_states(pat); o2.set_pattern_freq(freq); o2.set_pattern_invar(np.zeros(nobs))
    assert abs(o2.branch_logl(ops, a, b, L) - ll) > 1e-3
    r.close()
