"""GPU: +ASC (Lewis' ascertainment-bias correction) through the C ABI (b200lk_set_ascertainment) against the CPU oracle
(pinned against the compiled reference in test_asc_cpu.py): corrected logL, per-pattern logL, (df, ddf) at two lengths,
computeLikelihoodFromBuffer, the Newton-Raphson branch optimum; binary (morphological-style), DNA and a 5-state mixture."""
import numpy as np
import pytest

from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine
from oracle.oracle_engine import OracleEngine
from test_asc_cpu import asc_case, holder_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,safe", [("bin", True), ("dna", False), ("dna", True), ("n5mix", True), ("aa", False)])
def test_asc_lewis(kind, safe):
    tree, model, pat, freq, unobs, pat_all, freq_all = asc_case(kind, ntaxa=16, nsites=600)
    nobs, U = pat.shape[1], unobs.shape[0]
    nsite = int(freq.sum())
    g = LikelihoodEngine(model, tree.ntaxa, nobs + U, safe_numeric=safe)
    o = OracleEngine(model, tree.ntaxa, nobs + U, safe_numeric=safe)
    for e in (g, o):
        e.set_tip_states(pat_all); e.set_pattern_freq(freq_all); e.set_pattern_invar(np.zeros(nobs + U))
        e.set_ascertainment(1, U, nsite)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    np.testing.assert_allclose(g.get_pattern_lh()[:nobs], o.get_pattern_lh()[:nobs], rtol=1e-11, atol=0)
    for x in (L, 0.4 * L + 0.02):
        np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, x), o.branch_derv(ops[:0], a, b, x), rtol=1e-9, atol=1e-7)
        assert abs(g.logl_from_buffer(x) - o.logl_from_buffer(x)) < 1e-6
    # the derivatives are those of the corrected log-likelihood
    h = 1e-5
    df, _ = g.branch_derv(ops[:0], a, b, L)
    fd = (g.logl_from_buffer(L + h) - g.logl_from_buffer(L - h)) / (2 * h)
    assert abs(fd - df) < 1e-4 * max(1.0, abs(df))
    # switching the correction off gives the uncorrected likelihood of the same patterns
    g.set_ascertainment(0, 0, 0)
    ll_plain = g.branch_logl(ops[:0], a, b, L)
    assert abs(ll_plain - ll_g) > 1e-3
    g.close()


@pytest.mark.parametrize("kind,safe", [("bin", True), ("dna", False), ("dna", True), ("aa", False), ("n5mix", True)])
def test_asc_holder_missing_data(kind, safe):
    """+ASC_MIS (ASC_VARIANT_MISSING): per-pattern correction over the nstates x padded-nptn block of constant patterns that
    keep each pattern's gaps; the oracle is pinned against the compiled reference in test_asc_cpu.py"""
    tree, model, pat, freq, unobs, pat_all, freq_all = holder_case(kind, ntaxa=16, nsites=500 if kind != "aa" else 200)
    nobs, U = pat.shape[1], unobs.shape[0]
    g = LikelihoodEngine(model, tree.ntaxa, nobs + U, safe_numeric=safe)
    o = OracleEngine(model, tree.ntaxa, nobs + U, safe_numeric=safe)
    for e in (g, o):
        e.set_tip_states(pat_all); e.set_pattern_freq(freq_all); e.set_pattern_invar(np.zeros(nobs + U))
        e.set_ascertainment(2, U, int(freq.sum()))
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    np.testing.assert_allclose(g.get_pattern_lh()[:nobs], o.get_pattern_lh()[:nobs], rtol=1e-11, atol=0)
    for x in (L, 0.4 * L + 0.02):
        np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, x), o.branch_derv(ops[:0], a, b, x), rtol=1e-9, atol=1e-7)
        assert abs(g.logl_from_buffer(x) - o.logl_from_buffer(x)) < 1e-6
    np.testing.assert_allclose(g.get_pattern_lh()[:nobs], o.get_pattern_lh()[:nobs], rtol=1e-11, atol=0)
    # the derivatives are those of the corrected log-likelihood
    h = 1e-5
    df, _ = g.branch_derv(ops[:0], a, b, L)
    fd = (g.logl_from_buffer(L + h) - g.logl_from_buffer(L - h)) / (2 * h)
    assert abs(fd - df) < 1e-4 * max(1.0, abs(df))
    # Newton-Raphson on the corrected likelihood stops at a stationary point of the oracle's corrected likelihood
    xg, _ = g.newton_branch(ops[:0], a, b, 1e-6, L, 10.0, 1e-6, 100)
    if 1e-5 < xg < 9.0:
        dfx, ddfx = o.branch_derv(ops[:0], a, b, xg)
        assert abs(dfx) <= 1e-5 * abs(ddfx) + 1e-6, (xg, dfx, ddfx)
    # another correction than Lewis' on the same data
    g.set_ascertainment(0, 0, 0)
    assert abs(g.branch_logl(ops[:0], a, b, L) - ll_g) > 1e-3
    # a malformed block is refused
    with pytest.raises(Exception):
        g.set_ascertainment(2, U - 1, 0)
    g.close()
