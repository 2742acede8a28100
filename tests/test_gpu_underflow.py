"""GPU: the reference's repair of a non-finite tree log-likelihood (tree/phylokernelnew.h:3107-3132, 3371-3383): every pattern
whose log-likelihood is not finite gets LOG_SCALING_THRESHOLD*4, the tree logL is re-summed, and the call reports the
condition (the reference prints a warning; the C ABI returns B200LK_WARN_UNDERFLOW).  Forced here through ptn_invar (the only
per-pattern input that reaches log() unscaled); the CPU oracle restates the reference's block statement for statement."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine
from oracle.oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu
LOG_SCALING_THRESHOLD = -177.4456782233459932741
WARN_UNDERFLOW = 1


@pytest.mark.parametrize("kind,safe", [("dna", False), ("dna", True), ("aa", False), ("dna_g8", False)])
@pytest.mark.parametrize("save_log", [True, False])
def test_non_finite_logl_is_repaired_pattern_by_pattern(kind, safe, save_log):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=13, nsites=400, seed=11)
    bad = np.array([3, 17, pat.shape[1] - 1])
    inv = inv.copy()
    inv[bad[0]] = np.nan          # log(NaN)
    inv[bad[1]] = np.inf          # log(inf)
    inv[bad[2]] = -np.nan
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll_g = g.branch_logl(ops, a, b, L, save_log_value=save_log)
    assert g.last_status == WARN_UNDERFLOW
    ll_o = o.branch_logl(ops, a, b, L, save_log_value=save_log)
    assert o.last_status == WARN_UNDERFLOW
    assert np.isfinite(ll_g) and abs(ll_g - ll_o) < 1e-6 + 1e-12 * abs(ll_o), (ll_g, ll_o)
    plh_g, plh_o = g.get_pattern_lh(), o.get_pattern_lh()
    np.testing.assert_allclose(plh_g, plh_o, rtol=1e-11, atol=0)
    if save_log:
        assert np.all(plh_g[bad] == LOG_SCALING_THRESHOLD * 4)
    else:
        # save_log_value = false: _pattern_lh holds the raw likelihood, so only NaN/inf likelihoods count as non-finite there
        assert np.all(plh_g[bad] == 1.0)
        np.testing.assert_array_equal(g.get_pattern_scaling()[bad], np.full(3, LOG_SCALING_THRESHOLD * 4))
        np.testing.assert_allclose(g.get_pattern_scaling(), o.get_pattern_scaling(), rtol=0, atol=0)
    # the repaired values are what the callers of computePatternLikelihood see afterwards; a healthy evaluation clears the status
    g.set_pattern_invar(np.where(np.isfinite(inv), inv, 0.0)); o.set_pattern_invar(np.where(np.isfinite(inv), inv, 0.0))
    ll2 = g.branch_logl(ops[:0], a, b, L)
    assert g.last_status == 0 and abs(ll2 - o.branch_logl(ops[:0], a, b, L)) < 1e-6
    g.close()


def test_from_buffer_repair():
    """computeLikelihoodFromBuffer's repair (tree/phylokernelnew.h:3371-3383)"""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=11, nsites=300, seed=5)
    inv = inv.copy(); inv[7] = np.nan
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g.branch_derv(ops, a, b, L); o.branch_derv(ops, a, b, L)
    ll_g, ll_o = g.logl_from_buffer(L * 1.3), o.logl_from_buffer(L * 1.3)
    assert g.last_status == WARN_UNDERFLOW
    assert abs(ll_g - ll_o) < 1e-6 + 1e-12 * abs(ll_o)
    np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11)
    assert g.get_pattern_lh()[7] == LOG_SCALING_THRESHOLD * 4
    g.close()
