"""GPU: the E-step of the reference's EM algorithms on the device (b200lk_em_posterior: computePatternLhCat(wsl) + the loops of
RateFree::optimizeWithEM / ModelMixture::optimizeWithEM) against the CPU oracle, which tests/test_em_cpu.py pins against the
compiled reference; and the M-step hand-over (one posterior column becomes the ptn_freq of another engine, device to device)."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import models, plan
from iqtree2_b200.engine import EngineError, LikelihoodEngine
from oracle.oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu

WSL_RATECAT, WSL_MIXTURE, WSL_MIXTURE_RATECAT = 2, 3, 4


@pytest.mark.parametrize("kind,safe", [("dna", False), ("dna_inv", False), ("aa", False), ("aa_c4", False), ("aa_c4", True),
                                       ("aa_fused", False), ("n5mix", True), ("aa_c60", False)])
def test_em_posterior(kind, safe):
    kw = dict(ntaxa=8, nsites=60) if kind == "aa_c60" else dict(ntaxa=14, nsites=700)
    tree, model, pat, freq, inv = make_case(kind, **kw)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    with pytest.raises(EngineError):
        g.em_posterior(WSL_RATECAT, True)          # nothing evaluated yet
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    assert abs(g.branch_logl(ops, a, b, L) - o.branch_logl(ops, a, b, L)) < 1e-6
    for wsl in (WSL_RATECAT, WSL_MIXTURE, WSL_MIXTURE_RATECAT):
        if wsl == WSL_MIXTURE and model.nmix == 1:
            continue
        for add_invar in (True, False):
            prop_o, post_o = o.em_posterior(wsl, add_invar)
            prop_g, post_g = g.em_posterior(wsl, add_invar, want_posterior=True)
            assert post_g.shape == post_o.shape
            np.testing.assert_allclose(post_g, post_o, rtol=1e-9, atol=1e-13 * freq.max())
            np.testing.assert_allclose(prop_g, prop_o, rtol=1e-10)
            np.testing.assert_array_equal(g.em_posterior(wsl, add_invar), prop_g)      # deterministic reduction
    g.close()


def test_posterior_column_becomes_pattern_freq():
    """M-step of RateFree::optimizeWithEM (model/ratefree.cpp:617-650): a tree with one rate category whose ptn_freq are the
    posteriors of category c — built without the posteriors leaving the device"""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=14, nsites=900)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g.branch_logl(ops, a, b, L)
    prop, post = g.em_posterior(WSL_RATECAT, True, want_posterior=True)
    c = 2
    one = models.make_gtr_gamma(ncat=1)
    m1 = LikelihoodEngine(one, tree.ntaxa, pat.shape[1])
    m1.set_tip_states(pat); m1.set_pattern_invar(np.zeros(pat.shape[1]))
    m1.set_pattern_freq_from_posterior(g, c)
    ll_dev = m1.branch_logl(ops, a, b, L)
    m2 = LikelihoodEngine(one, tree.ntaxa, pat.shape[1])
    m2.set_tip_states(pat); m2.set_pattern_invar(np.zeros(pat.shape[1])); m2.set_pattern_freq(np.ascontiguousarray(post[:, c]))
    assert abs(ll_dev - m2.branch_logl(ops, a, b, L)) < 1e-9 * abs(ll_dev)
    np.testing.assert_array_equal(m1.get_pattern_lh(), m2.get_pattern_lh())
    with pytest.raises(EngineError):
        m1.set_pattern_freq_from_posterior(g, 99)
    for e in (g, m1, m2):
        e.close()
