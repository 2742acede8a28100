"""CPU: PhyloTree::computePatternLhCat(wsl) (tree/phylotree.cpp:1338-1393; WSL_RATECAT / WSL_MIXTURE / WSL_MIXTURE_RATECAT) in the
oracle against the compiled reference, and the E-step of RateFree::optimizeWithEM (model/ratefree.cpp:556-573) and
ModelMixture::optimizeWithEM (model/modelmixture.cpp:2008-2024) — two short loops over those categories, restated in the
oracle — against the same loops written with numpy on the reference's own _pattern_lh_cat."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from oracle import ref_engine
from oracle.oracle_engine import OracleEngine

WSL_RATECAT, WSL_MIXTURE, WSL_MIXTURE_RATECAT = 2, 3, 4


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind,safe", [("dna", False), ("dna_inv", False), ("aa_c4", False), ("aa_c4", True), ("aa_fused", False),
                                       ("n5mix", True)])
def test_pattern_lh_cat_wsl_and_em_step(kind, safe):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=12, nsites=300)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    o.branch_logl(ops, a, b, L)
    for wsl in (WSL_RATECAT, WSL_MIXTURE, WSL_MIXTURE_RATECAT):
        if wsl == WSL_MIXTURE and model.nmix == 1:
            continue          # getNumLhCat(WSL_MIXTURE) = 1 for a non-mixture: not a meaningful request
        ref = r.pattern_lh_cat_wsl(wsl)
        got = o.pattern_lh_cat_wsl(wsl)
        assert got.shape == ref.shape
        scale = np.abs(ref).max(axis=1, keepdims=True)
        np.testing.assert_allclose(got / scale, ref / scale, rtol=0, atol=1e-13)
        for add_invar in (True, False):
            lk = (inv if add_invar else 0.0) + ref.sum(axis=1)
            post_ref = ref * (freq / lk)[:, None]
            prop, post = o.em_posterior(wsl, add_invar)
            np.testing.assert_allclose(post, post_ref, rtol=1e-11, atol=1e-300)
            np.testing.assert_allclose(prop, post_ref.sum(axis=0), rtol=1e-11)
            if not add_invar or not inv.any():
                assert abs(prop.sum() - freq.sum()) < 1e-8 * freq.sum()      # posteriors of a pattern add up to its frequency
    r.close()
