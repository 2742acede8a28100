"""GPU: UFBoot RELL on the device (b200lk_set_boot_samples / b200lk_rell) against the reference's arithmetic restated in
numpy — IQTree::saveCurrentTree (tree/iqtree.cpp:3551-3637): pattern log-likelihoods cast to float, dotted in float with every
replicate's resampled pattern frequencies (BootValType = float, dotProductSIMD over 8-float lanes).  The device forms the same
float products and accumulates them in double, so it agrees with a float64 sum of float products to float rounding of the
result, and with the reference's 8-lane float accumulation within that accumulation's own error."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine

pytestmark = pytest.mark.gpu


def reference_rell(pattern_lh, boot):
    """dotProductSIMD<float, Vec8f>: 8 lane accumulators over consecutive 8-element groups, horizontal add at the end"""
    lh = pattern_lh.astype(np.float32)
    out = np.zeros(boot.shape[0], dtype=np.float32)
    n8 = lh.size // 8 * 8
    for b in range(boot.shape[0]):
        prod = lh * boot[b]
        lanes = prod[:n8].reshape(-1, 8).sum(axis=0, dtype=np.float32)
        out[b] = np.float32(lanes.sum(dtype=np.float32) + prod[n8:].sum(dtype=np.float32))
    return out


@pytest.mark.parametrize("kind,nsites", [("dna", 3000), ("aa", 901)])
def test_rell_matches_reference_arithmetic(kind, nsites):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=12, nsites=nsites)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g.branch_logl(ops, a, b, L)
    plh = g.get_pattern_lh()
    rng = np.random.default_rng(5)
    nsite_total = int(freq.sum())
    # Alignment::createBootstrapAlignment: multinomial resampling of the sites, expressed as pattern frequencies
    boot = rng.multinomial(nsite_total, freq / freq.sum(), size=37).astype(np.float32)
    g.set_boot_samples(boot)
    rell = g.rell()
    exact = ((plh.astype(np.float32)[None, :] * boot).astype(np.float64)).sum(axis=1)   # float products, exact sum
    np.testing.assert_allclose(rell, exact.astype(np.float32), rtol=2e-7)               # float rounding of the result only
    ref = reference_rell(plh, boot)
    np.testing.assert_allclose(rell, ref, rtol=3e-6)      # the reference's float accumulation error over ~1e3 terms
    # the unresampled alignment reproduces the tree log-likelihood
    g.set_boot_samples(freq.astype(np.float32)[None, :])
    assert abs(float(g.rell()[0]) - float((plh * freq).sum())) < 2e-6 * abs(float((plh * freq).sum()))
    # a new evaluation is picked up without re-uploading the samples
    g.branch_logl(ops[:0], a, b, L * 1.7)
    assert abs(float(g.rell()[0]) - float((g.get_pattern_lh() * freq).sum())) < 2e-6 * abs(float((plh * freq).sum()))
    g.set_boot_samples(np.zeros((0, pat.shape[1]), dtype=np.float32))
    g.close()
