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


def test_rell_against_the_reference_kernel():
    """pinned: the reference tree's own dotProduct member (PhyloTree::dotProductSIMD<float, Vec8f>, bound by setDotProductFMA;
    oracle/ref_harness.cpp iqref_rell restates only the float cast and zero padding of IQTree::saveCurrentTree around it)"""
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libiqref.so not built")
    tree, model, pat, freq, inv = make_case("dna", ntaxa=14, nsites=5000, seed=3, ambig=0.0)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    r = ref_engine.RefTree(tree, model, pat, freq)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g.branch_logl(ops, a, b, L)
    _, plh_ref = r.branch_logl(0)
    np.testing.assert_allclose(g.get_pattern_lh(), plh_ref, rtol=1e-11)
    rng = np.random.default_rng(8)
    boot = rng.multinomial(int(freq.sum()), freq / freq.sum(), size=100).astype(np.float32)
    g.set_boot_samples(boot)
    np.testing.assert_allclose(g.rell(), r.rell(plh_ref, boot), rtol=3e-6)   # the reference's 8-lane float accumulation error
    g.close(); r.close()


def reference_bookkeeping(state, rell, cur_logl, eps, uniform):
    """tree/iqtree.cpp:3613-3628, statement for statement (random_double(rstream) replaced by the caller's numbers)"""
    boot_logl, boot_counts, boot_orig = state
    updated = np.zeros(len(rell), dtype=np.uint8)
    for s in range(len(rell)):
        r = float(rell[s])
        better = r > boot_logl[s] + eps
        if not better and r > boot_logl[s] - eps:
            better = uniform[s] <= 1.0 / (boot_counts[s] + 1)
        if better:
            if r <= boot_logl[s] + eps:
                boot_counts[s] += 1
            else:
                boot_counts[s] = 1
            boot_logl[s] = max(boot_logl[s], r)
            boot_orig[s] = cur_logl
            updated[s] = 1
    return updated


def test_ufboot_bookkeeping_on_device():
    """a sequence of candidate trees (branch-length variants of one topology, close enough in logL for ties within
    ufboot_epsilon): boot_logl / boot_counts / boot_orig_logl and the replicates whose tree changes follow the reference's rule"""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=12, nsites=2500, seed=6)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    rng = np.random.default_rng(2)
    nboot = 64
    boot = rng.multinomial(int(freq.sum()), freq / freq.sum(), size=nboot).astype(np.float32)
    g.set_boot_samples(boot)
    state = [np.full(nboot, -np.finfo(np.float64).max), np.zeros(nboot, dtype=np.int64), np.full(nboot, -np.finfo(np.float64).max)]
    eps = 0.5
    ll0 = g.branch_logl(ops, a, b, L)
    for it, scale in enumerate([1.0, 1.002, 0.999, 1.3, 1.001, 0.7, 1.0005, 1.0]):
        ll = g.branch_logl(ops[:0], a, b, L * scale)
        rell = g.rell()
        u = rng.random(nboot)
        want = reference_bookkeeping(state, rell, ll, eps, u)
        upd, n = g.rell_update(ll, eps, u)
        np.testing.assert_array_equal(upd, want)
        assert n == int(want.sum())
        bl, bc, bo = g.get_boot_state()
        np.testing.assert_array_equal(bl, state[0]); np.testing.assert_array_equal(bc, state[1]); np.testing.assert_array_equal(bo, state[2])
    assert state[1].max() >= 2, "the sequence is expected to produce ties"
    g.boot_reset()
    bl, bc, bo = g.get_boot_state()
    assert np.all(bc == 0) and np.all(bl == -np.finfo(np.float64).max)
    g.close()
