"""GPU (needs >= 2 devices; skipped on a single-GPU box): the multi-device engine behind the same C ABI (cfg.device = -N,
iqtree2_b200/csrc/b200lk_multi.inc): one engine per device over pattern slices, scalars summed in device order.  Everything the
single engine returns must come back unchanged — per-pattern values bit for bit, sums to summation-order noise."""
import os

import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("kind,safe", [("dna", False), ("aa", False), ("dna_g8", True), ("aa_c4", False)])
def test_multi_device_engine_equals_single_engine(kind, safe):
    if _ndev() < 2:
        pytest.skip("needs at least two GPUs")
    tree, model, pat, freq, inv = make_case(kind, ntaxa=14, nsites=1500, seed=21)
    assert pat.shape[1] > 256
    one = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe, device=0)
    two = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe, device=-2)
    ops, a, b, L = plan.plan_branch(tree, 4)
    l1, l2 = one.branch_logl(ops, a, b, L), two.branch_logl(ops, a, b, L)
    assert abs(l1 - l2) < 1e-9 * abs(l1)
    np.testing.assert_array_equal(one.get_pattern_lh(), two.get_pattern_lh())
    np.testing.assert_array_equal(one.get_pattern_lh_cat(), two.get_pattern_lh_cat())
    for op in ops[-3:]:
        s = int(op["dst_slot"])
        np.testing.assert_array_equal(one.get_scale_num(s), two.get_scale_num(s))
        np.testing.assert_array_equal(one.get_partial(s), two.get_partial(s))
    one.theta_reset(); two.theta_reset()
    d1, d2 = one.branch_derv(ops[:0], a, b, L), two.branch_derv(ops[:0], a, b, L)
    np.testing.assert_allclose(d1, d2, rtol=1e-11)
    np.testing.assert_array_equal(one.get_theta(), two.get_theta())
    assert abs(one.logl_from_buffer(0.7 * L) - two.logl_from_buffer(0.7 * L)) < 1e-9 * abs(l1)
    # Newton-Raphson: the shards walk the same iterates (one reduction per step)
    one.theta_reset(); two.theta_reset()
    x1, _ = one.newton_branch(ops[:0], a, b, 1e-6, L, 10.0, 1e-6)
    x2, _ = two.newton_branch(ops[:0], a, b, 1e-6, L, 10.0, 1e-6)
    assert abs(x1 - x2) < 1e-6
    one.close(); two.close()


def test_reference_phylotree_on_two_gpus():
    """the UNMODIFIED reference PhyloTree (oracle/ref_b200_shim.cpp binding) with its kernel pointers on a two-device engine:
    optimizeAllBranches gives the lengths and logL of the same reference over its own AVX kernel"""
    if _ndev() < 2:
        pytest.skip("needs at least two GPUs")
    from oracle import ref_engine
    if not (ref_engine.available() and ref_engine.available_b200()):
        pytest.skip("oracle/_ref libraries not built")
    tree, model, pat, freq, inv = make_case("dna", ntaxa=24, nsites=3000, seed=9, gaps=0.02, ambig=0.0)
    cpu = ref_engine.RefTree(tree, model, pat, freq, nthreads=4)
    os.environ["B200LK_NDEV"] = "2"
    try:
        gpu = ref_engine.RefTree(tree, model, pat, freq, b200=True)
    finally:
        del os.environ["B200LK_NDEV"]
    l_cpu, l_gpu = cpu.optimize_all_branches(5, 0.001), gpu.optimize_all_branches(5, 0.001)
    assert abs(l_cpu - l_gpu) < 1e-6 * max(1.0, abs(l_cpu) * 1e-6) + 1e-6
    bc, bg = cpu.branches(), gpu.branches()
    for k in bc:
        assert abs(bc[k] - bg[k]) < 1e-6
    cpu.close(); gpu.close()


def test_multi_device_mixlen_and_site_models():
    """mixed branch lengths (every shard takes the table; gradient / Hessian summed in device order) and site-specific models (every
    shard takes its pattern slice of the eigensystems) on a two-device engine"""
    if _ndev() < 2:
        pytest.skip("needs at least two GPUs")
    from test_mixlen_cpu import make_mixlen_case, mixlen_engine, plan_rows
    tree, model, pat, freq, invar = make_mixlen_case("dna", 4, ntaxa=12, nsites=1500, seed=61)
    one = mixlen_engine(LikelihoodEngine, tree, model, pat, freq, invar, False, device=0)
    two = mixlen_engine(LikelihoodEngine, tree, model, pat, freq, invar, False, device=-2)
    ops, a, b, row = plan_rows(tree, 3)
    l1, l2 = one.branch_logl(ops, a, b, row), two.branch_logl(ops, a, b, row)
    assert abs(l1 - l2) < 1e-9 * abs(l1)
    np.testing.assert_array_equal(one.get_pattern_lh(), two.get_pattern_lh())
    g1, H1, ll1 = one.branch_derv_mixlen(ops[:0], a, b, row)
    g2, H2, ll2 = two.branch_derv_mixlen(ops[:0], a, b, row)
    np.testing.assert_allclose(g2, g1, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(H2, H1, rtol=1e-10, atol=1e-8)
    assert abs(ll1 - ll2) < 1e-9 * abs(ll1)
    for cur in range(model.ncat):
        np.testing.assert_allclose(two.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), one.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), rtol=1e-10, atol=1e-9)
    one.close(); two.close()

    from test_site_model_cpu import engine, make_site_case
    tree, model, pat, freq, invar = make_site_case(20, 4, ntaxa=11, nptn=777, seed=3)
    one = engine(LikelihoodEngine, tree, model, pat, freq, invar, False, device=0)
    two = engine(LikelihoodEngine, tree, model, pat, freq, invar, False, device=-2)
    ops, a, b, L = plan.plan_branch(tree, 3)
    l1, l2 = one.branch_logl(ops, a, b, L), two.branch_logl(ops, a, b, L)
    assert abs(l1 - l2) < 1e-10 * abs(l1)
    np.testing.assert_array_equal(one.get_pattern_lh(), two.get_pattern_lh())
    np.testing.assert_allclose(two.branch_derv(ops[:0], a, b, L), one.branch_derv(ops[:0], a, b, L), rtol=1e-10)
    one.close(); two.close()
