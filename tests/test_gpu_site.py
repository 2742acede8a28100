"""GPU: site-specific models (ModelSet / PMSF; iqtree2_b200/csrc/b200lk_site.cuh, the SITE_MODEL template branch of
tree/phylokernelnew.h) through the C ABI against the CPU restatement (oracle/lk_oracle_site.h, itself pinned to the compiled reference's
SITE_MODEL instantiations and to a per-pattern brute-force pruning in tests/test_site_model_cpu.py): every CLV of the traversal
(scale_num bit-exact), _pattern_lh, _pattern_lh_cat, theta, logL, (df, ddf), the from-buffer form, Newton; and the live reference."""
import numpy as np
import pytest

from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine
from oracle.oracle_engine import OracleEngine
from test_gpu_parity import assert_partials_close, newton_ref
from test_site_model_cpu import engine, make_site_case

pytestmark = pytest.mark.gpu


def compare_site(tree, model, pat, freq, invar, safe, branches, clv_tol=1e-11, **kw):
    g = engine(LikelihoodEngine, tree, model, pat, freq, invar, safe, **kw)
    o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for br in branches:
        ops, a, b, L = plan.plan_branch(tree, br)
        for save_log in (True, False):
            ll_g, ll_o = g.branch_logl(ops, a, b, L, save_log_value=save_log), o.branch_logl(ops, a, b, L, save_log_value=save_log)
            assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o), (br, ll_g, ll_o)
            np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), f"scale_num differs in slot {s}"
            assert_partials_close(g.get_partial(s), o.get_partial(s), clv_tol)
        assert_partials_close(g.get_pattern_lh_cat(), o.get_pattern_lh_cat())
        for x in (L, 0.37 * L):
            np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, x), o.branch_derv(ops[:0], a, b, x), rtol=1e-9, atol=1e-7)
        assert_partials_close(g.get_theta(), o.get_theta())
        for x in (L, 2.1 * L):
            lg, lo = g.logl_from_buffer(x), o.logl_from_buffer(x)
            assert abs(lg - lo) < 1e-9 * abs(lo)
        g.theta_reset(); o.theta_reset()
    g.close()


@pytest.mark.parametrize("nstates,ncat", [(4, 4), (4, 1), (5, 3), (12, 2), (20, 4), (20, 1), (61, 2)])
@pytest.mark.parametrize("safe", [False, True])
def test_site_parity(nstates, ncat, safe):
    """4 states (plain warp-interleaved slots), 12 / 20 / 61 states (the fragment-major slot layouts of the FP64-MMA engines, addressed
    through ClvIdx), an odd state count; a pattern count that is not a multiple of the tile; unknown and ambiguous tip states"""
    nptn = 333 if nstates <= 20 else 150
    tree, model, pat, freq, invar = make_site_case(nstates, ncat, ntaxa=13, nptn=nptn, seed=17)
    nb = len(tree.edges)
    compare_site(tree, model, pat, freq, invar, safe, [0, nb // 3, nb - 1, plan.pendant_branch_of(tree, 4)])


@pytest.mark.parametrize("safe", [False, True])
def test_site_parity_scaling_deep_tree(safe):
    tree, model, pat, freq, invar = make_site_case(4, 4, ntaxa=200, nptn=150, seed=7, tree_kind="caterpillar", mean_len=0.8, p_unknown=0.0)
    g = engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
    o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for br in (plan.pendant_branch_of(tree, 0), plan.pendant_branch_of(tree, 80)):
        ops, a, b, L = plan.plan_branch(tree, br)
        ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
        assert max(o.get_scale_num(int(op["dst_slot"])).max() for op in ops) >= 1, "case does not rescale"
        assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s))
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
        np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, L), o.branch_derv(ops[:0], a, b, L), rtol=1e-9, atol=1e-7)
    g.close()


@pytest.mark.parametrize("safe", [False, True])
@pytest.mark.parametrize("ncat", [1, 3, 4, 6])
def test_site_20_states_both_kernels(safe, ncat):
    """the FP64-MMA walker with the matrices resident in shared memory (k_site_walk20: 4-category chunks, so 1 / 3 / 6 categories
    exercise partial and multiple chunks) and the thread-per-pattern kernel (kernel variant 5) against the restatement and each other"""
    tree, model, pat, freq, invar = make_site_case(20, ncat, ntaxa=14, nptn=210, seed=53, pinvar=0.1 if ncat == 3 else 0.0)
    o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
    ops, a, b, L = plan.plan_branch(tree, 5)
    ll_o = o.branch_logl(ops, a, b, L)
    for variant in (0, 5):
        g = engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
        if variant:
            g.set_schedule(10 + variant)
        ll_g = g.branch_logl(ops, a, b, L)
        assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o), (variant, ll_g, ll_o)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s))
            assert_partials_close(g.get_partial(s), o.get_partial(s), 1e-11)
        g.close()


def test_site_20_states_deep_tree_rescales_on_the_mma_kernel():
    tree, model, pat, freq, invar = make_site_case(20, 4, ntaxa=120, nptn=96, seed=7, tree_kind="caterpillar", mean_len=0.8, p_unknown=0.0)
    for safe in (False, True):
        g = engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
        o = engine(OracleEngine, tree, model, pat, freq, invar, safe)
        ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
        ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
        assert max(o.get_scale_num(int(op["dst_slot"])).max() for op in ops) >= 1, "case does not rescale"
        assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s))
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
        g.close()


def test_site_with_invariant_sites():
    tree, model, pat, freq, invar = make_site_case(4, 4, ntaxa=10, nptn=400, seed=23, pinvar=0.15)
    assert (invar > 0).any()
    compare_site(tree, model, pat, freq, invar, False, [1, 5])


def test_site_incremental_equals_full():
    """resident CLVs: one branch length changes, only the op that depends on it is recomputed (cf. test_gpu_parity)"""
    tree, model, pat, freq, invar = make_site_case(20, 4, ntaxa=16, nptn=300, seed=41)
    g = engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g.branch_logl(ops, a, b, L)
    new_ops = ops.copy()
    new_ops[-1]["left"]["length"] *= 2.5
    ll_inc = g.branch_logl(new_ops[-1:], a, b, L)
    g2 = engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    ll_full = g2.branch_logl(new_ops, a, b, L)
    assert ll_inc == ll_full
    g.close(); g2.close()


def test_site_newton_branch_and_unsupported_calls():
    tree, model, pat, freq, invar = make_site_case(4, 4, ntaxa=12, nptn=500, seed=29)
    g = engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    o = engine(OracleEngine, tree, model, pat, freq, invar, False)
    for br in (2, plan.pendant_branch_of(tree, 3)):
        ops, a, b, L = plan.plan_branch(tree, br)
        o.branch_logl(ops, a, b, L)
        o.theta_reset()
        x_g, steps = g.newton_branch(ops, a, b, 1e-6, L, 10.0, 1e-6)
        x_o = newton_ref(lambda x: o.branch_derv(ops[:0], a, b, x), 1e-6, L, 10.0, 1e-6, 100)
        assert abs(x_g - x_o) < 1e-6, (x_g, x_o)
        assert steps >= 1
    with pytest.raises(Exception):
        g.set_ascertainment(1, 4, 500)
    g.close()


def test_site_on_a_multi_device_engine_slices_the_eigensystems():
    """cfg.device = -N: every shard takes its pattern slice of the per-pattern eigensystems (one visible GPU: two shards on it)"""
    tree, model, pat, freq, invar = make_site_case(4, 4, ntaxa=11, nptn=777, seed=3)
    g1 = engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    try:
        g2 = engine(LikelihoodEngine, tree, model, pat, freq, invar, False, device=-2)
    except Exception as ex:   # fewer than two devices and no same-device sharding
        g1.close()
        pytest.skip(str(ex))
    ops, a, b, L = plan.plan_branch(tree, 3)
    l1, l2 = g1.branch_logl(ops, a, b, L), g2.branch_logl(ops, a, b, L)
    assert abs(l1 - l2) < 1e-10 * abs(l1)
    np.testing.assert_allclose(g2.branch_derv(ops[:0], a, b, L), g1.branch_derv(ops[:0], a, b, L), rtol=1e-10)
    g1.close(); g2.close()


def test_site_against_the_live_reference():
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libiqref.so not built")
    for nstates, ncat, safe in ((4, 4, False), (20, 4, True), (20, 1, False)):
        tree, model, pat, freq, invar = make_site_case(nstates, ncat, ntaxa=30, nptn=1200, seed=31)
        r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=4)
        g = engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
        for leaf in (0, 11, 25):
            ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, leaf))
            ll_r, plh_r = r.branch_logl(leaf)
            ll_g = g.branch_logl(ops, a, b, L)
            assert abs(ll_r - ll_g) < 1e-9 * abs(ll_r)
            np.testing.assert_allclose(g.get_pattern_lh(), plh_r, rtol=1e-11, atol=0)
            np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, L), r.branch_derv(leaf), rtol=1e-9)
            assert abs(g.logl_from_buffer(L) - r.logl_from_buffer()) < 1e-9 * abs(ll_r)
            g.theta_reset()
        g.close(); r.close()
