"""GPU: mixed branch lengths (heterotachy; b200lk_set_mixlen, iqtree2_b200/csrc/b200lk_mixlen.cuh + resolve_len_cat in every info
kernel) through the C ABI against the CPU restatement (oracle/lk_oracle_mixlen.h, pinned to the compiled reference run as a mixlen
tree in tests/test_mixlen_cpu.py): CLVs (scale_num bit-exact), _pattern_lh, logL, the from-buffer form, the gradient / Hessian form of
computeLikelihoodDerv, computeLikelihoodDervMixlen for every class — on every tuned walker (4 states x any categories, 20 states,
general n) — and against the live reference."""
import numpy as np
import pytest

from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine
from oracle.oracle_engine import OracleEngine
from test_gpu_parity import assert_partials_close
from test_mixlen_cpu import make_mixlen_case, mixlen_engine, mixlen_table, plan_rows

pytestmark = pytest.mark.gpu


def compare_mixlen(tree, model, pat, freq, invar, safe, branches, clv_tol=1e-11):
    g = mixlen_engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
    o = mixlen_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for br in branches:
        ops, a, b, row = plan_rows(tree, br)
        ll_g, ll_o = g.branch_logl(ops, a, b, row), o.branch_logl(ops, a, b, row)
        assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o), (br, ll_g, ll_o)
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), f"scale_num differs in slot {s}"
            assert_partials_close(g.get_partial(s), o.get_partial(s), clv_tol)
        assert_partials_close(g.get_pattern_lh_cat(), o.get_pattern_lh_cat())
        gg, Hg, lg = g.branch_derv_mixlen(ops[:0], a, b, row)
        go, Ho, lo = o.branch_derv_mixlen(ops[:0], a, b, row)
        np.testing.assert_allclose(gg, go, rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(Hg, Ho, rtol=1e-9, atol=1e-6)
        assert abs(lg - lo) < 1e-9 * abs(lo)
        for cur in range(model.ncat):
            np.testing.assert_allclose(g.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), o.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), rtol=1e-9, atol=1e-7)
        other = (int(row) + 1) % len(tree.lengths)   # the resident theta buffer at another row's lengths
        assert abs(g.logl_from_buffer(other) - o.logl_from_buffer(other)) < 1e-9 * abs(ll_o)
        g.theta_reset(); o.theta_reset()
    g.close()


@pytest.mark.parametrize("kind,ncat", [("dna", 4), ("dna", 2), ("dna", 3), ("aa", 4), ("aa", 2), ("n5", 3), ("n12", 2), ("aa_mix", 2)])
@pytest.mark.parametrize("safe", [False, True])
def test_mixlen_parity(kind, ncat, safe):
    if kind == "aa_mix" and not safe:
        pytest.skip("mixtures run the SAFE kernel")
    tree, model, pat, freq, invar = make_mixlen_case(kind, ncat, ntaxa=13, nsites=400, seed=17)
    nb = len(tree.lengths)
    compare_mixlen(tree, model, pat, freq, invar, safe, [0, nb // 3, nb - 1, plan.pendant_branch_of(tree, 4)])


def test_mixlen_with_invariant_sites_and_a_new_table():
    """lengths change (an optimiser step): the table is set again, stale CLVs are recomputed, results follow the oracle's"""
    tree, model, pat, freq, invar = make_mixlen_case("dna", 4, ntaxa=10, nsites=500, seed=23, pinvar=0.15)
    assert (invar > 0).any()
    g = mixlen_engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    o = mixlen_engine(OracleEngine, tree, model, pat, freq, invar, False)
    ops, a, b, row = plan_rows(tree, 2)
    tab = mixlen_table(tree, model)
    for scale in (1.0, 0.7, 1.9):
        for e in (g, o):
            e.set_mixlen(tab * scale)
            e.theta_reset()
        ll_g, ll_o = g.branch_logl(ops, a, b, row), o.branch_logl(ops, a, b, row)
        assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o)
        gg, Hg, _ = g.branch_derv_mixlen(ops[:0], a, b, row)
        go, Ho, _ = o.branch_derv_mixlen(ops[:0], a, b, row)
        np.testing.assert_allclose(gg, go, rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(Hg, Ho, rtol=1e-9, atol=1e-6)
    g.close()


def test_mixlen_refusals_and_switching_off():
    tree, model, pat, freq, invar = make_mixlen_case("dna", 4, ntaxa=8, nsites=200, seed=29)
    g = mixlen_engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    ops, a, b, row = plan_rows(tree, 1)
    g.branch_logl(ops, a, b, row)
    with pytest.raises(Exception):
        g.branch_derv(ops[:0], a, b, row)            # the scalar derivative has no meaning on a mixlen tree
    with pytest.raises(Exception):
        g.branch_logl(ops, a, b, 0.37)               # not a row
    with pytest.raises(Exception):
        g.branch_logl(ops, a, b, float(len(tree.lengths)))   # past the table
    # off again: plain lengths, same answer as an engine that never had a table
    g.set_mixlen(None)
    g.theta_reset()
    ops2, a2, b2, L2 = plan.plan_branch(tree, 1)
    h = LikelihoodEngine(model, tree.ntaxa, pat.shape[1])
    h.set_tip_states(pat); h.set_pattern_freq(freq); h.set_pattern_invar(invar)
    assert g.branch_logl(ops2, a2, b2, L2) == h.branch_logl(ops2, a2, b2, L2)
    np.testing.assert_allclose(g.branch_derv(ops2[:0], a2, b2, L2), h.branch_derv(ops2[:0], a2, b2, L2), rtol=1e-12)
    g.close(); h.close()


def test_mixlen_against_the_live_reference():
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libiqref.so not built")
    for kind, ncat, safe in (("dna", 4, False), ("aa", 4, True), ("aa", 3, False)):
        tree, model, pat, freq, invar = make_mixlen_case(kind, ncat, ntaxa=30, nsites=1500, seed=31)
        r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=4)
        g = mixlen_engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
        for leaf in (0, 11, 25):
            ops, a, b, row = plan_rows(tree, plan.pendant_branch_of(tree, leaf))
            ll_r, plh_r = r.branch_logl(leaf)
            ll_g = g.branch_logl(ops, a, b, row)
            assert abs(ll_r - ll_g) < 1e-9 * abs(ll_r)
            np.testing.assert_allclose(g.get_pattern_lh(), plh_r, rtol=1e-11, atol=0)
            g_r, H_r, l_r = r.branch_derv_mixlen(leaf)
            g_g, H_g, l_g = g.branch_derv_mixlen(ops[:0], a, b, row)
            np.testing.assert_allclose(g_g, g_r, rtol=1e-9, atol=1e-7)
            np.testing.assert_allclose(H_g, H_r, rtol=1e-9, atol=1e-6)
            assert abs(l_g - l_r) < 1e-9 * abs(l_r)
            for cur in range(model.ncat):
                np.testing.assert_allclose(g.branch_derv_mixlen_cur(ops[:0], a, b, row, cur), r.branch_derv_mixlen_cur(leaf, cur), rtol=1e-9, atol=1e-7)
            g.theta_reset()
        g.close(); r.close()
