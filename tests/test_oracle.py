"""CPU tests: the oracle against the independent brute-force pruning (Route C), and its scaling paths."""
import numpy as np
import pytest

import bruteforce as bf
from cases import load_engine, make_case
from iqtree2_b200 import models, plan, treesim
from oracle.oracle_engine import OracleEngine


@pytest.mark.parametrize("kind", ["dna", "dna_inv", "aa", "n5mix", "bin", "n7"])
@pytest.mark.parametrize("safe", [False, True])
def test_oracle_matches_bruteforce(kind, safe):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=10, nsites=120)
    e = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    for br in (0, 3, plan.pendant_branch_of(tree, 2)):
        ops, a, b, L = plan.plan_branch(tree, br)
        ll = e.branch_logl(ops, a, b, L)
        ref_ll, ref_df, ref_ddf = bf.tree_derivatives(tree, model, pat, freq, br, ptn_invar=inv)
        assert abs(ll - ref_ll) < 1e-8
        ref_ptn = bf.pattern_loglik(tree, model, pat, br, ptn_invar=inv)
        np.testing.assert_allclose(e.get_pattern_lh(), ref_ptn, rtol=1e-10, atol=1e-10)
        e.theta_reset()
        df, ddf = e.branch_derv(ops[:0], a, b, L)
        assert abs(df - ref_df) < 1e-7 * max(1.0, abs(ref_df))
        assert abs(ddf - ref_ddf) < 1e-7 * max(1.0, abs(ref_ddf))
        assert abs(e.logl_from_buffer(L) - ll) < 1e-9
        # derivative at a different length, theta reused
        df2, _ = e.branch_derv(ops[:0], a, b, L * 1.7)
        _, r_df2, _ = bf.tree_derivatives(tree, model, pat, freq, br, length=L * 1.7, ptn_invar=inv)
        assert abs(df2 - r_df2) < 1e-7 * max(1.0, abs(r_df2))


def test_oracle_logl_is_branch_independent():
    tree, model, pat, freq, inv = make_case("dna", ntaxa=14, nsites=200)
    e = load_engine(OracleEngine, tree, model, pat, freq, inv)
    vals = []
    for br in range(len(tree.edges)):
        ops, a, b, L = plan.plan_branch(tree, br)
        vals.append(e.branch_logl(ops, a, b, L))
    assert np.ptp(vals) < 1e-8


@pytest.mark.parametrize("safe", [False, True])
def test_oracle_scaling_triggers_and_stays_consistent(safe):
    """Deep caterpillar tree with long branches: partials underflow 2^-256 and scale_num becomes non-zero;
    the log-likelihood must still agree with the brute-force value."""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=220, nsites=40, tree_kind="caterpillar", mean_len=0.6, gaps=0.0, ambig=0.0)
    e = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    ll = e.branch_logl(ops, a, b, L)
    assert e.get_scale_num(a[1]).max() >= 1
    ref = bf.tree_loglik(tree, model, pat, freq, br)
    assert abs(ll - ref) < 1e-6
    e.theta_reset()
    e.branch_derv(ops[:0], a, b, L)
    assert abs(e.logl_from_buffer(L) - ll) < 1e-8


def test_tip_partial_lh_matches_numpy():
    for model in (models.make_gtr_gamma(), models.make_lg_gamma(), models.make_lg_c60_gamma(nclasses=3)):
        e = OracleEngine(model, 4, 8)
        np.testing.assert_allclose(e.get_tip_partial_lh(), models.tip_partial_lh(model), rtol=0, atol=1e-15)


def test_save_log_value_false():
    tree, model, pat, freq, inv = make_case("dna", ntaxa=8, nsites=100)
    e = load_engine(OracleEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 0)
    ll = e.branch_logl(ops, a, b, L, save_log_value=True)
    logs = e.get_pattern_lh()
    ll2 = e.branch_logl(ops[:0], a, b, L, save_log_value=False)
    assert abs(ll - ll2) < 1e-10
    np.testing.assert_allclose(np.log(e.get_pattern_lh()) + e.get_pattern_scaling(), logs, rtol=1e-13)


def test_pattern_lh_cat_sums_to_pattern_lh():
    tree, model, pat, freq, inv = make_case("aa", ntaxa=8, nsites=60)
    e = load_engine(OracleEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 1)
    e.branch_logl(ops, a, b, L)
    cat = e.get_pattern_lh_cat()
    sc = e.get_scale_num(a[1])[:, 0].astype(float)
    if b[0] == 1:
        sc = sc + e.get_scale_num(b[1])[:, 0]
    np.testing.assert_allclose(np.log(np.abs(cat.sum(axis=1)) + inv) + sc * (-177.4456782233459932741), e.get_pattern_lh(), rtol=1e-12)
