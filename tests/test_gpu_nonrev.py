"""GPU: the non-reversible kernels (iqtree2_b200/csrc/b200lk_nonrev.cuh; tree/phylokernelnonrev.h) through the C ABI against the CPU
restatement (oracle/lk_oracle_nonrev.h, itself pinned to the compiled reference and to a brute-force pruning in
tests/test_nonrev_cpu.py) on rooted trees: every CLV of the traversal (scale_num bit-exact), _pattern_lh, logL, (df, ddf), on
root-ward, leaf-ward, pendant and root branches; and against the live compiled reference."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine
from nonrev_cases import make_nonrev_case
from oracle.oracle_engine import OracleEngine
from test_gpu_parity import assert_partials_close, newton_ref
from test_nonrev_cpu import nonrev_engine

pytestmark = pytest.mark.gpu


def compare_nonrev(tree, model, pat, freq, invar, safe, branches, clv_tol=1e-11):
    g = nonrev_engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
    o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    root_br = plan.pendant_branch_of(tree, tree.ntaxa - 1)
    for br in branches:
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
        assert abs(ll_g - ll_o) < 1e-6, (br, ll_g, ll_o)
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), f"scale_num differs in slot {s}"
            assert_partials_close(g.get_partial(s), o.get_partial(s), clv_tol)
        assert_partials_close(g.get_pattern_lh_cat(), o.get_pattern_lh_cat())
        if br != root_br:
            for x in (L, 0.37 * L):
                np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, x), o.branch_derv(ops[:0], a, b, x), rtol=1e-9, atol=1e-7)
    g.close()


@pytest.mark.parametrize("nstates,ncat,nmix", [(4, 4, 1), (4, 1, 1), (5, 3, 2), (12, 2, 1), (20, 4, 1), (20, 2, 2)])
@pytest.mark.parametrize("safe", [False, True])
def test_nonrev_parity(nstates, ncat, nmix, safe):
    """4 states (plain warp-interleaved slots), 12 and 20 states (the fragment-major slot layouts of the FP64-MMA engines, addressed
    through ClvIdx), odd state counts, mixtures; a pattern count that is not a multiple of the tile"""
    tree, model, pat, freq, invar = make_nonrev_case(nstates, ncat, nmix, ntaxa=13, nptn=333, seed=17)
    nb = len(tree.edges)
    compare_nonrev(tree, model, pat, freq, invar, safe, [0, nb // 3, nb - 1, plan.pendant_branch_of(tree, 4), plan.pendant_branch_of(tree, tree.ntaxa - 1)])


@pytest.mark.parametrize("safe", [False, True])
def test_nonrev_parity_scaling_deep_tree(safe):
    tree, model, pat, freq, invar = make_nonrev_case(4, 4, 1, ntaxa=200, nptn=150, seed=7, tree_kind="caterpillar", mean_len=0.8, p_unknown=0.0)
    g = nonrev_engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
    o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for br in (plan.pendant_branch_of(tree, 0), len(tree.edges) // 2, plan.pendant_branch_of(tree, tree.ntaxa - 1)):
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
        assert max(o.get_scale_num(int(op["dst_slot"])).max() for op in ops) >= 1, "case does not rescale"
        assert abs(ll_g - ll_o) < 1e-6
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s))
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
    g.close()


def test_nonrev_with_invariant_sites_and_ambiguity_codes():
    tree, model, pat, freq, invar = make_nonrev_case(4, 4, 1, ntaxa=10, nptn=400, seed=23, pinvar=0.15)
    assert (invar > 0).any()
    compare_nonrev(tree, model, pat, freq, invar, False, [1, 5])


@pytest.mark.parametrize("kind", ["dna", "aa"])
def test_reversible_model_through_the_nonrev_kernel_unrooted(kind):
    """--kernel-nonrev on an unrooted tree: no root leaf, no direction, matrix rows weighted by the state frequencies"""
    tree, model, pat, freq, invar = make_case(kind, ntaxa=12, nsites=300, seed=4)
    n, nmix = model.nstates, model.nmix
    z = lambda a: np.stack([a, np.zeros_like(a)], axis=-1)
    model_nr = type(model)(**{k: getattr(model, k) for k in model.__dataclass_fields__})
    model_nr.nonrev = dict(ceval=z(model.eval.reshape(nmix, n)), cevec=z(model.evec.reshape(nmix, n, n)), cinv_evec=z(model.inv_evec.reshape(nmix, n, n)),
                           rate_matrix=np.stack(model.Q), rooted=False)
    rev = LikelihoodEngine(model, tree.ntaxa, pat.shape[1])
    nr = LikelihoodEngine(model_nr, tree.ntaxa, pat.shape[1])
    for e in (rev, nr):
        e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(invar)
    for br in (0, 3, plan.pendant_branch_of(tree, 2)):
        ops, a, b, L = plan.plan_branch(tree, br)
        l1 = rev.branch_logl(ops, a, b, L)
        l2 = nr.branch_logl(ops, a, b, L)
        assert abs(l1 - l2) < 1e-8 * abs(l1)
        np.testing.assert_allclose(nr.get_pattern_lh(), rev.get_pattern_lh(), rtol=1e-10)
        rev.theta_reset()
        np.testing.assert_allclose(nr.branch_derv(ops[:0], a, b, L), rev.branch_derv(ops[:0], a, b, L), rtol=1e-8)
    rev.close(); nr.close()


def test_nonrev_newton_branch_and_unsupported_calls():
    tree, model, pat, freq, invar = make_nonrev_case(4, 4, 1, ntaxa=12, nptn=500, seed=29)
    g = nonrev_engine(LikelihoodEngine, tree, model, pat, freq, invar, False)
    o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, False)
    for br in (2, plan.pendant_branch_of(tree, 3)):
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        o.branch_logl(ops, a, b, L)
        x_g, steps = g.newton_branch(ops, a, b, 1e-6, L, 10.0, 1e-6)
        x_o = newton_ref(lambda x: o.branch_derv(ops[:0], a, b, x), 1e-6, L, 10.0, 1e-6, 100)
        assert abs(x_g - x_o) < 1e-6, (x_g, x_o)
        assert steps >= 1
    with pytest.raises(Exception):
        g.logl_from_buffer(0.1)      # the reference has no from-buffer form for non-reversible models
    g.close()


def test_nonrev_against_the_live_reference():
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libiqref.so not built")
    for nstates, ncat, safe in ((4, 4, False), (20, 4, True)):
        tree, model, pat, freq, invar = make_nonrev_case(nstates, ncat, 1, ntaxa=30, nptn=1200, seed=31)
        r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=4)
        g = nonrev_engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
        for leaf in (0, 11, pat.shape[0]):
            ops, a, b, L = plan.plan_branch_nonrev(tree, plan.pendant_branch_of(tree, leaf))
            ll_r, plh_r = r.branch_logl(leaf)
            ll_g = g.branch_logl(ops, a, b, L)
            assert abs(ll_r - ll_g) < 1e-6
            np.testing.assert_allclose(g.get_pattern_lh(), plh_r, rtol=1e-11, atol=0)
            if leaf != pat.shape[0]:
                np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, L), r.branch_derv(leaf), rtol=1e-9)
        g.close(); r.close()


@pytest.mark.parametrize("safe", [False, True])
@pytest.mark.parametrize("ncat", [4, 1, 3])
def test_nonrev_4_state_walker_and_thread_per_op_kernel_agree(ncat, safe):
    """k_nonrev_walk4 (register-chained walker; 4 states x 1 or 4 categories) and k_nonrev_partials (kernel variant 5, also what
    3 categories run on) against the restatement: CLVs, scale counts, logL — on a deep tree that rescales and on a bushy one"""
    for tree_kind, ntaxa, mean_len in (("yule", 40, 0.1), ("caterpillar", 150, 0.8)):
        tree, model, pat, freq, invar = make_nonrev_case(4, ncat, 1, ntaxa=ntaxa, nptn=200, seed=43, tree_kind=tree_kind, mean_len=mean_len,
                                                         p_unknown=0.0 if tree_kind == "caterpillar" else 0.05)
        o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
        br = plan.pendant_branch_of(tree, 0)
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        ll_o = o.branch_logl(ops, a, b, L)
        for variant in (0, 5):
            g = nonrev_engine(LikelihoodEngine, tree, model, pat, freq, invar, safe)
            if variant:
                g.set_schedule(10 + variant)
            ll_g = g.branch_logl(ops, a, b, L)
            assert abs(ll_g - ll_o) < 1e-9 * abs(ll_o), (tree_kind, variant, ll_g, ll_o)
            for op in ops:
                s = int(op["dst_slot"])
                assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), (tree_kind, variant, s)
                assert_partials_close(g.get_partial(s), o.get_partial(s), 1e-11)
            # a partial traversal on resident CLVs: only the last op again (its children come from memory, nothing is chained)
            assert g.branch_logl(ops[-1:], a, b, L) == ll_g
            g.close()
