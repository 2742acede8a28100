"""CPU tier: the non-reversible restatement (oracle/lk_oracle_nonrev.h, tree/phylokernelnonrev.h) against an independent
brute-force pruning of the rooted tree, on every branch (the transposition of the matrices on root-ward branches makes the
likelihood independent of the evaluation branch), its derivatives against finite differences, and — a reversible model through
the non-reversible kernel on an unrooted tree (--kernel-nonrev) — against the reversible restatement."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200 import plan
from nonrev_cases import brute_force_rooted, make_nonrev_case
from oracle.oracle_engine import OracleEngine


def nonrev_engine(cls, tree, model, pat, freq, invar, safe, **kw):
    ntaxa = pat.shape[0]
    e = cls(model, ntaxa, pat.shape[1], nslots=tree.ntaxa - 2 + 2, safe_numeric=safe, **kw)
    e.set_tip_states(pat)
    e.set_pattern_freq(freq)
    e.set_pattern_invar(invar)
    return e


@pytest.mark.parametrize("nstates,ncat,nmix", [(4, 4, 1), (4, 1, 1), (5, 3, 2), (20, 2, 1)])
@pytest.mark.parametrize("safe", [False, True])
def test_nonrev_oracle_matches_brute_force_on_every_branch(nstates, ncat, nmix, safe):
    tree, model, pat, freq, invar = make_nonrev_case(nstates, ncat, nmix, ntaxa=8, nptn=60, seed=5)
    want = brute_force_rooted(tree, model, pat)
    o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    for br in range(len(tree.edges)):
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        ll = o.branch_logl(ops, a, b, L)
        np.testing.assert_allclose(o.get_pattern_lh(), want, rtol=1e-10, atol=1e-10, err_msg=f"branch {br} {a} {b}")
        assert abs(ll - float((want * freq).sum())) < 1e-8


def test_nonrev_oracle_scaling_deep_tree():
    tree, model, pat, freq, invar = make_nonrev_case(4, 4, 1, ntaxa=150, nptn=40, seed=7, tree_kind="caterpillar", mean_len=0.8, p_unknown=0.0)
    want = brute_force_rooted(tree, model, pat)
    for safe in (False, True):
        o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
        for br in (0, len(tree.edges) // 2, plan.pendant_branch_of(tree, tree.ntaxa - 1)):
            ops, a, b, L = plan.plan_branch_nonrev(tree, br)
            o.branch_logl(ops, a, b, L)
            assert max(o.get_scale_num(int(op["dst_slot"])).max() for op in ops) >= 1, "case does not rescale"
            np.testing.assert_allclose(o.get_pattern_lh(), want, rtol=1e-9)


@pytest.mark.parametrize("safe", [False, True])
def test_nonrev_oracle_derivatives_are_the_derivatives_of_its_logl(safe):
    tree, model, pat, freq, invar = make_nonrev_case(4, 4, 1, ntaxa=9, nptn=80, seed=9)
    o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    root_br = plan.pendant_branch_of(tree, tree.ntaxa - 1)
    for br in range(len(tree.edges)):
        if br == root_br:
            continue   # the branch to the root leaf has no length
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        o.branch_logl(ops, a, b, L)
        df, ddf = o.branch_derv(ops[:0], a, b, L)
        h = 1e-5
        lp, lm, l0 = o.branch_logl(ops[:0], a, b, L + h), o.branch_logl(ops[:0], a, b, L - h), o.branch_logl(ops[:0], a, b, L)
        assert abs(df - (lp - lm) / (2 * h)) < 1e-5 * max(1.0, abs(df)), (br, a, b)
        assert abs(ddf - (lp - 2 * l0 + lm) / (h * h)) < 2e-3 * max(1.0, abs(ddf)), (br, a, b)


@pytest.mark.parametrize("kind", ["dna", "aa", "n5mix"])
def test_reversible_model_through_the_nonrev_kernel_on_an_unrooted_tree(kind):
    """--kernel-nonrev: useRevKernel() is false for ANY model (model/modelsubst.h:73-75); the tree stays unrooted, no branch has a
    direction, the evaluation branch's matrix rows are weighted by the state frequencies (phylokernelnonrev.h:1112-1120)."""
    tree, model, pat, freq, invar = make_case(kind, ntaxa=10, nsites=150, seed=4)
    n, nmix = model.nstates, model.nmix
    z = lambda a: np.stack([a, np.zeros_like(a)], axis=-1)
    model_nr = type(model)(**{k: getattr(model, k) for k in model.__dataclass_fields__})
    model_nr.nonrev = dict(ceval=z(model.eval.reshape(nmix, n)), cevec=z(model.evec.reshape(nmix, n, n)), cinv_evec=z(model.inv_evec.reshape(nmix, n, n)),
                           rate_matrix=np.stack(model.Q), rooted=False)
    rev = OracleEngine(model, tree.ntaxa, pat.shape[1])
    nr = OracleEngine(model_nr, tree.ntaxa, pat.shape[1])
    for e in (rev, nr):
        e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(invar)
    for br in (0, 3, plan.pendant_branch_of(tree, 2)):
        ops, a, b, L = plan.plan_branch(tree, br)
        l1 = rev.branch_logl(ops, a, b, L)
        l2 = nr.branch_logl(*plan.plan_branch_nonrev(tree, br, rooted=False))
        assert abs(l1 - l2) < 1e-8 * abs(l1)
        np.testing.assert_allclose(nr.get_pattern_lh(), rev.get_pattern_lh(), rtol=1e-10)
        rev.theta_reset()
        d1 = rev.branch_derv(ops[:0], a, b, L)
        d2 = nr.branch_derv(ops[:0], a, b, L)
        np.testing.assert_allclose(d2, d1, rtol=1e-8)


# ---- the restatement against the compiled reference (oracle/_ref/libiqref.so: tree/phylokernelnonrev.h, Vec4d AVX+FMA) --------------
def _ref_available():
    from oracle import ref_engine
    return ref_engine.available()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("nstates,ncat,nmix,safe", [(4, 4, 1, False), (4, 4, 1, True), (4, 1, 1, False), (5, 4, 2, True), (5, 4, 1, False), (20, 2, 1, False), (20, 4, 1, True)])
def test_nonrev_oracle_matches_the_reference_kernel(nstates, ncat, nmix, safe):
    """every pendant branch and the root branch: logL, _pattern_lh, (df, ddf) of the reference's computeNonrevLikelihoodBranch /
    computeNonrevLikelihoodDerv on a ROOTED tree (direction flags, root leaf) against the restatement driven by plan_branch_nonrev.
    (Shapes with block * nstates not a multiple of the SIMD width — e.g. 5 states x 3 categories — cannot be compared: the reference
    itself crashes there, computeNonrevLikelihoodBranch carves its tip buffer at buffer_partial_lh + block*nstates without rounding up
    (phylokernelnonrev.h:1099-1100) and then reads it with aligned vector loads.)"""
    from oracle import ref_engine
    tree, model, pat, freq, invar = make_nonrev_case(nstates, ncat, nmix, ntaxa=11, nptn=160, seed=13)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=2)
    o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
    if nstates % 2 == 1:
        o.set_reference_quirks(1)
    for leaf in list(range(0, pat.shape[0], 3)) + [pat.shape[0]]:      # pat.shape[0] = the root leaf
        br = plan.pendant_branch_of(tree, leaf)
        ops, a, b, L = plan.plan_branch_nonrev(tree, br)
        ll_r, plh_r = r.branch_logl(leaf)
        ll_o = o.branch_logl(ops, a, b, L)
        assert abs(ll_r - ll_o) < 1e-9 * abs(ll_r), (leaf, ll_r, ll_o)
        np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12, atol=0)
        if leaf != pat.shape[0]:
            d_r = r.branch_derv(leaf)
            d_o = o.branch_derv(ops[:0], a, b, L)
            np.testing.assert_allclose(d_o, d_r, rtol=1e-9)
    r.close()


@pytest.mark.skipif(not _ref_available(), reason="oracle/_ref/libiqref.so not built")
def test_nonrev_oracle_matches_the_reference_on_a_deep_tree():
    from oracle import ref_engine
    tree, model, pat, freq, invar = make_nonrev_case(4, 4, 1, ntaxa=150, nptn=64, seed=7, tree_kind="caterpillar", mean_len=0.8, p_unknown=0.0)
    for safe in (False, True):
        r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=1)
        o = nonrev_engine(OracleEngine, tree, model, pat, freq, invar, safe)
        for leaf in (0, 77, 150):
            ops, a, b, L = plan.plan_branch_nonrev(tree, plan.pendant_branch_of(tree, leaf))
            ll_r, plh_r = r.branch_logl(leaf)
            o.branch_logl(ops, a, b, L)
            np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-11, atol=0)
        r.close()
