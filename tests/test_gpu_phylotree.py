"""GPU: the tree-level API (C++ host mirror over the CUDA engine) against the Python restatement over the CPU oracle.
BASELINE config 1 is example.phy + a committed NJ tree + GTR+G4: fixed-tree logL and branch-length optimisation.
Bars: tree logL 1e-6, per-pattern logL 1e-11 relative, optimised branch lengths 1e-6."""
import numpy as np
import pytest

from iqtree2_b200.phylotree import PhyloTree
from test_tree_logic_cpu import example_case, make_pytree

pytestmark = pytest.mark.gpu


def make_gpu_tree(names, pat, freq, model, newick, **kw):
    t = PhyloTree(**kw)
    t.set_alignment(names, pat, freq, model)
    t.set_model(model)
    t.read_tree(newick)
    return t


def test_config1_fixed_tree_logl_and_pattern_lh():
    names, pat, freq, model, newick = example_case()
    g = make_gpu_tree(names, pat, freq, model, newick)
    o = make_pytree(names, pat, freq, model, newick)
    ll_g, ll_o = g.compute_likelihood(), o.compute_likelihood()
    assert abs(ll_g - ll_o) < 1e-6
    plh, cat = g.compute_pattern_likelihood(with_cat=True)
    np.testing.assert_allclose(plh, o.engine.get_pattern_lh(), rtol=1e-11)
    assert abs(float((plh * freq).sum()) - ll_g) < 1e-7
    # per-category log-likelihoods recombine to the pattern log-likelihood
    np.testing.assert_allclose(np.log(np.exp(cat - plh[:, None]).sum(axis=1)), 0.0, atol=1e-9)
    a, b, ln = g.branches()
    for br in range(0, len(ln), 4):
        assert abs(g.compute_likelihood_branch(br) - ll_o) < 1e-6


@pytest.mark.parametrize("fused", [True, False])
def test_config1_branch_length_optimisation(fused):
    names, pat, freq, model, newick = example_case()
    g = make_gpu_tree(names, pat, freq, model, newick, fused_newton=fused)
    o = make_pytree(names, pat, freq, model, newick)
    ll_g = g.optimize_all_branches(iterations=100, tolerance=0.001)
    ll_o = o.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    _, _, ln = g.branches()
    np.testing.assert_allclose(ln, o.lengths(), rtol=0, atol=1e-6)
    # same amount of partial-likelihood work as the reference's planner would schedule
    assert g.num_partial_lh_computations == o.num_partial
    # and a fresh full evaluation at the optimised lengths agrees
    g.clear_all_partial_lh(); o.clear_all()
    assert abs(g.compute_likelihood() - o.compute_likelihood()) < 1e-6


def test_incremental_invalidation_matches_full_recompute():
    names, pat, freq, model, newick = example_case()
    g = make_gpu_tree(names, pat, freq, model, newick)
    ll0 = g.compute_likelihood()
    n0 = g.num_partial_lh_computations
    assert n0 == 15   # 17 taxa -> 15 internal CLVs
    g.optimize_one_branch(7)
    g.optimize_one_branch(20)
    ll_inc = g.compute_likelihood_branch(3)
    assert g.num_partial_lh_computations - n0 < 3 * 15
    newick2 = g.get_tree()
    g2 = make_gpu_tree(names, pat, freq, model, newick2)
    # get_tree prints 10 significant digits; compare at that resolution
    assert abs(g2.compute_likelihood() - ll_inc) < 1e-4
    assert ll_inc >= ll0 - 1e-9


def test_model_change_invalidates_everything():
    from iqtree2_b200 import models
    names, pat, freq, model, newick = example_case()
    g = make_gpu_tree(names, pat, freq, model, newick)
    ll0 = g.compute_likelihood()
    model2 = models.make_gtr_gamma(rates6=(1.0, 2.0, 1.0, 1.0, 2.0, 1.0), pi=(0.25, 0.25, 0.25, 0.25), alpha=1.0)
    g.set_model(model2)
    ll1 = g.compute_likelihood()
    o = make_pytree(names, pat, freq, model2, newick)
    assert abs(ll1 - o.compute_likelihood()) < 1e-6
    assert abs(ll1 - ll0) > 1.0


def test_protein_tree_api_and_safe_kernel():
    from cases import make_case
    tree, model, pat, freq, inv = make_case("aa", ntaxa=14, nsites=300, seed=8)
    newick = tree.to_newick()
    for safe in (False, True):
        g = make_gpu_tree(tree.names, pat, freq, model, newick, safe_numeric=safe)
        g.set_pattern_invar(inv)
        from oracle.oracle_engine import OracleEngine
        from pytree import PyTree

        def factory(ntaxa, nslots):
            e = OracleEngine(model, ntaxa, pat.shape[1], nslots=nslots, safe_numeric=safe)
            e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(inv)
            return e
        o = PyTree(factory, tree.names, newick)
        assert abs(g.compute_likelihood() - o.compute_likelihood()) < 1e-6
        ll_g = g.optimize_all_branches(iterations=3)
        ll_o = o.optimize_all_branches(iterations=3)
        assert abs(ll_g - ll_o) < 1e-6
        np.testing.assert_allclose(g.branches()[2], o.lengths(), atol=1e-6)


@pytest.mark.parametrize("kind,schedule,mem", [("dna", 0, 0.0), ("dna", 1, 0.0), ("dna", 0, 0.5), ("aa", 0, 0.0), ("aa", 1, 0.3),
                                               ("aa_c4", 0, 0.0), ("n7", 0, 0.0)])
def test_mem_save_slot_recycling_on_gpu(kind, schedule, mem):
    """-mem (LM_MEM_SAVE): a bounded number of recycled CLV slots gives the same likelihoods as one slot per node, for
    the walker kernels (in-order op list, register chaining) and the levelled schedule (WAR/WAW-aware levels)."""
    from cases import make_case
    tree, model, pat, freq, inv = make_case(kind, ntaxa=28, nsites=260, seed=11)
    newick = tree.to_newick()
    full = make_gpu_tree(tree.names, pat, freq, model, newick)
    full.set_pattern_invar(inv)
    ms = make_gpu_tree(tree.names, pat, freq, model, newick, schedule=schedule)
    ms.set_mem_save(True, mem)
    ms.set_pattern_invar(inv)
    assert ms.max_lh_slots == max(int(np.log2(28)) + 1, int(np.floor(mem * 26)))
    nbr = len(full.branches()[0])
    for br in np.random.default_rng(3).permutation(nbr)[:12]:
        ll_f, ll_m = full.compute_likelihood_branch(int(br)), ms.compute_likelihood_branch(int(br))
        assert abs(ll_f - ll_m) < 1e-9 * abs(ll_f), (br, ll_f, ll_m)
        np.testing.assert_allclose(ms.compute_pattern_likelihood(), full.compute_pattern_likelihood(), rtol=1e-12)
        d_f, d_m = full.compute_likelihood_derv(int(br)), ms.compute_likelihood_derv(int(br))
        np.testing.assert_allclose(d_m, d_f, rtol=1e-9)
    assert ms.num_partial_lh_computations >= full.num_partial_lh_computations
    # branch-length optimisation under the slot budget reaches the same optimum
    ll_f = full.optimize_all_branches(iterations=2)
    ll_m = ms.optimize_all_branches(iterations=2)
    assert abs(ll_f - ll_m) < 1e-6
    np.testing.assert_allclose(ms.branches()[2], full.branches()[2], atol=1e-6)


@pytest.mark.parametrize("kind", ["config1", "aa"])
def test_pipelined_sweep_equals_branch_by_branch(kind):
    """optimizeAllBranches with every sweep queued as a whole on the device (branch lengths device-resident between branches,
    Newton-Raphson and the divergence test in k_newton_coop, one synchronisation per sweep) gives the lengths of the
    branch-by-branch path — the same arithmetic in the same order — and therefore the reference's (1e-6)."""
    if kind == "config1":
        names, pat, freq, model, newick = example_case()
    else:
        from cases import make_case
        tree, model, pat, freq, _ = make_case("aa", ntaxa=30, nsites=400, seed=4, mean_len=0.2)
        names, newick = tree.names, tree.to_newick(fmt="%.17g")
    a = make_gpu_tree(names, pat, freq, model, newick)
    b = make_gpu_tree(names, pat, freq, model, newick)
    b.set_pipelined_sweep(True)
    la = a.optimize_all_branches(iterations=100, tolerance=0.001)
    lb = b.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(la - lb) < 1e-8, (la, lb)
    np.testing.assert_allclose(b.branches()[2], a.branches()[2], rtol=0, atol=1e-12)
    o = make_pytree(names, pat, freq, model, newick)
    o.optimize_all_branches(iterations=100, tolerance=0.001)
    np.testing.assert_allclose(b.branches()[2], o.lengths(), rtol=0, atol=1e-6)
    # the resident CLVs are consistent with the final lengths: a fresh evaluation agrees
    ll = b.compute_likelihood()
    b.clear_all_partial_lh()
    assert abs(b.compute_likelihood() - ll) < 1e-7


@pytest.mark.parametrize("kind,start", [("dna", 1.0), ("aa", 0.5)])
def test_tree_length_scaling(kind, start):
    """PhyloTree::optimizeTreeLengthScaling (tree/phylotree.cpp:2457-2488: Brent's method, every evaluation a full traversal)
    against the Python restatement over the oracle, which test_reference_parity.py pins against the compiled reference"""
    from cases import make_case
    tree, model, pat, freq, inv = make_case(kind, ntaxa=14, nsites=500 if kind == "dna" else 200, seed=13)
    tree.lengths = tree.lengths * 1.6
    names, newick = tree.names, tree.to_newick(fmt="%.17g")
    g = make_gpu_tree(names, pat, freq, model, newick)
    g.set_pattern_invar(inv)
    o = make_pytree(names, pat, freq, model, newick, inv=inv)
    ll0 = g.compute_likelihood()
    _, _, len0 = g.branches()
    n0 = g.num_partial_lh_computations
    ll_g, sc_g = g.optimize_tree_length_scaling(0.01, start, 100.0, 0.001)
    ll_o, sc_o = o.optimize_tree_length_scaling(0.01, start, 100.0, 0.001)
    assert abs(sc_g - sc_o) < 1e-6, (sc_g, sc_o)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    assert ll_g > ll0
    np.testing.assert_allclose(g.branches()[2], len0 * (sc_g / start), rtol=1e-12)
    # every Brent evaluation was a full traversal (T-2 partial updates)
    assert (g.num_partial_lh_computations - n0) % (len(names) - 2) == 0
    g.scale_length(2.0)
    np.testing.assert_allclose(g.branches()[2], len0 * (sc_g / start) * 2.0, rtol=1e-12)
    assert g.compute_likelihood() < ll_g
