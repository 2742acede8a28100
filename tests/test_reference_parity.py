"""CPU: the oracle (oracle/lk_oracle.c) and the host-side tree logic pinned against the REAL reference kernel —
IQ-TREE 2.4.0's own translation units compiled in place into oracle/_ref/libiqref.so (oracle/build_ref.sh) and driven
through PhyloTree's public members by oracle/ref_harness.cpp.  Bars (BASELINE.json north_star): scale_num bit-exact,
per-pattern logL 1e-11 relative, tree logL 1e-6, optimised branch lengths 1e-6.  The oracle follows the reference's
summation orders, so the observed differences are at the 1e-14 level and the asserts are set far tighter than the bars."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from oracle import ref_engine
from oracle.oracle_engine import OracleEngine
from refmap import branch_splits, pytree_leafset, pytree_splits, slot_leafsets
from test_tree_logic_cpu import example_case, make_pytree

pytestmark = pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built (needs /root/reference)")

# (kind, safe) — safe is what the reference itself would pick (tree/phylotreesse.cpp:95-96: SAFE unless nstates in {4,20}),
# plus explicit -safe runs for DNA / protein
CASES = [("dna", False), ("dna", True), ("dna_inv", False), ("dna_g1", False), ("dna_g8", False), ("aa", False), ("aa", True),
         ("aa_c4", False), ("aa_c4", True), ("aa_fused", False), ("aa_c60", False), ("bin", True), ("codon", True), ("n5mix", True),
         ("n7", True)]


def run_pair(kind, safe, ntaxa=14, nsites=400, leaf=0, **kw):
    if kind == "aa_c60":
        ntaxa, nsites = 8, 60
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=nsites, **kw)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe)
    assert r.info()["safe_numeric"] == safe and r.info()["vector_size"] == 4
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    o.set_reference_quirks(1)
    br = plan.pendant_branch_of(tree, leaf)
    ops, a, b, L = plan.plan_branch(tree, br)
    return tree, model, pat, freq, inv, r, o, br, ops, a, b, L


@pytest.mark.parametrize("kind,safe", CASES)
def test_oracle_matches_reference_kernel(kind, safe):
    tree, model, pat, freq, inv, r, o, br, ops, a, b, L = run_pair(kind, safe)
    np.testing.assert_array_equal(r.ptn_invar(), inv)                      # computePtnInvar + computeConst
    np.testing.assert_allclose(o.get_tip_partial_lh().ravel(), r.tip_partial_lh().ravel(), rtol=1e-13, atol=1e-15)
    ll_r, plh_r, cat_r = r.branch_logl(0, want_cat=True)
    ll_o = o.branch_logl(ops, a, b, L)
    assert abs(ll_o - ll_r) < 1e-9, (ll_o, ll_r)
    np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12)
    cmax = np.abs(cat_r).max(axis=1, keepdims=True)   # categories that contribute nothing to a pattern carry cancellation noise
    np.testing.assert_allclose(o.get_pattern_lh_cat().reshape(cat_r.shape) / cmax, cat_r / cmax, rtol=1e-11, atol=1e-13)
    assert r.last_traversal_size() == len(ops)
    # every CLV of the traversal: scale_num bit-exact, partial_lh to 1e-12 of the pattern's largest entry
    ref = r.directed_partials()
    sets = slot_leafsets(tree, br)
    assert len(ops) == tree.ntaxa - 2
    for op in ops:
        slot = int(op["dst_slot"])
        part_r, sc_r = ref[sets[slot]]
        np.testing.assert_array_equal(o.get_scale_num(slot).reshape(sc_r.shape), sc_r)
        part_o = o.get_partial(slot).reshape(part_r.shape)
        scale = np.abs(part_r).max(axis=1, keepdims=True)
        np.testing.assert_allclose(part_o / scale, part_r / scale, rtol=0, atol=1e-12)
    # derivatives (theta buffer, df, ddf) and the from-buffer likelihood
    df_r, ddf_r = r.branch_derv(0)
    df_o, ddf_o = o.branch_derv(ops[:0], a, b, L)
    assert abs(df_o - df_r) <= 1e-9 * max(1.0, abs(df_r)), (df_o, df_r)
    assert abs(ddf_o - ddf_r) <= 1e-9 * max(1.0, abs(ddf_r)), (ddf_o, ddf_r)
    th_r, bs_r = r.theta()
    th_o = o.get_theta().reshape(th_r.shape)
    scale = np.abs(th_r).max(axis=1, keepdims=True) + 1e-300
    np.testing.assert_allclose(th_o / scale, th_r / scale, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(o.get_buffer_scale_all(), bs_r)
    assert abs(o.logl_from_buffer(L) - r.logl_from_buffer()) < 1e-9


def test_reference_dotproducttriple_double_counts_last_element():
    """Documents the reference defect behind B200LK_QUIRK_DOTTRIPLE_ODD: with odd nstates and an even block the
    reference's (df, ddf) are NOT the derivatives of its own logL; without the quirk the oracle's are."""
    tree, model, pat, freq, inv, r, o, br, ops, a, b, L = run_pair("n5mix", True)
    assert model.nstates % 2 == 1 and (model.nstates * model.ncat_mix) % 2 == 0
    o.branch_logl(ops, a, b, L)
    df_q, _ = o.branch_derv(ops[:0], a, b, L)
    df_r, _ = r.branch_derv(0)
    assert abs(df_q - df_r) <= 1e-9 * max(1.0, abs(df_r))
    o.set_reference_quirks(0)
    o.theta_reset()
    df_m, ddf_m = o.branch_derv(ops[:0], a, b, L)
    h = 1e-6
    num = (o.logl_from_buffer(L + h) - o.logl_from_buffer(L - h)) / (2 * h)
    assert abs(df_m - num) < 1e-4 * max(1.0, abs(num))
    assert abs(df_r - num) > 1e-2 * max(1.0, abs(num))


@pytest.mark.parametrize("kind,safe", [("dna", False), ("dna", True), ("aa", False)])
def test_scale_num_parity_on_deep_trees(kind, safe):
    """caterpillar trees with long branches: scale_num climbs well above 1 and must stay bit-exact"""
    tree, model, pat, freq, inv, r, o, br, ops, a, b, L = run_pair(kind, safe, ntaxa=500 if kind == "dna" else 200, nsites=40,
                                                               tree_kind="caterpillar", mean_len=1.5, gaps=0.0, ambig=0.0)
    ll_r, plh_r = r.branch_logl(0)
    ll_o = o.branch_logl(ops, a, b, L)
    ref = r.directed_partials(want_partial=False)
    sets = slot_leafsets(tree, br)
    top = 0
    for op in ops:
        slot = int(op["dst_slot"])
        sc_o = o.get_scale_num(slot)
        np.testing.assert_array_equal(sc_o.reshape(ref[sets[slot]][1].shape), ref[sets[slot]][1])
        top = max(top, int(sc_o.max()))
    assert top >= 2, top
    np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12)
    assert abs(ll_o - ll_r) < 1e-8 * max(1.0, abs(ll_r) * 1e-3)


def test_reference_threads_agree():
    tree, model, pat, freq, inv = make_case("dna", ntaxa=20, nsites=3000)
    r1 = ref_engine.RefTree(tree, model, pat, freq, nthreads=1)
    r4 = ref_engine.RefTree(tree, model, pat, freq, nthreads=4)
    assert r4.info()["num_threads"] == 4
    (l1, p1), (l4, p4) = r1.branch_logl(3), r4.branch_logl(3)
    np.testing.assert_array_equal(p1, p4)
    assert abs(l1 - l4) < 1e-8


def test_config1_tree_level_logic_matches_reference():
    """BASELINE config 1 (example.phy + NJ tree, GTR+G4): computeLikelihood and optimizeAllBranches of the compiled
    reference against the Python restatement of the tree logic over the oracle."""
    names, pat, freq, model, newick = example_case()
    r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
    t = make_pytree(names, pat, freq, model, newick)
    ll_r, plh_r = r.compute_likelihood()
    ll_t = t.compute_likelihood()
    assert abs(ll_t - ll_r) < 1e-8
    np.testing.assert_allclose(t.engine.get_pattern_lh(), plh_r, rtol=1e-12)
    assert r.last_traversal_size() == t.num_partial == 15
    ll_r = r.optimize_all_branches(100, 0.001, 100)
    ll_t = t.optimize_all_branches(100, 0.001, 100)
    assert abs(ll_t - ll_r) < 1e-6, (ll_t, ll_r)
    ref_len = r.branches()
    for split, ln in zip(pytree_splits(t), t.lengths()):
        assert abs(ref_len[split] - ln) < 1e-6, (sorted(split), ref_len[split], ln)


def test_branch_optimisation_order_and_lengths_on_a_larger_tree():
    """40 taxa x 3000 sites with gaps: same branch order (computeBestTraversal), same Newton-Raphson path — lengths agree to
    ~1e-15, far inside the 1e-6 bar (the NR tolerance is params->min_branch_length, main/phyloanalysis.cpp:4424-4430)"""
    tree, model, pat, freq, inv = make_case("dna", ntaxa=40, nsites=3000, seed=5, gaps=0.02, ambig=0.0)
    names = ["T%d" % i for i in range(tree.ntaxa)]
    tree.names = names
    newick = tree.to_newick(fmt="%.12g")
    r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
    t = make_pytree(names, pat, freq, model, newick)
    assert abs(r.compute_likelihood()[0] - t.compute_likelihood()) < 1e-8
    far = t.farthest_leaf(); t.farthest_leaf(far)
    nodes, nodes2 = [], []
    t.pre_order_branches(far, None, nodes, nodes2)
    spl = pytree_splits(t)
    assert r.best_traversal() == [spl[n1.find(n2).id] for n1, n2 in zip(nodes, nodes2)]
    ll_r, ll_t = r.optimize_all_branches(100, 0.001, 100), t.optimize_all_branches(100, 0.001, 100)
    assert abs(ll_r - ll_t) < 1e-7
    ref_len = r.branches()
    for split, ln in zip(spl, t.lengths()):
        assert abs(ref_len[split] - ln) < 1e-9


@pytest.mark.parametrize("mem,tree_kind", [(0.0, "yule"), (0.4, "yule"), (0.0, "caterpillar")])
def test_mem_save_slot_table_matches_reference(mem, tree_kind):
    """-mem (LM_MEM_SAVE): the host mirror's MemSlotVector against the reference's (tree/memslot.cpp) — same slot budget,
    the same number of partial updates for every move of the evaluation branch, and the reference's own logL is unchanged
    by the mode (so an op list that recomputes the same directed partials is the same computation)."""
    from iqtree2_b200.phylotree import PhyloTree
    tree, model, pat, freq, inv = make_case("dna", ntaxa=26, nsites=80, tree_kind=tree_kind)
    r_full = ref_engine.RefTree(tree, model, pat, freq)
    ll_true = r_full.branch_logl(0)[0]
    r_full.close()
    ref_engine.set_mem_save(True, mem)   # a process-wide switch (Params::getInstance()): it must stay on while r lives
    try:
        r = ref_engine.RefTree(tree, model, pat, freq)
        t = PhyloTree()
        t.set_plan_only(True)
        t.set_alignment(tree.names, pat, freq, model)
        t.set_model(model)
        t.read_tree(tree.to_newick())
        t.set_mem_save(True, mem)
        assert t.max_lh_slots == r.max_lh_slots
        o = load_engine(OracleEngine, tree, model, pat, freq, inv, nslots=t.max_lh_slots + 2)
        pa, pb, _ = t.branches()
        pendant = {int(min(x, y)): k for k, (x, y) in enumerate(zip(pa, pb)) if min(x, y) < tree.ntaxa}
        for leaf in [0, 7, 19, 3, 3, 25, 12, 0, 1, 24, 8]:
            ll_r = r.branch_logl(leaf)[0]
            assert abs(ll_r - ll_true) < 1e-8
            ops, a, b, L = t.plan_branch(pendant[leaf])
            assert len(ops) == r.last_traversal_size(), (leaf, len(ops), r.last_traversal_size())
            assert abs(o.branch_logl(ops, a, b, L) - ll_r) < 1e-8
        r.close(); o.close(); t.close()
    finally:
        ref_engine.set_mem_save(False)


@pytest.mark.parametrize("kind,nni5", [("dna", True), ("dna", False), ("aa", True)])
def test_nni_evaluation_matches_reference(kind, nni5):
    """PhyloTree::getBestNNIForBran (tree/phylotree.cpp:4069-4322) on every inner branch: the compiled reference against the
    Python restatement over the oracle — same two moves (identified by the swapped subtrees), scores to 1e-6, the five
    re-optimised lengths to 1e-6 (observed ~1e-12), and the tree, its lengths and its likelihood untouched afterwards."""
    tree, model, pat, freq, inv = make_case(kind, ntaxa=12, nsites=500 if kind == "dna" else 200, seed=11, gaps=0.02, ambig=0.0)
    names = ["T%d" % i for i in range(tree.ntaxa)]
    tree.names = names
    newick = tree.to_newick(fmt="%.12g")
    r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
    t = make_pytree(names, pat, freq, model, newick)
    ll0 = t.compute_likelihood()
    assert abs(r.compute_likelihood()[0] - ll0) < 1e-8
    len0 = t.lengths().copy()
    spl = pytree_splits(t)
    inner = [(ab, ba) for ab, ba in t.branches if not ab.node.is_leaf() and not ba.node.is_leaf()]
    assert len(inner) == tree.ntaxa - 3
    for ab, ba in inner:
        ref = r.best_nni(spl[ab.id], nni5=nni5)
        # the end the reference takes as node1 (its getBranches orientation) decides the order of the five branch optimisations:
        # its first swapped subtree hangs off node1, its second off node2
        node1, node2 = ba.node, ab.node
        if not ref[0]["swap"][1] <= pytree_leafset(node2, node1):
            node1, node2 = node2, node1
        got = t.get_best_nni_for_bran(node1, node2, nni5=nni5)
        key = lambda m: frozenset(m["swap"])
        ref_by = {key(m): m for m in ref}
        for mv in got:
            sw = (pytree_leafset(mv["swap"][0], node1), pytree_leafset(mv["swap"][1], node2))
            rm = ref_by[frozenset(sw)]
            assert abs(mv["score"] - rm["score"]) < 1e-6, (mv["score"], rm["score"])
            assert abs(mv["new_len"][0] - rm["new_len"][0]) < 1e-6
            if nni5:
                # outer branches keyed by the subtree they lead to (after the swap)
                mine = {}
                for i, x in enumerate(mv["outer"]):
                    mine[pytree_leafset(x, node1 if node1.find(x) else node2)] = mv["new_len"][1 + i]
                theirs = dict(zip(rm["outer"], rm["new_len"][1:]))
                assert set(mine) == set(theirs)
                for k2 in mine:
                    assert abs(mine[k2] - theirs[k2]) < 1e-6, (sorted(k2), mine[k2], theirs[k2])
        np.testing.assert_array_equal(t.lengths(), len0)
        assert pytree_splits(t) == spl
    # nothing is stale: the likelihood at the old branch and after a full recomputation is the old one
    assert abs(t.compute_likelihood() - ll0) < 1e-9
    t.clear_all()
    assert abs(t.compute_likelihood() - ll0) < 1e-9
    assert abs(r.compute_likelihood()[0] - ll0) < 1e-8
    # applying the better move of the last branch gives the tree whose likelihood the evaluation promised
    best = max(got, key=lambda m: m["score"])
    t.do_nni(node1, node2, best)
    assert abs(t.compute_likelihood() - best["score"]) < 1e-6
    t.clear_all()
    assert abs(t.compute_likelihood() - best["score"]) < 1e-6


@pytest.mark.parametrize("kind,start,bounds", [("dna", 1.0, (0.01, 100.0)), ("dna", 2.5, (0.01, 100.0)), ("aa", 0.4, (1e-4, 10.0))])
def test_tree_length_scaling_matches_reference(kind, start, bounds):
    """PhyloTree::optimizeTreeLengthScaling (tree/phylotree.cpp:2457-2488; Brent's method over clearAllPartialLH + computeLikelihood,
    the M-step of RateFree::optimizeWithEM and `--tree-len-scale`): same evaluation points, hence the same optimum to the last digits"""
    tree, model, pat, freq, inv = make_case(kind, ntaxa=12, nsites=400 if kind == "dna" else 150, seed=13)
    tree.lengths = tree.lengths * 1.7          # a tree that is too long: the optimum is well below 1
    names = ["T%d" % i for i in range(tree.ntaxa)]
    tree.names = names
    newick = tree.to_newick(fmt="%.17g")
    r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
    t = make_pytree(names, pat, freq, model, newick)
    ll0 = t.compute_likelihood()
    assert abs(r.compute_likelihood()[0] - ll0) < 1e-8
    len0 = t.lengths().copy()
    ll_r, sc_r = r.optimize_tree_length_scaling(bounds[0], start, bounds[1], 0.001)
    ll_t, sc_t = t.optimize_tree_length_scaling(bounds[0], start, bounds[1], 0.001)
    assert abs(sc_r - sc_t) < 1e-9, (sc_r, sc_t)
    assert abs(ll_r - ll_t) < 1e-7, (ll_r, ll_t)
    assert ll_t >= ll0 - 1e-9
    np.testing.assert_allclose(t.lengths(), len0 * (sc_t / start), rtol=1e-12)
    ref_len = r.branches()
    for split, ln in zip(pytree_splits(t), t.lengths()):
        assert abs(ref_len[split] - ln) < 1e-9
