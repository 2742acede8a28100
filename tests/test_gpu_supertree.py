"""GPU: the edge-linked partition model (b200::PhyloSuperTreePlen over one device engine per partition, all partitions
queued concurrently) against its restatement over the CPU oracle (tests/pysupertree.py, pinned against the compiled
reference on induced subtrees in test_supertree_cpu.py).  Bars: super-tree logL 1e-6, partition scores 1e-6,
linked derivatives 1e-9 relative, optimised super branch lengths 1e-6."""
import numpy as np
import pytest

from iqtree2_b200.supertree import PhyloSuperTree
from supercases import make_supercase
from test_supertree_cpu import make_pysuper

pytestmark = pytest.mark.gpu


def make_gpu_super(tree, parts, num_threads=0, safe=False):
    g = PhyloSuperTree(tree.names, num_threads=num_threads)
    for p in parts:
        g.add_partition(p["pat"], p["freq"], p["model"], part_rate=p["rate"], ptn_invar=p["inv"], safe_numeric=safe)
    g.read_tree(tree.to_newick(fmt="%.17g"))
    return g


@pytest.mark.parametrize("safe", [False, True])
def test_supertree_logl_scores_and_derivatives(safe):
    tree, parts = make_supercase(ntaxa=14, nparts=7, missing=0.2)
    g = make_gpu_super(tree, parts, safe=safe)
    o = make_pysuper(tree, parts)
    ll_g, ll_o = g.compute_likelihood(), o.compute_likelihood()
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    np.testing.assert_allclose(g.part_scores(), o.cur_score, rtol=0, atol=1e-6)
    _, _, ln = g.branches()
    for bid in (0, 5, 11, len(ln) - 1):
        assert abs(g.compute_likelihood_branch(bid) - ll_o) < 1e-6
        for x in (ln[bid], 0.5 * ln[bid] + 0.01):
            np.testing.assert_allclose(g.compute_likelihood_derv(bid, x), o.func_derv_at(bid, x), rtol=1e-9, atol=1e-7)
        # computeFuncDerv leaves the branch at the last evaluated length (as the reference's does): put it back and drop
        # the partials that saw the other value
        g.set_branch_length(bid, ln[bid], clear_partials=True)
        o._set_lengths(ln[bid])
        o.clear_all()
    g.close()


def test_supertree_thread_count_does_not_change_results():
    tree, parts = make_supercase(ntaxa=10, nparts=9, seed=8)
    res = []
    for nt in (1, 4):
        g = make_gpu_super(tree, parts, num_threads=nt)
        ll = g.compute_likelihood()
        d = g.compute_likelihood_derv(3, 0.07)
        res.append((ll, d, g.part_scores().copy()))
        g.close()
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    np.testing.assert_array_equal(res[0][2], res[1][2])


def test_supertree_branch_length_optimisation():
    tree, parts = make_supercase(ntaxa=12, nparts=5, seed=3, sites=(100, 250))
    g = make_gpu_super(tree, parts)
    o = make_pysuper(tree, parts)
    ll_g = g.optimize_all_branches(iterations=100, tolerance=0.001)
    ll_o = o.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    _, _, ln = g.branches()
    np.testing.assert_allclose(ln, o.lengths(), rtol=0, atol=1e-6)
    # a fresh evaluation at the optimised lengths agrees with the running score
    g.clear_all_partial_lh()
    assert abs(g.compute_likelihood() - ll_g) < 1e-6


def test_supertree_incremental_after_one_branch():
    """optimizeOneBranch invalidates only the partitions' reverse partials of that branch; the next evaluation at another
    branch must equal a full recomputation."""
    tree, parts = make_supercase(ntaxa=11, nparts=4, seed=5)
    g = make_gpu_super(tree, parts)
    g.compute_likelihood()
    g.optimize_one_branch(4)
    ll_inc = g.compute_likelihood_branch(9)
    g.clear_all_partial_lh()
    ll_full = g.compute_likelihood_branch(9)
    assert abs(ll_inc - ll_full) < 1e-7


def test_grouped_launches_equal_per_partition_streams(monkeypatch):
    """The grouped launches (b200lk_group_*: one launch per kernel type for all partitions) run the single-engine kernels' own
    device functions with per-member reduction slots: the partition scores, the linked derivatives and the optimised lengths are
    those of one engine + stream per partition, bit for bit — in a handful of launches per evaluation instead of ~5 per partition."""
    tree, parts = make_supercase(ntaxa=16, nparts=13, seed=12, missing=0.1)
    g = make_gpu_super(tree, parts)
    monkeypatch.setenv("B200ST_NO_GROUP", "1")
    s = make_gpu_super(tree, parts)
    ll_s = s.compute_likelihood_branch(5)
    monkeypatch.delenv("B200ST_NO_GROUP")
    ll_g = g.compute_likelihood_branch(5)
    assert ll_g == ll_s
    g.clear_all_partial_lh()               # (the first evaluation also created the engines and uploaded the models)
    l0 = g.launch_count
    assert g.compute_likelihood_branch(5) == ll_s
    n_first = g.launch_count - l0
    np.testing.assert_array_equal(g.part_scores(), s.part_scores())
    _, _, ln = g.branches()
    monkeypatch.setenv("B200ST_NO_GROUP", "1")
    d_s = s.compute_likelihood_derv(5, float(ln[5]) * 0.8)
    monkeypatch.delenv("B200ST_NO_GROUP")
    l1 = g.launch_count
    d_g = g.compute_likelihood_derv(5, float(ln[5]) * 0.8)
    assert d_g == d_s
    # two kernel families (4 and 20 states): staging + walker + conversion + root per family at most
    assert n_first <= 10 and g.launch_count - l1 <= 8, (n_first, g.launch_count - l1)
    monkeypatch.setenv("B200ST_NO_GROUP", "1")
    s.set_branch_length(5, ln[5], clear_partials=True)
    ll_os = s.optimize_all_branches(iterations=3, tolerance=0.001)
    monkeypatch.delenv("B200ST_NO_GROUP")
    g.set_branch_length(5, ln[5], clear_partials=True)
    ll_og = g.optimize_all_branches(iterations=3, tolerance=0.001)
    assert ll_og == ll_os
    np.testing.assert_array_equal(g.branches()[2], s.branches()[2])
    g.close(); s.close()
