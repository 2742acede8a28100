"""GPU: NNI evaluation for the edge-linked partition model (b200::PhyloSuperTreePlen::getBestNNIForBran / doNNIAll; the role of
tree/phylosupertreeplen.cpp:1100-1500 with every partition tree on the full topology).  Two anchors:
  * a single partition with rate 1 is the plain model: scores, lengths and moves equal PhyloTree::getBestNNIForBran (host-driven
    Newton-Raphson on both sides), which is pinned against the compiled reference;
  * several partitions with their own rates: the score an evaluation reports is the likelihood of the super tree after the move is
    applied with the reported lengths and everything is recomputed from scratch, and the evaluation leaves the tree untouched.
tests/test_host_mirror_cpu.py runs the same checks on the CPU with the host mirrors linked against the oracle."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200.phylotree import PhyloTree
from iqtree2_b200.supertree import PhyloSuperTree
from supercases import make_supercase

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nni5", [True, False])
def test_single_partition_equals_plain_tree(nni5):
    tree, model, pat, freq, inv = make_case("dna", ntaxa=12, nsites=400, seed=17, gaps=0.02, ambig=0.0)
    names = ["T%d" % i for i in range(tree.ntaxa)]
    tree.names = names
    nwk = tree.to_newick(fmt="%.17g")
    s = PhyloSuperTree(names)
    s.add_partition(pat, freq, model, part_rate=1.0, ptn_invar=inv)
    s.read_tree(nwk)
    t = PhyloTree(fused_newton=False)
    t.set_alignment(names, pat, freq, model); t.set_pattern_invar(inv); t.set_model(model); t.read_tree(nwk)
    assert abs(s.compute_likelihood() - t.compute_likelihood()) < 1e-8
    a, b, _ = t.branches()
    sa, sb, _ = s.branches()
    np.testing.assert_array_equal(a, sa); np.testing.assert_array_equal(b, sb)
    n = 0
    for br in range(len(a)):
        if a[br] < tree.ntaxa or b[br] < tree.ntaxa:
            continue
        for rev in (False, True):
            ms, mt = s.get_best_nni_for_bran(br, nni5=nni5, reverse=rev), t.get_best_nni_for_bran(br, nni5=nni5, reverse=rev)
            for x, y in zip(ms, mt):
                assert x["swap"] == y["swap"] and x["outer"] == y["outer"]
                assert abs(x["score"] - y["score"]) < 1e-6, (br, x["score"], y["score"])
                np.testing.assert_allclose(x["new_len"], y["new_len"], rtol=0, atol=1e-7)
        n += 1
    assert n == tree.ntaxa - 3
    s.close()


def test_partitioned_nni_scores_are_the_likelihoods_of_the_moved_trees():
    tree, parts = make_supercase(ntaxa=12, nparts=5, seed=9, missing=0.2)

    def build():
        g = PhyloSuperTree(tree.names)
        for p in parts:
            g.add_partition(p["pat"], p["freq"], p["model"], part_rate=p["rate"], ptn_invar=p["inv"])
        g.read_tree(tree.to_newick(fmt="%.17g"))
        return g

    g = build()
    ll0 = g.compute_likelihood()
    a, b, len0 = g.branches()
    inner = [br for br in range(len(a)) if a[br] >= tree.ntaxa and b[br] >= tree.ntaxa]
    better = 0
    for br in inner[::2]:
        moves = g.get_best_nni_for_bran(br)
        a2, b2, len2 = g.branches()
        np.testing.assert_array_equal(a2, a); np.testing.assert_array_equal(b2, b); np.testing.assert_array_equal(len2, len0)
        for mv in moves:
            h = build()
            h.compute_likelihood()
            h.do_nni(br, mv["move"], new_len=mv["new_len"])
            h.clear_all_partial_lh()
            assert abs(h.compute_likelihood() - mv["score"]) < 1e-6, (br, mv["move"], mv["score"])
            better += mv["score"] > ll0
            h.close()
    # nothing stale after the evaluations; incremental application of a move agrees with the promise as well
    assert abs(g.compute_likelihood() - ll0) < 1e-7
    g.clear_all_partial_lh()
    assert abs(g.compute_likelihood() - ll0) < 1e-7
    mv = max(g.get_best_nni_for_bran(inner[1]), key=lambda m: m["score"])
    g.do_nni(inner[1], mv["move"], new_len=mv["new_len"])
    assert abs(g.compute_likelihood() - mv["score"]) < 1e-6
    g.close()
