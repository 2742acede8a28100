"""GPU: the edge-unlinked partition model (b200::PhyloSuperTree over independent partition trees, b200ut_*; the reference's
PhyloSuperTree, tree/phylosupertree.cpp:653-791): logL = sum of the partitions' logL on their induced subtrees, pattern
log-likelihoods concatenated, optimizeAllBranches = each partition's own with tolerance / min(ntrees, 10) — against the
Python restatement of the tree logic over the CPU oracle, partition by partition."""
import numpy as np
import pytest

from iqtree2_b200.phylotree import PhyloTree
from iqtree2_b200.supertree import PhyloSuperTreeUnlinked
from supercases import induced_subtree, make_supercase
from test_tree_logic_cpu import make_pytree

pytestmark = pytest.mark.gpu


def build(nparts=5, num_threads=0):
    tree, parts = make_supercase(ntaxa=14, nparts=nparts, seed=7, missing=0.2)
    st = PhyloSuperTreeUnlinked(num_threads=num_threads)
    refs = []
    for p in parts:
        sub, kept = induced_subtree(tree, p["present"])
        pat = np.ascontiguousarray(p["pat"][kept, :])
        nwk = sub.to_newick(fmt="%.17g")
        g = PhyloTree()
        g.set_alignment(sub.names, pat, p["freq"], p["model"]); g.set_pattern_invar(p["inv"]); g.set_model(p["model"])
        g.read_tree(nwk)
        st.add_partition(g)
        refs.append(make_pytree(sub.names, pat, p["freq"], p["model"], nwk, inv=p["inv"]))
    return st, refs, parts


@pytest.mark.parametrize("num_threads", [1, 4])
def test_unlinked_partitions(num_threads):
    st, refs, parts = build(num_threads=num_threads)
    exp = [r.compute_likelihood() for r in refs]
    ll, plh = st.compute_likelihood(want_pattern_lh=True)
    assert abs(ll - sum(exp)) < 1e-6
    np.testing.assert_allclose(st.part_scores(), exp, rtol=0, atol=1e-6)
    np.testing.assert_allclose(plh, np.concatenate([r.engine.get_pattern_lh() for r in refs]), rtol=1e-11)
    plh2, ll2 = st.compute_pattern_likelihood()
    np.testing.assert_array_equal(plh2, plh)
    assert abs(ll2 - ll) < 1e-9
    tol = 0.001 / min(len(refs), 10)
    exp_opt = [r.optimize_all_branches(iterations=100, tolerance=tol) for r in refs]
    got = st.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(got - sum(exp_opt)) < 1e-6, (got, sum(exp_opt))
    assert got > ll
    for g, r in zip(st.parts, refs):
        np.testing.assert_allclose(g.branches()[2], r.lengths(), rtol=0, atol=1e-6)
    # each partition kept its own lengths; a fresh evaluation agrees with the running score
    st.clear_all_partial_lh()
    assert abs(st.compute_likelihood() - got) < 1e-6
    st.close()
