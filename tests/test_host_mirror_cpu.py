"""CPU: the product's C++ host mirrors (b200::PhyloTree, PhyloSuperTreePlen, PhyloSuperTree; iqtree2_b200/csrc/*.cpp) linked against
the CPU oracle instead of the CUDA library (oracle/b200lk_over_oracle.c -> oracle/_build/libb200pt_over_oracle.so, test
infrastructure), so that their LOGIC — traversal planning, optimizeOneBranch / optimizeAllBranches, getBestNNIForBran on the spare
slots, doNNI, optimizeTreeLengthScaling, the partition drivers — is covered without a GPU.  The checks are the tree-level GPU tests
themselves, run with the library loader pointed at that build: same assertions, same bars (reference golden vectors, the Python
restatement pinned against the compiled reference)."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libb200pt_over_oracle.so")


@pytest.fixture()
def over_oracle(monkeypatch):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    lib = ctypes.CDLL(LIB)
    import iqtree2_b200.phylotree as pt
    import iqtree2_b200.supertree as st
    monkeypatch.setattr(pt, "load_library", lambda: lib)
    monkeypatch.setattr(st, "load_library", lambda: lib)
    return lib


def test_nni_against_reference_golden_on_cpu(over_oracle):
    import test_gpu_nni as m
    m.test_nni_against_reference_golden(True)
    m.test_nni_against_reference_golden(False)


@pytest.mark.parametrize("kind,nni5,safe", [("dna", True, False), ("dna", False, False), ("aa", True, False)])
def test_nni_evaluation_on_cpu(over_oracle, kind, nni5, safe):
    import test_gpu_nni as m
    m.test_nni_evaluation_on_resident_clvs(kind, nni5, safe)


@pytest.mark.parametrize("nni5", [True, False])
def test_do_nni_on_cpu(over_oracle, nni5):
    import test_gpu_nni as m
    m.test_do_nni_applies_the_evaluated_move(nni5)


@pytest.mark.parametrize("fused", [True, False])
def test_config1_branch_length_optimisation_on_cpu(over_oracle, fused):
    import test_gpu_phylotree as m
    m.test_config1_fixed_tree_logl_and_pattern_lh()
    m.test_config1_branch_length_optimisation(fused)


def test_invalidation_and_scaling_on_cpu(over_oracle):
    import test_gpu_phylotree as m
    m.test_incremental_invalidation_matches_full_recompute()
    m.test_model_change_invalidates_everything()
    m.test_tree_length_scaling("dna", 1.0)


def test_partition_drivers_on_cpu(over_oracle):
    import test_gpu_supertree_unlinked as u
    u.test_unlinked_partitions(2)


@pytest.mark.parametrize("safe", [False, True])
def test_edge_linked_supertree_on_cpu(over_oracle, safe):
    import test_gpu_supertree as m
    m.test_supertree_logl_scores_and_derivatives(safe)


def test_edge_linked_supertree_optimisation_on_cpu(over_oracle):
    import test_gpu_supertree as m
    # (not test_supertree_thread_count_does_not_change_results: bit-identical sums for any number of queueing threads are a
    # property of the CUDA engine's ordered reduction; the oracle's OpenMP reductions change with the threads it is given)
    m.test_supertree_branch_length_optimisation()
    m.test_supertree_incremental_after_one_branch()


@pytest.mark.parametrize("asc_type", [1, 2])
def test_polytomies_asc_and_robust_phylo_through_the_tree_api_on_cpu(over_oracle, asc_type):
    import test_gpu_phylotree_multifurc_asc as m
    import test_gpu_robust as r
    from oracle import ref_engine
    if ref_engine.available():
        m.test_host_mirror_with_asc(asc_type)
        m.test_host_mirror_on_polytomous_tree("dna" if asc_type == 1 else "aa")
    r.test_robust_phylo_through_the_tree_api()


def test_linked_partition_nni_on_cpu(over_oracle):
    import test_gpu_supertree_nni as m
    m.test_single_partition_equals_plain_tree(True)
    m.test_single_partition_equals_plain_tree(False)
    m.test_partitioned_nni_scores_are_the_likelihoods_of_the_moved_trees()


@pytest.mark.parametrize("kind,mem", [("dna", 0.0), ("aa", 0.3)])
def test_mem_save_and_safe_kernel_through_the_tree_api_on_cpu(over_oracle, kind, mem):
    import test_gpu_phylotree as m
    m.test_mem_save_slot_recycling_on_gpu(kind, 0, mem)
    if kind == "aa":
        m.test_protein_tree_api_and_safe_kernel()
        m.test_pipelined_sweep_equals_branch_by_branch("aa")   # no device length table here: falls back to branch by branch


def test_rereading_the_tree_drops_every_cached_node_pointer(over_oracle):
    """the linked model resolves a partition's branch through its node lists (nodeById): reading a tree again must not leave
    pointers to the old nodes behind"""
    import numpy as np
    from supercases import make_supercase
    from test_gpu_supertree import make_gpu_super
    tree, parts = make_supercase(ntaxa=10, nparts=3, seed=5)
    g = make_gpu_super(tree, parts)
    ll = g.compute_likelihood()
    d = g.compute_likelihood_derv(3, 0.05)
    for _ in range(3):
        g.read_tree(tree.to_newick(fmt="%.17g"))
        assert abs(g.compute_likelihood() - ll) < 1e-9
        np.testing.assert_allclose(g.compute_likelihood_derv(3, 0.05), d, rtol=1e-10)
    g.close()


def test_refused_engine_call_rolls_the_host_state_back(over_oracle):
    """A b200lk_* call that refuses a planned traversal must not leave the tree believing those CLVs exist (the planner marks
    them computed before the call): the next evaluation has to recompute everything and give the right answer."""
    import numpy as np
    from cases import make_case
    from iqtree2_b200.phylotree import PhyloTree
    tree, model, pat, freq, inv = make_case("dna", ntaxa=12, nsites=200, seed=4)
    for mem_save in (False, True):
        t = PhyloTree()
        t.set_alignment(tree.names, pat, freq, model)
        t.set_model(model)
        t.read_tree(tree.to_newick(fmt="%.17g"))
        if mem_save:
            t.set_mem_save(True, 0.5)
        ll0 = t.compute_likelihood_branch(3)
        t.set_branch_length(7, 0.23)            # invalidates the partials that saw branch 7
        want = None
        over_oracle.b200lk_test_fail_op_call_in(1)
        with pytest.raises(Exception, match="branch_logl"):
            t.compute_likelihood_branch(3)
        ll1 = t.compute_likelihood_branch(3)    # must not reuse never-written CLVs
        t.clear_all_partial_lh()
        want = t.compute_likelihood_branch(3)
        assert abs(ll1 - want) < 1e-9 and abs(ll1 - ll0) > 1e-3
        t.close()
