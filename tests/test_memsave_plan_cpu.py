"""CPU: the C++ host mirror's traversal planner (b200::PhyloTree::computeTraversalInfo, LM_PER_NODE slot re-orientation and
the -mem slot table b200::MemSlotVector = tree/memslot.cpp) in plan-only mode; the op lists it emits are executed by
the CPU oracle, so slot recycling is checked for correctness without a GPU:

  * every branch of the tree, visited in a scrambled order, gives the same logL as a from-scratch full-memory evaluation;
  * LM_MEM_SAVE never names a slot >= max_lh_slots and works down to the reference's minimum log2(T)+1 slots;
  * moving to a neighbouring branch recomputes only a few partials in LM_PER_NODE and more (but still correct) with few slots.
"""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.phylotree import PhyloTree
from oracle.oracle_engine import OracleEngine


def make_planner(tree, model, pat, freq, mem_save=None):
    t = PhyloTree()
    t.set_plan_only(True)
    t.set_alignment(tree.names, pat, freq, model)
    t.set_model(model)
    t.read_tree(tree.to_newick())
    if mem_save is not None:
        t.set_mem_save(True, mem_save)
    return t


@pytest.mark.parametrize("kind,ntaxa,tree_kind", [("dna", 24, "yule"), ("dna", 33, "caterpillar"), ("aa", 16, "yule")])
@pytest.mark.parametrize("mem", [None, 0.0, 0.5])
def test_planner_op_lists_on_oracle(kind, ntaxa, tree_kind, mem):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=120, tree_kind=tree_kind)
    # truth: full traversal with the numpy planner, one slot per node
    o_full = load_engine(OracleEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 0)
    ll_true = o_full.branch_logl(ops, a, b, L)

    t = make_planner(tree, model, pat, freq, mem)
    nslots = t.max_lh_slots
    if mem is None:
        assert nslots == ntaxa - 2
    elif mem == 0.0:
        assert nslots == int(np.log2(ntaxa)) + 1
    else:
        assert nslots == max(int(np.floor(0.5 * (ntaxa - 2))), int(np.log2(ntaxa)) + 1)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, nslots=nslots + 2)
    nbr = len(t.branches()[0])
    order = np.random.default_rng(5).permutation(nbr)
    total_ops = 0
    for br in list(order) + list(order[:5]):
        ops, a, b, L = t.plan_branch(int(br))
        for op in ops:
            assert 0 <= op["dst_slot"] < nslots
            for side in ("left", "right"):
                if op[side]["kind"] == plan.KIND_SLOT:
                    assert 0 <= op[side]["id"] < nslots and op[side]["id"] != op["dst_slot"]
        total_ops += len(ops)
        ll = o.branch_logl(ops, a, b, L)
        assert abs(ll - ll_true) < 1e-8 * abs(ll_true), (br, ll, ll_true)
    assert total_ops == t.num_partial_lh_computations
    if mem is None:
        # incremental: far fewer updates than a full traversal per visited branch
        assert total_ops < (nbr + 5) * (ntaxa - 2) / 2
    o.close(); o_full.close(); t.close()


def test_mem_save_neighbouring_branches_are_incremental():
    tree, model, pat, freq, inv = make_case("dna", ntaxa=40, nsites=60)
    t = make_planner(tree, model, pat, freq, 0.6)
    ops, *_ = t.plan_branch(0)
    assert len(ops) == 38
    # re-planning the same branch needs nothing: both ends are still resident
    ops, *_ = t.plan_branch(0)
    assert len(ops) == 0
    t.close()
