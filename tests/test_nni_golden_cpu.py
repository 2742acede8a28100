"""CPU: the Python restatement of getBestNNIForBran over the oracle against golden vectors minted from the compiled reference
(tests/golden/ref_config1_nni.json, tests/golden/make_nni_golden.py): BASELINE config 1, every inner branch, nni5 and
central-branch-only.  The file travels with the repo; tests/test_gpu_nni.py checks the CUDA path against the same numbers."""
import json
import os

import pytest

from refmap import pytree_leafset, pytree_splits
from test_tree_logic_cpu import example_case, make_pytree

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_config1_nni.json")


def golden_cases(nni5):
    g = json.load(open(GOLD))
    return g["tree"], [c for c in g["cases"] if c["nni5"] == nni5]


def check_moves(got, case, leafset_of_swap, leafset_of_outer, nni5):
    """got: moves with score / new_len / swap / outer in the caller's node terms; leafset_*: node -> frozenset of taxa"""
    ref_by = {frozenset((frozenset(m["swap"][0]), frozenset(m["swap"][1]))): m for m in case["moves"]}
    assert len(ref_by) == 2
    for mv in got:
        rm = ref_by[frozenset(leafset_of_swap(mv))]
        assert abs(mv["score"] - rm["score"]) < 1e-6, (mv["score"], rm["score"])            # tree logL bar
        assert abs(mv["new_len"][0] - rm["new_len"][0]) < 1e-6                               # branch-length bar
        if nni5:
            mine = dict(zip(leafset_of_outer(mv), mv["new_len"][1:]))
            theirs = {frozenset(k): v for k, v in zip(rm["outer"], rm["new_len"][1:])}
            assert set(mine) == set(theirs)
            for k in mine:
                assert abs(mine[k] - theirs[k]) < 1e-6


@pytest.mark.parametrize("nni5", [True, False])
def test_pytree_nni_against_reference_golden(nni5):
    names, pat, freq, model, newick = example_case()
    gold_tree, cases = golden_cases(nni5)
    assert gold_tree == newick
    t = make_pytree(names, pat, freq, model, newick)
    assert abs(t.compute_likelihood() - cases[0]["logl_before"]) < 1e-6
    spl = pytree_splits(t)
    by_split = {frozenset(c["split"]): c for c in cases}
    n = 0
    for ab, ba in t.branches:
        if ab.node.is_leaf() or ba.node.is_leaf():
            continue
        case = by_split[spl[ab.id]]
        # which end the reference took as node1 decides the order of the five branch optimisations: its first swapped subtree
        # hangs off node1, its second off node2
        node1, node2 = ba.node, ab.node
        if not frozenset(case["moves"][0]["swap"][1]) <= pytree_leafset(node2, node1):
            node1, node2 = node2, node1
        got = t.get_best_nni_for_bran(node1, node2, nni5=nni5)
        check_moves(got, case,
                    lambda mv: (pytree_leafset(mv["swap"][0], node1), pytree_leafset(mv["swap"][1], node2)),
                    lambda mv: [pytree_leafset(x, node1 if node1.find(x) else node2) for x in mv["outer"]], nni5)
        n += 1
    assert n == len(cases) == len(names) - 3
