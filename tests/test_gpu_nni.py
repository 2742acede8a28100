"""GPU: NNI evaluation on resident CLVs — PhyloTree::getBestNNIForBran / doNNI / changeNNIBrans
(tree/phylotree.cpp:3947-4322) of the C++ host mirror over the CUDA engine against the Python restatement over the CPU oracle,
which tests/test_reference_parity.py pins against the compiled reference.  Moves, outer branches and nodes are matched by the
leaf sets they cover.  Bars: scores 1e-6, re-optimised branch lengths 1e-6, per-pattern logL 1e-11 relative."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200.phylotree import PhyloTree
from refmap import pytree_leafset, pytree_splits
from test_tree_logic_cpu import example_case, make_pytree

pytestmark = pytest.mark.gpu


def gpu_leafsets(g, ntaxa):
    """{(node, dad): leaf set of node's subtree seen from dad} of the C++ tree, from its branch list"""
    a, b, _ = g.branches()
    adj = {}
    for x, y in zip(a, b):
        adj.setdefault(int(x), []).append(int(y))
        adj.setdefault(int(y), []).append(int(x))
    memo = {}

    def rec(node, dad):
        if (node, dad) not in memo:
            if node < ntaxa:
                memo[(node, dad)] = frozenset([node])
            else:
                s = frozenset()
                for v in adj[node]:
                    if v != dad:
                        s |= rec(v, node)
                memo[(node, dad)] = s
        return memo[(node, dad)]

    for x, y in zip(a, b):
        rec(int(x), int(y)); rec(int(y), int(x))
    return rec, adj


def build(kind, safe=False):
    if kind == "config1":
        names, pat, freq, model, newick = example_case()
        inv = None
    else:
        tree, model, pat, freq, inv = make_case(kind, ntaxa=12, nsites=500 if kind.startswith("dna") else 200, seed=11, gaps=0.02, ambig=0.01)
        names = ["T%d" % i for i in range(tree.ntaxa)]
        tree.names = names
        newick = tree.to_newick(fmt="%.17g")
    g = PhyloTree(safe_numeric=safe)
    g.set_alignment(names, pat, freq, model)
    if inv is not None:
        g.set_pattern_invar(inv)
    g.set_model(model)
    g.read_tree(newick)
    t = make_pytree(names, pat, freq, model, newick, inv=inv)
    return g, t, names, pat, freq, model, inv


@pytest.mark.parametrize("kind,nni5,safe", [("dna", True, False), ("dna", False, False), ("aa", True, False), ("dna_inv", True, True),
                                            ("config1", True, False)])
def test_nni_evaluation_on_resident_clvs(kind, nni5, safe):
    g, t, names, pat, freq, model, inv = build(kind, safe)
    nt = len(names)
    ll0 = g.compute_likelihood()
    assert abs(t.compute_likelihood() - ll0) < 1e-6
    a, b, len0 = g.branches()
    rec, adj = gpu_leafsets(g, nt)
    spl = pytree_splits(t)
    allt = frozenset(range(nt))
    checked = 0
    for br in range(len(a)):
        na, nb = int(a[br]), int(b[br])
        if na < nt or nb < nt:
            continue
        side_b = rec(nb, na)
        split = allt - side_b if 0 in side_b else side_b
        ab, ba = t.branches[spl.index(split)]
        # pytree nodes in the orientation of the C++ tree's (end a, end b)
        node1, node2 = (ba.node, ab.node) if pytree_leafset(ab.node, ba.node) == side_b else (ab.node, ba.node)
        got = g.get_best_nni_for_bran(br, nni5=nni5, want_pattern_lh=True)
        exp = t.get_best_nni_for_bran(node1, node2, nni5=nni5)
        exp_by = {frozenset((pytree_leafset(m["swap"][0], node1), pytree_leafset(m["swap"][1], node2))): m for m in exp}
        for mv in got:
            sw = frozenset((rec(mv["swap"][0], na), rec(mv["swap"][1], nb)))
            em = exp_by[sw]
            assert abs(mv["score"] - em["score"]) < 1e-6, (br, mv["score"], em["score"])
            assert abs(mv["new_len"][0] - em["new_len"][0]) < 1e-6
            assert abs(float((mv["pattern_lh"] * freq).sum()) - mv["score"]) < 1e-6
            if nni5:
                mine = {rec(x, na if x in adj[na] else nb): mv["new_len"][1 + i] for i, x in enumerate(mv["outer"])}
                theirs = {pytree_leafset(x, node1 if node1.find(x) else node2): em["new_len"][1 + i] for i, x in enumerate(em["outer"])}
                assert set(mine) == set(theirs)
                for k in mine:
                    assert abs(mine[k] - theirs[k]) < 1e-6, (br, sorted(k), mine[k], theirs[k])
        # the tree is as before: topology, lengths
        a2, b2, len2 = g.branches()
        np.testing.assert_array_equal(a2, a); np.testing.assert_array_equal(b2, b)
        np.testing.assert_array_equal(len2, len0)
        checked += 1
    assert checked == nt - 3
    # same planning on both sides: the evaluation recomputed exactly the partials the reference's flags ask for
    assert g.num_partial_lh_computations == t.num_partial
    # nothing is stale: the resident CLVs still give the old likelihood, and so does a full recomputation
    assert abs(g.compute_likelihood() - ll0) < 1e-7
    g.clear_all_partial_lh()
    assert abs(g.compute_likelihood() - ll0) < 1e-7


@pytest.mark.parametrize("nni5", [True, False])
def test_do_nni_applies_the_evaluated_move(nni5):
    g, t, names, pat, freq, model, inv = build("dna")
    nt = len(names)
    g.compute_likelihood()
    a, b, _ = g.branches()
    inner = [br for br in range(len(a)) if a[br] >= nt and b[br] >= nt]
    br = inner[len(inner) // 2]
    moves = g.get_best_nni_for_bran(br, nni5=nni5)
    best = max(moves, key=lambda m: m["score"])
    g.do_nni(br, best["move"], new_len=best["new_len"])
    # the likelihood the evaluation promised, on the incrementally updated CLVs and after a full recomputation
    assert abs(g.compute_likelihood() - best["score"]) < 1e-6
    g.clear_all_partial_lh()
    assert abs(g.compute_likelihood() - best["score"]) < 1e-6
    # and it is the likelihood of that topology: a fresh oracle-side tree read from the Newick string agrees
    t2 = make_pytree(names, pat, freq, model, g.get_tree(), inv=inv)
    assert abs(t2.compute_likelihood() - best["score"]) < 1e-6
    # the swapped subtrees changed sides
    rec, adj = gpu_leafsets(g, nt)
    na, nb = int(a[br]), int(b[br])
    assert best["swap"][1] in adj[na] and best["swap"][0] in adj[nb]
    # without lengths the move keeps every branch length
    _, _, l1 = g.branches()
    g.do_nni(br, 0)
    _, _, l2 = g.branches()
    np.testing.assert_array_equal(l1, l2)
    assert np.isfinite(g.compute_likelihood())


@pytest.mark.parametrize("nni5", [True, False])
def test_nni_against_reference_golden(nni5):
    """the CUDA path against numbers minted from the compiled reference (tests/golden/ref_config1_nni.json): BASELINE config 1,
    every inner branch, both moves — scores and re-optimised lengths to 1e-6"""
    from test_nni_golden_cpu import check_moves, golden_cases
    g, t, names, pat, freq, model, inv = build("config1")
    nt = len(names)
    gold_tree, cases = golden_cases(nni5)
    by_split = {frozenset(c["split"]): c for c in cases}
    assert abs(g.compute_likelihood() - cases[0]["logl_before"]) < 1e-6
    a, b, _ = g.branches()
    rec, adj = gpu_leafsets(g, nt)
    allt = frozenset(range(nt))
    n = 0
    for br in range(len(a)):
        na, nb = int(a[br]), int(b[br])
        if na < nt or nb < nt:
            continue
        side = rec(nb, na)
        case = by_split[allt - side if 0 in side else side]
        # node1 = the end the reference's first swapped subtree hangs off
        reverse = not frozenset(case["moves"][0]["swap"][1]) <= side
        n1, n2 = (nb, na) if reverse else (na, nb)
        got = g.get_best_nni_for_bran(br, nni5=nni5, reverse=reverse)
        check_moves(got, case,
                    lambda mv: (rec(mv["swap"][0], n1), rec(mv["swap"][1], n2)),
                    lambda mv: [rec(x, n1 if x in adj[n1] else n2) for x in mv["outer"]], nni5)
        n += 1
    assert n == len(cases)
