"""GPU: the host mirror (b200pt_*) on a POLYTOMOUS tree and with +ASC against the compiled reference reading the same Newick:
tree logL, and optimizeAllBranches — same traversal planning over multifurcating nodes, same Newton iterations — with the
optimised branch lengths compared split by split (1e-6)."""
import re

import numpy as np
import pytest

from iqtree2_b200.phylotree import PhyloTree
from multifurc import newick
from oracle import ref_engine
from test_asc_cpu import asc_case, holder_case
from test_multifurcation_cpu import polytomy_case

pytestmark = pytest.mark.gpu


def splits_of(nwk, names):
    """{frozenset(alignment rows on the side without row 0): length} of every branch of a Newick string"""
    row = {n: i for i, n in enumerate(names)}
    tok = re.findall(r"[(),;]|:[^(),;:]+|[^(),;:]+", nwk)
    out, stack, i = {}, [], 0
    cur = None
    all_rows = frozenset(range(len(names)))

    def put(s, ln):
        s = frozenset(s)
        if 0 in s:
            s = all_rows - s
        if s and s != all_rows:
            out[s] = out.get(s, 0.0) + ln

    while i < len(tok):
        t = tok[i]
        if t == "(":
            stack.append([])
        elif t == ")":
            kids = stack.pop()
            cur = set().union(*kids)
            if stack:
                stack[-1].append(cur)
        elif t == "," or t == ";":
            pass
        elif t.startswith(":"):
            put(cur, float(t[1:]))
        else:
            cur = {row[t]}
            stack[-1].append(cur)
        i += 1
    return out


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind", ["dna", "aa"])
def test_host_mirror_on_polytomous_tree(kind):
    tree, model, pat, freq, inv, adj = polytomy_case(kind, False, ntaxa=24, nsites=500, ncollapse=4, seed=13)
    nwk = newick(adj, tree.names)
    g = PhyloTree()
    g.set_alignment(tree.names, pat, freq, model); g.set_pattern_invar(inv); g.set_model(model); g.read_tree(nwk)
    r = ref_engine.RefTree(None, model, pat, freq, newick=nwk, names=tree.names)
    ll_g, (ll_r, _) = g.compute_likelihood(), r.compute_likelihood()
    assert abs(ll_g - ll_r) < 1e-6, (ll_g, ll_r)
    og, orf = g.optimize_all_branches(iterations=100, tolerance=0.001), r.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(og - orf) < 1e-6, (og, orf)
    sg, sr = splits_of(g.get_tree(), tree.names), r.branches()
    assert set(sg) == set(sr)
    for k in sr:
        assert abs(sg[k] - sr[k]) < 1e-6, (sorted(k), sg[k], sr[k])
    r.close(); g.close()


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("asc_type", [1, 2])
def test_host_mirror_with_asc(asc_type):
    """computeLikelihood and optimizeAllBranches with +ASC (Lewis) and +ASC_MIS (Holder) against the compiled reference"""
    if asc_type == 1:
        tree, model, pat, freq, unobs, pat_all, freq_all = asc_case("bin", ntaxa=14, nsites=500)
    else:
        tree, model, pat, freq, unobs, pat_all, freq_all = holder_case("bin", ntaxa=14, nsites=500)
    nobs, U = pat.shape[1], unobs.shape[0]
    g = PhyloTree(safe_numeric=True)
    g.set_alignment(tree.names, pat_all, freq_all, model); g.set_model(model)
    g.set_ascertainment(asc_type, U, nobs)          # the harness alignment reports one site per pattern
    nwk = tree.to_newick(fmt="%.17g")
    g.read_tree(nwk)
    ref_engine.set_asc_next(asc_type, unobs)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=True)
    ll_g, (ll_r, _) = g.compute_likelihood(), r.compute_likelihood()
    assert abs(ll_g - ll_r) < 1e-6, (ll_g, ll_r)
    og, orf = g.optimize_all_branches(iterations=100, tolerance=0.001), r.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(og - orf) < 1e-6, (og, orf)
    sg, sr = splits_of(g.get_tree(), tree.names), r.branches()
    for k in sr:
        assert abs(sg[k] - sr[k]) < 1e-6, (sorted(k), sg[k], sr[k])
    r.close(); g.close()
