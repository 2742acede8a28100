"""CPU: the partition-model restatement (tests/pysupertree.py over the scalar oracle) against the compiled reference
(oracle/_ref/libiqref.so): with a partition's absent taxa coded as rows of STATE_UNKNOWN on the FULL topology, the
partition log-likelihood equals the reference's on the INDUCED subtree with summed branch lengths — the form
PhyloSuperTree evaluates (tree/phylosupertree.cpp:697-726; linkTrees/mapTrees) — so sum_p logL_p is the reference's
super-tree log-likelihood; and the linked derivatives are the derivatives of that sum."""
import numpy as np
import pytest

from iqtree2_b200 import treesim
from oracle import ref_engine
from oracle.oracle_engine import OracleEngine
from pysupertree import PySuperTree
from supercases import induced_subtree, make_supercase


def oracle_factory(part):
    def factory(ntaxa, nslots):
        e = OracleEngine(part["model"], ntaxa, part["pat"].shape[1], nslots=nslots)
        e.set_tip_states(part["pat"]); e.set_pattern_freq(part["freq"]); e.set_pattern_invar(part["inv"])
        return e
    return factory


def make_pysuper(tree, parts):
    return PySuperTree(tree.names, tree.to_newick(fmt="%.17g"), [oracle_factory(p) for p in parts], [p["rate"] for p in parts])


def test_induced_subtree_shapes():
    tree, _ = make_supercase(ntaxa=12, nparts=1)
    keep = np.ones(12, bool); keep[[2, 5, 6]] = False
    sub, kept = induced_subtree(tree, keep)
    assert sub.ntaxa == 9 and len(sub.lengths) == 15
    assert abs(sub.lengths.sum() - tree.lengths.sum()) > 0   # pendant branches of the dropped leaves are gone
    keep[:] = True
    sub, _ = induced_subtree(tree, keep)
    assert np.isclose(sub.lengths.sum(), tree.lengths.sum())


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
def test_supertree_logl_equals_reference_on_induced_subtrees():
    tree, parts = make_supercase(ntaxa=14, nparts=6, missing=0.25)
    st = make_pysuper(tree, parts)
    total = st.compute_likelihood()
    ref_total = 0.0
    for k, p in enumerate(parts):
        assert (~p["present"]).sum() > 0 or k > 0
        sub, kept = induced_subtree(tree, p["present"])
        sub.lengths = sub.lengths * p["rate"]
        pat = p["pat"][kept, :]
        r = ref_engine.RefTree(sub, p["model"], pat, p["freq"])
        ll, _ = r.compute_likelihood()
        r.close()
        assert abs(ll - st.cur_score[k]) < 1e-7 * max(1.0, abs(ll) * 1e-3), (k, ll, st.cur_score[k])
        ref_total += ll
    assert abs(total - ref_total) < 1e-6


def test_linked_derivatives_are_derivatives_of_the_sum():
    tree, parts = make_supercase(ntaxa=10, nparts=4, seed=7)
    st = make_pysuper(tree, parts)
    st.compute_likelihood()
    for bid in (0, 3, 8):
        ab, ba = st.branches[bid]
        x = ab.length
        df, ddf = st.func_derv_at(bid, x)
        h = 1e-5
        def f(v):
            ab.length = ba.length = v
            for p in st.parts:
                p.clear_all()
            return st.branch_logl(ab, ba.node)
        fp, fm, f0 = f(x + h), f(x - h), f(x)
        assert abs((fp - fm) / (2 * h) - df) < 1e-4 * max(1.0, abs(df))
        assert abs((fp - 2 * f0 + fm) / (h * h) - ddf) < 2e-2 * max(1.0, abs(ddf))
        ab.length = ba.length = x


def test_branch_optimisation_reaches_stationary_point_of_the_sum():
    tree, parts = make_supercase(ntaxa=9, nparts=3, seed=2, sites=(80, 160))
    st = make_pysuper(tree, parts)
    ll0 = st.compute_likelihood()
    ll1 = st.optimize_all_branches(iterations=50, tolerance=1e-4)
    assert ll1 > ll0
    for bid in range(len(st.branches)):
        ab, _ = st.branches[bid]
        if 1e-5 < ab.length < 9.0:
            df, _ = st.func_derv_at(bid, ab.length)
            assert abs(df) < 0.5, (bid, df)   # Newton tolerance is on the length (1e-6), |df| ~ |ddf| * 1e-6 plus sweep coupling
