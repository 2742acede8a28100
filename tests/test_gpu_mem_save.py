"""GPU: the reference's -mem mode (LM_MEM_SAVE, tree/phylotree.cpp:999-1021, tree/memslot.cpp) on the CUDA engine: op lists planned
by the host mirror with a bounded number of recycled CLV slots — a slot serves several nodes inside one traversal, and later
traversals start from whatever survived — must give what the CPU oracle gives on the same op lists, call after call."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200.engine import LikelihoodEngine
from iqtree2_b200.phylotree import PhyloTree
from oracle.oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,safe", [("dna", False), ("aa", False), ("aa_c4", False), ("aa_c4", True), ("dna_g8", False), ("aa_fused", False)])
@pytest.mark.parametrize("frac", [0.25, 0.6])
def test_recycled_slots_over_several_traversals(kind, safe, frac):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=36, nsites=500 if kind.startswith("aa") else 900, seed=17)
    ntaxa = tree.ntaxa
    pt = PhyloTree()
    pt.set_plan_only(True)
    pt.set_alignment(tree.names, np.zeros((ntaxa, 1), dtype=np.uint8), np.ones(1), model)
    pt.set_model(model)
    pt.read_tree(tree.to_newick(fmt="%.17g"))
    pt.set_mem_save(True, frac)
    nslots = pt.max_lh_slots + 2
    assert nslots - 2 < ntaxa - 2
    g = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe, nslots=nslots)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe, nslots=nslots)
    pa, pb, _ = pt.branches()
    first = [k for k in range(len(pa)) if min(pa[k], pb[k]) == 0][0]
    inner = [k for k in range(len(pa)) if min(pa[k], pb[k]) >= ntaxa]
    ll0 = None
    for br in [first] + inner[::4] + [first]:
        ops, a, b, L = pt.plan_branch(br)
        ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
        assert abs(ll_g - ll_o) < 1e-6 + 1e-12 * abs(ll_o), (br, len(ops), ll_g, ll_o)
        np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11)
        for op in ops:
            s = int(op["dst_slot"])
            assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), (br, s)
        if ll0 is None:
            ll0 = ll_g
        assert abs(ll_g - ll0) < 1e-7 * abs(ll0)   # pulley principle
        g.theta_reset(); o.theta_reset()   # another branch: the caller's theta_computed = false (tree/phylotree.cpp:2608)
        dg, do = g.branch_derv(ops[:0], a, b, L), o.branch_derv(ops[:0], a, b, L)
        np.testing.assert_allclose(dg, do, rtol=1e-9, atol=1e-9)
    g.close()
    pt.close()
