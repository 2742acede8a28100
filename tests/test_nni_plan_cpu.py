"""CPU: PhyloTree::doNNI of the C++ host mirror (neighbour swap, slot re-orientation before the swap, partial_lh_computed flags
cleared on both sides; tree/phylotree.cpp:3947-4036) in plan-only mode.  The op lists planned after each move are executed by the
CPU oracle on the slots the earlier lists filled: the incremental result must be the likelihood of the new topology, which a
fresh tree read from the mirror's Newick string gives independently."""
import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200.phylotree import PhyloTree
from oracle.oracle_engine import OracleEngine
from test_tree_logic_cpu import make_pytree


@pytest.mark.parametrize("kind,ntaxa", [("dna", 16), ("aa", 12)])
def test_do_nni_plans_incremental_updates(kind, ntaxa):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=150, seed=8)
    names = ["T%d" % i for i in range(ntaxa)]
    tree.names = names
    t = PhyloTree()
    t.set_plan_only(True)
    t.set_alignment(names, pat, freq, model); t.set_model(model)
    t.read_tree(tree.to_newick(fmt="%.17g"))
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, nslots=ntaxa)      # T-2 slots + the two spares
    a, b, _ = t.branches()
    inner = [br for br in range(len(a)) if a[br] >= ntaxa and b[br] >= ntaxa]
    ops, ea, eb, L = t.plan_branch(0)
    assert len(ops) == ntaxa - 2
    ll = o.branch_logl(ops, ea, eb, L)
    assert abs(ll - make_pytree(names, pat, freq, model, t.get_tree(), inv=inv).compute_likelihood()) < 1e-7
    rng = np.random.default_rng(3)
    seen = {t.get_tree()}
    total = 0
    for step in range(12):
        br = int(rng.choice(inner))
        t.do_nni(br, int(rng.integers(2)), reverse=bool(rng.integers(2)))
        nwk = t.get_tree()
        seen.add(nwk)
        at = int(rng.integers(len(a)))
        ops, ea, eb, L = t.plan_branch(at)
        assert len(ops) <= ntaxa - 2
        total += len(ops)
        ll = o.branch_logl(ops, ea, eb, L)
        fresh = make_pytree(names, pat, freq, model, nwk, inv=inv).compute_likelihood()
        assert abs(ll - fresh) < 1e-7 * max(1.0, abs(fresh)), (step, br, ll, fresh)
    assert len(seen) > 6                                       # the moves really changed the topology
    assert total < 12 * (ntaxa - 2) * 0.8                      # an NNI invalidates some partials, not the whole tree
    o.close(); t.close()
