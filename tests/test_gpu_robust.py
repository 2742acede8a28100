"""GPU: the robust-phylo variants of computeLikelihoodBranch (--robust-phy keep, --robust-median; tree/phylokernelnew.h:3134-3173)
through the C ABI (b200lk_set_robust_phylo) against the CPU oracle, which tests/test_robust_cpu.py pins against the compiled
reference.  Compressed patterns with frequencies > 1: the device path ranks sites without expanding them."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import EngineError, LikelihoodEngine
from oracle.oracle_engine import OracleEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,nsites,keep,median", [("dna", 2001, 0.9, False), ("dna", 2000, 0.5, False), ("aa", 300, 0.73, False),
                                                     ("dna", 2001, 1.0, True), ("dna", 2000, 1.0, True), ("bin", 1000, 1.0, True)])
def test_robust_phylo(kind, nsites, keep, median):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=14, nsites=nsites, seed=5, gaps=0.03, ambig=0.0)
    safe = model.nstates not in (4, 20)
    assert freq.max() > 1 and int(freq.sum()) == nsites
    n = pat.shape[1]
    g = LikelihoodEngine(model, tree.ntaxa, n, safe_numeric=safe)
    o = OracleEngine(model, tree.ntaxa, n, safe_numeric=safe)
    for e in (g, o):
        e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(inv)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll_plain = g.branch_logl(ops, a, b, L)
    assert abs(ll_plain - o.branch_logl(ops, a, b, L)) < 1e-6
    for e in (g, o):
        e.set_robust_phylo(keep, median)
    ll_g, ll_o = g.branch_logl(ops[:0], a, b, L), o.branch_logl(ops[:0], a, b, L)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11)
    assert ll_g > ll_plain if not median else ll_g > ll_plain / nsites * 4
    # computeLikelihoodFromBuffer has no robust branch in the reference either
    g.branch_derv(ops[:0], a, b, L)
    assert abs(g.logl_from_buffer(L) - ll_plain) < 1e-6
    # switched off again
    g.set_robust_phylo(1.0, False)
    assert abs(g.branch_logl(ops[:0], a, b, L) - ll_plain) < 1e-9
    # not together with +ASC (outError in the reference, :3136-3138)
    g.set_robust_phylo(keep, median)
    with pytest.raises(EngineError):
        g.set_ascertainment(1, 2, nsites)
    g.close()


def test_robust_phylo_through_the_tree_api():
    from iqtree2_b200.phylotree import PhyloTree
    tree, model, pat, freq, inv = make_case("dna", ntaxa=14, nsites=1500, seed=6, gaps=0.02, ambig=0.0)
    g = PhyloTree()
    g.set_alignment(tree.names, pat, freq, model); g.set_pattern_invar(inv); g.set_model(model)
    g.read_tree(tree.to_newick(fmt="%.17g"))
    g.set_robust_phylo(0.8, False)          # before the engine exists: applied by initializeAllPartialLh
    ll = g.compute_likelihood()
    plh = g.compute_pattern_likelihood()
    sites = np.sort(np.repeat(plh, freq.astype(int)))
    assert abs(ll - sites[int((1.0 - 0.8) * len(sites)):].sum()) < 1e-6
    g.set_robust_phylo(1.0, True)           # on the live engine
    g.clear_all_partial_lh()
    m = len(sites)
    assert abs(g.compute_likelihood() - (sites[m // 2] if m & 1 else 0.5 * (sites[m // 2] + sites[m // 2 - 1]))) < 1e-9
    g.set_robust_phylo(1.0, False)
    g.clear_all_partial_lh()
    assert abs(g.compute_likelihood() - float((plh * freq).sum())) < 1e-6
