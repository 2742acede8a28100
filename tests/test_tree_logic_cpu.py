"""CPU: BASELINE config 1 (example.phy, 17 taxa x 1998 sites, GTR+G4) through the tree-level logic over the oracle:
fixed-tree logL against the brute-force pruning, and Newton-Raphson branch-length optimisation reaching a stationary
point (df ~ 0 on every branch, logL not decreasing)."""
import os

import numpy as np

import bruteforce as bf
from iqtree2_b200 import alignment, models, treesim
from oracle.oracle_engine import OracleEngine
from pytree import PyTree

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def example_case():
    names, seqs = alignment.read_phylip(os.path.join(GOLD, "example.phy"))
    pat, freq = alignment.compress(alignment.encode(seqs, "DNA"))
    model = models.make_gtr_gamma(rates6=(1.2, 4.1, 0.9, 1.1, 4.6, 1.0), pi=(0.36, 0.26, 0.13, 0.25), alpha=0.6)
    newick = open(os.path.join(GOLD, "example_nj.nwk")).read().strip()
    return names, pat, freq, model, newick


def make_pytree(names, pat, freq, model, newick, inv=None):
    def factory(ntaxa, nslots):
        e = OracleEngine(model, ntaxa, pat.shape[1], nslots=nslots)
        e.set_tip_states(pat); e.set_pattern_freq(freq)
        e.set_pattern_invar(np.zeros(pat.shape[1]) if inv is None else inv)
        return e
    return PyTree(factory, names, newick)


def pytree_to_arrays(t):
    """PyTree -> treesim.Tree (for the brute-force checker)."""
    edges, lens = [], []
    for ab, ba in t.branches:
        edges.append([ba.node.id, ab.node.id]); lens.append(ab.length)
    return treesim.Tree(len(t.names), np.array(edges), np.array(lens), t.names)


def test_example_phy_fixed_tree_logl():
    names, pat, freq, model, newick = example_case()
    assert pat.shape == (17, 1152) and freq.sum() == 1998
    t = make_pytree(names, pat, freq, model, newick)
    ll = t.compute_likelihood()
    ref = bf.tree_loglik(pytree_to_arrays(t), model, pat, freq, branch=0)
    assert abs(ll - ref) < 1e-6, (ll, ref)
    # the same value from every branch
    for ab, ba in t.branches[::5]:
        assert abs(t.branch_logl(ab, ba.node) - ll) < 1e-7


def test_example_phy_branch_optimisation_reaches_stationary_point():
    names, pat, freq, model, newick = example_case()
    t = make_pytree(names, pat, freq, model, newick)
    ll0 = t.compute_likelihood()
    ll1 = t.optimize_all_branches(iterations=100, tolerance=0.001)
    assert ll1 > ll0
    # independent check of the optimum: brute-force logL at the optimised lengths, and df ~ 0 per branch
    ref = bf.tree_loglik(pytree_to_arrays(t), model, pat, freq, branch=0)
    assert abs(ll1 - ref) < 1e-6
    for ab, ba in t.branches:
        if ab.length <= 2e-6:
            continue
        t.theta_computed = False
        df, ddf = t.branch_derv(ab, ba.node)
        assert abs(df) < 0.5 and ddf < 0, (ab.id, ab.length, df, ddf)
    # partial updates stay incremental: far fewer than a full traversal per Newton step
    assert t.num_partial < 60 * 15
