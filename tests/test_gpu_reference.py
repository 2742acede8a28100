"""GPU: the CUDA engine (through the C ABI) against the REAL reference kernel — via the committed golden vectors
(tests/golden/ref_*.npz, minted by make_golden.py from oracle/_ref/libiqref.so) and, when the prebuilt libiqref.so
travelled with the snapshot, directly against it on larger seeded inputs.  Bars: scale_num bit-exact, per-pattern logL
1e-11 relative, tree logL 1e-6, optimised branch lengths 1e-6."""
import os

import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.engine import LikelihoodEngine
from oracle import ref_engine
from refmap import pytree_splits, slot_leafsets
from test_golden_reference import GOLD, GOLDEN_CASES, check_engine_against_golden, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("schedule", [1, 2])
@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_cuda_engine_against_reference_golden(name, schedule):
    tree, model, pat, freq, inv, safe, g, ops, a, b, L = load_case(name)
    e = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe, schedule=schedule)
    check_engine_against_golden(e, g, ops, a, b, L, inv)
    e.close()


@pytest.mark.parametrize("fused", [True, False])
def test_config1_golden_branch_lengths(fused):
    from iqtree2_b200.phylotree import PhyloTree
    from test_tree_logic_cpu import example_case, make_pytree
    g = np.load(os.path.join(GOLD, "ref_config1_example.npz"))
    names, pat, freq, model, newick = example_case()
    t = PhyloTree(fused_newton=fused)
    t.set_alignment(names, pat, freq, model)
    t.set_model(model)
    t.read_tree(newick)
    assert abs(t.compute_likelihood() - float(g["logl"])) < 1e-6
    np.testing.assert_allclose(t.compute_pattern_likelihood(), g["pattern_lh"], rtol=1e-11)
    assert abs(t.optimize_all_branches(iterations=100, tolerance=0.001) - float(g["opt_logl"])) < 1e-6
    # the C++ host mirror numbers its branches like tests/pytree.PyTree (same Newick reader order)
    _, _, ln = t.branches()
    np.testing.assert_allclose(ln, g["opt_lengths"], rtol=0, atol=1e-6)


needs_ref = pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so did not travel with the snapshot")


@needs_ref
@pytest.mark.parametrize("kind,safe,ntaxa,nsites", [("dna", False, 64, 6000), ("dna", True, 64, 3000), ("dna_inv", False, 40, 4000),
                                                    ("aa", False, 48, 1500), ("aa", True, 30, 700), ("aa_c4", False, 20, 600), ("aa_c4", True, 16, 300),
                                                    ("aa_fused", False, 20, 600), ("aa_c60", False, 9, 200), ("codon", True, 12, 300)])
def test_cuda_engine_against_live_reference(kind, safe, ntaxa, nsites):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=nsites, seed=11)
    r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=4)
    e = load_engine(LikelihoodEngine, tree, model, pat, freq, inv, safe=safe)
    e.set_reference_quirks(1)
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    ll_r, plh_r = r.branch_logl(0)
    ll_e = e.branch_logl(ops, a, b, L)
    assert abs(ll_e - ll_r) < 1e-6
    np.testing.assert_allclose(e.get_pattern_lh(), plh_r, rtol=1e-11)
    ref = r.directed_partials(want_partial=False)
    sets = slot_leafsets(tree, br)
    for op in ops:
        slot = int(op["dst_slot"])
        np.testing.assert_array_equal(e.get_scale_num(slot).reshape(ref[sets[slot]][1].shape), ref[sets[slot]][1])
    df_r, ddf_r = r.branch_derv(0)
    df_e, ddf_e = e.branch_derv(ops[:0], a, b, L)
    assert abs(df_e - df_r) <= 1e-8 * max(1.0, abs(df_r)) and abs(ddf_e - ddf_r) <= 1e-8 * max(1.0, abs(ddf_r))
    e.close()


@needs_ref
def test_cuda_tree_against_live_reference_optimisation():
    """a second, larger branch-length optimisation than config 1: 40 taxa x 3000 sites GTR+G4, lengths within 1e-6"""
    from iqtree2_b200.phylotree import PhyloTree
    from test_tree_logic_cpu import make_pytree
    tree, model, pat, freq, inv = make_case("dna", ntaxa=40, nsites=3000, seed=5, gaps=0.02, ambig=0.0)
    names = ["T%d" % i for i in range(tree.ntaxa)]
    tree.names = names
    newick = tree.to_newick(fmt="%.12g")
    r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
    t = PhyloTree()
    t.set_alignment(names, pat, freq, model)
    t.set_model(model)
    t.read_tree(newick)
    assert abs(t.compute_likelihood() - r.compute_likelihood()[0]) < 1e-6
    ll_r = r.optimize_all_branches(100, 0.001, 100)
    ll_t = t.optimize_all_branches(iterations=100, tolerance=0.001)
    assert abs(ll_t - ll_r) < 1e-6
    ref_len = r.branches()
    p = make_pytree(names, pat, freq, model, newick)   # only for its branch numbering / split keys
    _, _, ln = t.branches()
    for split, x in zip(pytree_splits(p), ln):
        assert abs(ref_len[split] - x) < 1e-6
