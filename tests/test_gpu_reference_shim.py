"""GPU: the drop-in, end to end.  The UNMODIFIED reference's PhyloTree (compiled from /root/reference by oracle/build_ref.sh) with
its five kernel pointers bound to the product's C ABI by oracle/ref_b200_shim.cpp — the binding INTEGRATION.md describes, compiled
for real — against the same reference running its own AVX/FMA kernel.  Everything above the kernels is the reference's code in
both arms: computeTraversalInfo, reorientPartialLh, optimizeOneBranch / minimizeNewton, optimizeAllBranches, getBestNNIForBran (on
nni_partial_lh = the engine's spare slots), optimizeTreeLengthScaling, computePatternLhCat.  Bars: tree logL 1e-6, per-pattern logL
1e-11 relative, branch lengths 1e-6."""
import numpy as np
import pytest

from cases import make_case
from oracle import ref_engine
from refmap import branch_splits
from test_tree_logic_cpu import example_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (ref_engine.available() and ref_engine.available_b200()), reason="oracle/_ref libraries not built")]


def pair(kind):
    if kind == "config1":
        names, pat, freq, model, newick = example_case()
        mk = lambda b200: ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names, b200=b200)
        return mk(False), mk(True), None, len(names)
    tree, model, pat, freq, inv = make_case(kind, ntaxa=14, nsites=600 if kind.startswith("dna") else 250, seed=21, gaps=0.03, ambig=0.01)
    safe = model.nstates not in (4, 20)
    mk = lambda b200: ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, b200=b200)
    return mk(False), mk(True), tree, tree.ntaxa


@pytest.mark.parametrize("kind", ["config1", "dna_inv", "aa", "aa_c4", "bin"])
def test_reference_control_flow_over_the_cuda_kernels(kind):
    cpu, gpu, tree, nt = pair(kind)
    ll_c, plh_c = cpu.compute_likelihood()
    ll_g, plh_g = gpu.compute_likelihood()
    assert abs(ll_c - ll_g) < 1e-6, (ll_c, ll_g)
    np.testing.assert_allclose(plh_g, plh_c, rtol=1e-11)
    calls, launches = gpu.b200_kernel_calls()
    assert calls >= 1 and launches >= 1          # the bound arm really went through the C ABI
    # per-category likelihoods through the reference's own computePatternLhCat(wsl)
    for wsl in (2, 4):
        a, b = cpu.pattern_lh_cat_wsl(wsl), gpu.pattern_lh_cat_wsl(wsl)
        scale = np.abs(a).max(axis=1, keepdims=True)
        np.testing.assert_allclose(b / scale, a / scale, rtol=0, atol=1e-11)
    # the reference's optimizeAllBranches (its minimizeNewton over the engine's derivatives)
    oc, og = cpu.optimize_all_branches(100, 0.001, 100), gpu.optimize_all_branches(100, 0.001, 100)
    assert abs(oc - og) < 1e-6, (oc, og)
    bc, bg = cpu.branches(), gpu.branches()
    assert set(bc) == set(bg)
    for k in bc:
        assert abs(bc[k] - bg[k]) < 1e-6, (sorted(k), bc[k], bg[k])
    # the reference's getBestNNIForBran on every inner branch: the copies it makes point into nni_partial_lh = the spare slots
    inner = [s for s in bc if 1 < len(s) < nt - 1]
    assert len(inner) == nt - 3
    for s in inner:
        mc, mg = cpu.best_nni(s), gpu.best_nni(s)
        for x, y in zip(mc, mg):
            assert x["swap"] == y["swap"]
            assert abs(x["score"] - y["score"]) < 1e-6, (sorted(s), x["score"], y["score"])
            np.testing.assert_allclose(y["new_len"], x["new_len"], rtol=0, atol=1e-6)
    assert abs(gpu.compute_likelihood()[0] - og) < 1e-6          # the tree is untouched after the evaluations
    # Brent over full-tree evaluations
    (lc, sc), (lg, sg) = cpu.optimize_tree_length_scaling(0.01, 1.0, 100.0, 0.001), gpu.optimize_tree_length_scaling(0.01, 1.0, 100.0, 0.001)
    assert abs(sc - sg) < 1e-6 and abs(lc - lg) < 1e-6
    calls2, launches2 = gpu.b200_kernel_calls()
    assert calls2 > calls + 100 and launches2 > launches + 100
    cpu.close(); gpu.close()


@pytest.mark.parametrize("kind", ["config1", "aa"])
def test_reference_with_the_fused_newton_hook(kind):
    """the optional hook: the reference's optimizeOneBranch with its minimizeNewton loop replaced by one b200lk_newton_branch call;
    its optimizeAllBranches and getBestNNIForBran then give the lengths and scores of the plain binding and of the CPU kernel"""
    cpu, gpu, tree, nt = pair(kind)
    gpu.b200_set_fused_newton(True)
    assert abs(cpu.compute_likelihood()[0] - gpu.compute_likelihood()[0]) < 1e-6
    calls0, _ = gpu.b200_kernel_calls()
    oc, og = cpu.optimize_all_branches(100, 0.001, 100), gpu.optimize_all_branches(100, 0.001, 100)
    assert abs(oc - og) < 1e-6, (oc, og)
    bc, bg = cpu.branches(), gpu.branches()
    for k in bc:
        assert abs(bc[k] - bg[k]) < 1e-6, (sorted(k), bc[k], bg[k])
    for s in [s for s in bc if 1 < len(s) < nt - 1][::2]:
        for x, y in zip(cpu.best_nni(s), gpu.best_nni(s)):
            assert abs(x["score"] - y["score"]) < 1e-6
            np.testing.assert_allclose(y["new_len"], x["new_len"], rtol=0, atol=1e-6)
    cpu.close(); gpu.close()


@pytest.mark.parametrize("asc_type", [1, 2])
def test_reference_with_asc_over_the_cuda_kernels(asc_type):
    """+ASC (Lewis) and +ASC_MIS (Holder) through the binding: ModelFactory::unobserved_ptns become the engine's last patterns"""
    from test_asc_cpu import asc_case, holder_case
    tree, model, pat, freq, unobs, pat_all, freq_all = (asc_case if asc_type == 1 else holder_case)("bin", ntaxa=14, nsites=500)
    ref_engine.set_asc_next(asc_type, unobs)
    cpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=True)
    ref_engine.set_asc_next(asc_type, unobs, b200=True)
    gpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=True, b200=True)
    (ll_c, plh_c), (ll_g, plh_g) = cpu.compute_likelihood(), gpu.compute_likelihood()
    assert abs(ll_c - ll_g) < 1e-6, (ll_c, ll_g)
    np.testing.assert_allclose(plh_g, plh_c, rtol=1e-11)
    oc, og = cpu.optimize_all_branches(100, 0.001, 100), gpu.optimize_all_branches(100, 0.001, 100)
    assert abs(oc - og) < 1e-6, (oc, og)
    bc, bg = cpu.branches(), gpu.branches()
    for k in bc:
        assert abs(bc[k] - bg[k]) < 1e-6
    cpu.close(); gpu.close()


@pytest.mark.parametrize("nstates,ncat,safe", [(4, 4, False), (20, 2, True)])
def test_reference_rooted_tree_with_a_nonreversible_model_over_the_cuda_kernels(nstates, ncat, safe):
    """the reference's own PhyloTree, ROOTED (root leaf, branch directions from computeBranchDirection), with a non-reversible model:
    its computeTraversalInfo / optimizeOneBranch / optimizeAllBranches over the product's non-reversible kernels
    (b200lk_set_model_nonrev; child kinds and the transposition flag marshalled from isRootLeaf / PhyloNeighbor::direction) against
    the same tree over the reference's computeNonrev* kernels (tree/phylokernelnonrev.h)"""
    from nonrev_cases import make_nonrev_case
    tree, model, pat, freq, invar = make_nonrev_case(nstates, ncat, 1, ntaxa=12, nptn=300, seed=37)
    cpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=2)
    gpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, b200=True)
    ll_c, plh_c = cpu.compute_likelihood()
    ll_g, plh_g = gpu.compute_likelihood()
    assert abs(ll_c - ll_g) < 1e-6, (ll_c, ll_g)
    np.testing.assert_allclose(plh_g, plh_c, rtol=1e-11)
    for leaf in (0, 5, pat.shape[0]):          # pendant branches and the branch of the root leaf
        (lc, pc), (lg, pg) = cpu.branch_logl(leaf), gpu.branch_logl(leaf)
        assert abs(lc - lg) < 1e-6
        np.testing.assert_allclose(pg, pc, rtol=1e-11)
        if leaf != pat.shape[0]:
            np.testing.assert_allclose(gpu.branch_derv(leaf), cpu.branch_derv(leaf), rtol=1e-9)
    oc, og = cpu.optimize_all_branches(5, 0.001, 100), gpu.optimize_all_branches(5, 0.001, 100)
    assert abs(oc - og) < 1e-6, (oc, og)
    assert og > ll_g
    bc, bg = cpu.branches(), gpu.branches()
    assert set(bc) == set(bg)
    for k in bc:
        assert abs(bc[k] - bg[k]) < 1e-6, (sorted(k), bc[k], bg[k])
    calls, launches = gpu.b200_kernel_calls()
    assert calls > 20 and launches > 20
    cpu.close(); gpu.close()


@pytest.mark.parametrize("nstates,ncat,safe", [(4, 4, False), (20, 4, True)])
def test_reference_tree_with_a_site_specific_model_over_the_cuda_kernels(nstates, ncat, safe):
    """ModelSet / PMSF: the binding de-interleaves the per-pattern eigensystems the reference's kernels read (V-interleaved,
    tree/phylotreesse.cpp:271-281) into b200lk_set_site_model; the reference's own optimizeAllBranches then runs over the site kernels"""
    from test_site_model_cpu import make_site_case
    tree, model, pat, freq, invar = make_site_case(nstates, ncat, ntaxa=12, nptn=300, seed=43)
    cpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=2)
    gpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, b200=True)
    for leaf in (0, 5, 9):
        (lc, pc), (lg, pg) = cpu.branch_logl(leaf), gpu.branch_logl(leaf)
        assert abs(lc - lg) < 1e-9 * abs(lc)
        np.testing.assert_allclose(pg, pc, rtol=1e-11)
        np.testing.assert_allclose(gpu.branch_derv(leaf), cpu.branch_derv(leaf), rtol=1e-9)
        assert abs(gpu.logl_from_buffer() - cpu.logl_from_buffer()) < 1e-9 * abs(lc)
    oc, og = cpu.optimize_all_branches(5, 0.001, 100), gpu.optimize_all_branches(5, 0.001, 100)
    assert abs(oc - og) < 1e-6, (oc, og)
    bc, bg = cpu.branches(), gpu.branches()
    for k in bc:
        assert abs(bc[k] - bg[k]) < 1e-6
    cpu.close(); gpu.close()


@pytest.mark.parametrize("kind,ncat,safe", [("dna", 4, False), ("aa", 3, True)])
def test_reference_mixlen_tree_over_the_cuda_kernels(kind, ncat, safe):
    """a tree with mixed branch lengths (PhyloNodeMixlen nodes, isMixlen()): the binding sends the [branch id][class] table
    (b200lk_set_mixlen), passes Neighbor::id where the kernels take a length, and serves computeLikelihoodDerv's gradient / Hessian form
    and computeLikelihoodDervMixlen from the CUDA kernels"""
    from test_mixlen_cpu import make_mixlen_case
    tree, model, pat, freq, invar = make_mixlen_case(kind, ncat, ntaxa=12, nsites=400, seed=47)
    cpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, nthreads=2)
    gpu = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe, b200=True)
    for leaf in (0, 5, 9):
        (lc, pc), (lg, pg) = cpu.branch_logl(leaf), gpu.branch_logl(leaf)
        assert abs(lc - lg) < 1e-9 * abs(lc)
        np.testing.assert_allclose(pg, pc, rtol=1e-11)
        g_c, H_c, l_c = cpu.branch_derv_mixlen(leaf)
        g_g, H_g, l_g = gpu.branch_derv_mixlen(leaf)
        np.testing.assert_allclose(g_g, g_c, rtol=1e-9, atol=1e-7)
        np.testing.assert_allclose(H_g, H_c, rtol=1e-9, atol=1e-6)
        assert abs(l_g - l_c) < 1e-9 * abs(l_c)
        assert abs(gpu.logl_from_buffer() - cpu.logl_from_buffer()) < 1e-9 * abs(lc)
        for cur in range(ncat):
            np.testing.assert_allclose(gpu.branch_derv_mixlen_cur(leaf, cur), cpu.branch_derv_mixlen_cur(leaf, cur), rtol=1e-9, atol=1e-7)
    calls, launches = gpu.b200_kernel_calls()
    assert calls > 10 and launches > 10
    cpu.close(); gpu.close()
