"""GPU: BASELINE.json configs 2 and 3 at their FULL sizes (1000 taxa x 1 000 000 patterns GTR+G4: 131 GB of resident CLVs;
200 x 100 000 LG+G4), checked through properties that do not need a full-size CPU run:
  * patterns are independent, so the first `SLICE` patterns of the full-size run must reproduce the CPU oracle's run on just
    those patterns: per-pattern log-likelihood (1e-11 relative) and the scale counts of the evaluation-branch CLV (bit-exact);
  * sum(ptn_freq * _pattern_lh) is the returned logL;
  * the pulley principle: another evaluation branch gives the same logL (needs every CLV of the other orientation);
  * (df, ddf) are the derivatives of logL along the branch (central differences of two more evaluations)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (the workload definitions and input generator of the benchmark)
from iqtree2_b200 import plan  # noqa: E402
from iqtree2_b200.engine import LikelihoodEngine  # noqa: E402
from oracle.oracle_engine import OracleEngine  # noqa: E402

pytestmark = pytest.mark.gpu
SLICE = 1536


@pytest.mark.parametrize("workload", ["aa_200x100k_lg_g4", "dna_1000x1M_gtr_g4"])
def test_baseline_config_at_full_size(workload):
    import torch
    ntaxa, nptn, mk, _ = bench.WORKLOADS[workload]
    model = mk()
    need = (ntaxa + 2) * ((nptn + 127) // 128 * 128) * (8 * model.nstates * model.ncat_mix + 2) * 1.03
    if torch.cuda.get_device_properties(0).total_memory < need:
        pytest.skip("device memory too small for the full-size workload")
    tree, tips, freq, _ = bench.make_inputs(ntaxa, nptn, model)
    try:
        g = LikelihoodEngine(model, ntaxa, nptn)
    except Exception as ex:   # another process holds device memory: nothing to learn from an allocation failure
        if "cudaMalloc" in str(ex) or "memory" in str(ex).lower():
            pytest.skip(f"device memory not available for the full-size workload: {ex}")
        raise
    g.set_tip_states(tips); g.set_pattern_freq(freq); g.set_pattern_invar(np.zeros(nptn))
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    ll = g.branch_logl(ops, a, b, L)
    plh = g.get_pattern_lh()
    assert np.isfinite(ll) and abs(float(plh.sum()) - ll) < 1e-9 * abs(ll)
    # the oracle on a slice of the same patterns
    o = OracleEngine(model, ntaxa, SLICE)
    o.set_tip_states(np.ascontiguousarray(tips[:, :SLICE])); o.set_pattern_freq(freq[:SLICE]); o.set_pattern_invar(np.zeros(SLICE))
    o.branch_logl(ops, a, b, L)
    np.testing.assert_allclose(plh[:SLICE], o.get_pattern_lh(), rtol=1e-11, atol=0)
    slot = a[1]
    sc_g, sc_o = g.get_scale_num(slot)[:SLICE], o.get_scale_num(slot)
    assert np.array_equal(sc_g, sc_o)
    assert sc_o.max() >= 1 or ntaxa < 500, "the full-size DNA case is expected to exercise 2^256 rescaling"
    # derivatives against central differences of logL along the evaluation branch
    df, ddf = g.branch_derv(ops[:0], a, b, L)
    h = 1e-4 * max(L, 1e-2)
    lp, lm = g.branch_logl(ops[:0], a, b, L + h), g.branch_logl(ops[:0], a, b, L - h)
    assert abs((lp - lm) / (2 * h) - df) < 1e-5 * max(1.0, abs(df)) + 1e-6 * abs(ll) * h
    assert abs((lp - 2 * ll + lm) / (h * h) - ddf) < 2e-3 * max(1.0, abs(ddf))
    # pulley principle: an internal branch far from the first one
    ops2, a2, b2, L2 = plan.plan_branch(tree, len(tree.lengths) // 2)
    ll2 = g.branch_logl(ops2, a2, b2, L2)
    assert abs(ll2 - ll) < 1e-9 * abs(ll), (ll, ll2)
    g.close()


def test_config4_c60_at_full_size_with_mem_save():
    """BASELINE config 4 (100 taxa x 50 000 patterns, LG+C60+G4: 240 categories, 1.92 GB per CLV) as bench.py runs it: the
    reference's -mem mode (LM_MEM_SAVE) planned by the host mirror, slots recycled inside one traversal."""
    import torch
    from iqtree2_b200.phylotree import PhyloTree
    ntaxa, nptn, mk, _ = bench.WORKLOADS["aa_100x50k_lg_c60_g4"]
    model = mk()
    B = model.nstates * model.ncat_mix
    slot_bytes = ((nptn + 127) // 128 * 128) * (8 * B + 2)
    dev_total = torch.cuda.get_device_properties(0).total_memory
    mem_frac = min(1.0, (0.70 * dev_total / slot_bytes - 4) / (ntaxa - 2))
    if mem_frac * (ntaxa - 2) < 12:
        pytest.skip("device memory too small for the full-size workload")
    tree, tips, freq, _ = bench.make_inputs(ntaxa, nptn, model)
    pt = PhyloTree()
    pt.set_plan_only(True)
    pt.set_alignment(tree.names, np.zeros((ntaxa, 1), dtype=np.uint8), np.ones(1), model)
    pt.set_model(model)
    pt.read_tree(tree.to_newick(fmt="%.17g"))   # the exact branch lengths
    pt.set_mem_save(True, mem_frac)
    nslots = pt.max_lh_slots + 2
    assert nslots - 2 < ntaxa - 2, "this test is about recycled slots"
    pa, pb, _ = pt.branches()
    pbr = [k for k in range(len(pa)) if min(pa[k], pb[k]) == 0][0]
    ops, a, b, L = pt.plan_branch(pbr)
    assert len(ops) == ntaxa - 2
    try:
        g = LikelihoodEngine(model, ntaxa, nptn, nslots=nslots)
    except Exception as ex:
        if "cudaMalloc" in str(ex) or "memory" in str(ex).lower():
            pytest.skip(f"device memory not available for the full-size workload: {ex}")
        raise
    g.set_tip_states(tips); g.set_pattern_freq(freq); g.set_pattern_invar(np.zeros(nptn))
    ll = g.branch_logl(ops, a, b, L)
    plh = g.get_pattern_lh()
    assert np.isfinite(ll) and abs(float(plh.sum()) - ll) < 1e-9 * abs(ll)
    sc_full = g.get_scale_num(a[1])
    df, ddf = g.branch_derv(ops[:0], a, b, L)
    h = 1e-4 * max(L, 1e-2)
    lp, lm = g.branch_logl(ops[:0], a, b, L + h), g.branch_logl(ops[:0], a, b, L - h)
    assert abs((lp - lm) / (2 * h) - df) < 1e-5 * max(1.0, abs(df)) + 1e-6 * abs(ll) * h
    # pulley principle under slot recycling: the planner keeps track of which CLVs survived the first traversal
    other = [k for k in range(len(pa)) if min(pa[k], pb[k]) >= ntaxa][len(pa) // 4]
    ops2, a2, b2, L2 = pt.plan_branch(other)
    ll2 = g.branch_logl(ops2, a2, b2, L2)
    assert abs(ll2 - ll) < 1e-9 * abs(ll), (ll, ll2)
    g.close()
    pt.close()
    # the same patterns with one slot per internal node (LM_PER_NODE) on a shard that fits: identical per-pattern results
    SH = 4096
    ops1, a1, b1, L1 = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    g1 = LikelihoodEngine(model, ntaxa, SH)
    g1.set_tip_states(np.ascontiguousarray(tips[:, :SH])); g1.set_pattern_freq(freq[:SH]); g1.set_pattern_invar(np.zeros(SH))
    g1.branch_logl(ops1, a1, b1, L1)
    np.testing.assert_allclose(plh[:SH], g1.get_pattern_lh(), rtol=1e-12, atol=0)
    g1.close()
    # the CPU oracle on a slice (240 categories x 400 flops per node: keep it small)
    SL = 128
    o = OracleEngine(model, ntaxa, SL)
    o.set_tip_states(np.ascontiguousarray(tips[:, :SL])); o.set_pattern_freq(freq[:SL]); o.set_pattern_invar(np.zeros(SL))
    o.branch_logl(ops1, a1, b1, L1)
    np.testing.assert_allclose(plh[:SL], o.get_pattern_lh(), rtol=1e-11, atol=0)
    assert np.array_equal(sc_full[:SL], o.get_scale_num(a1[1]))


def test_config5_500_partitions_against_the_reference_per_partition():
    """BASELINE config 5 as bench.py builds it (500 partitions x 100 taxa, 80 % GTR+G4 / 20 % LG+G4, 10 % of the taxa absent per
    partition): super-tree logL, partition scores and the linked (df, ddf) against the compiled reference kernel run per partition
    tree and combined as PhyloSuperTreePlen::computeFuncDerv does (tree/phylosupertreeplen.cpp:362-412: lengths x part_rate,
    df += rate*df_p, ddf += rate^2*ddf_p)."""
    from iqtree2_b200 import treesim
    from iqtree2_b200.supertree import PhyloSuperTree
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libiqref.so not built")
    ntaxa, nparts = 100, 500
    tree, parts = bench.make_super_inputs(ntaxa, nparts)
    st = PhyloSuperTree(tree.names)
    for p in parts:
        st.add_partition(p["pat"], p["freq"], p["model"], part_rate=p["rate"])
    st.read_tree(tree.to_newick(fmt="%.17g"))
    sa, sb, ln = st.branches()
    br = [k for k in range(len(sa)) if min(sa[k], sb[k]) == 0][0]   # the pendant branch of taxon 0
    ll = st.compute_likelihood_branch(br)
    scores = st.part_scores().copy()
    df, ddf = st.compute_likelihood_derv(br, float(ln[br]))
    st.close()
    ll_ref = df_ref = ddf_ref = 0.0
    mag = 0.0
    for k, p in enumerate(parts):
        scaled = treesim.Tree(tree.ntaxa, tree.edges, tree.lengths * p["rate"], tree.names)
        r = ref_engine.RefTree(scaled, p["model"], p["pat"], p["freq"], nthreads=4)
        t = r.time_eval(0, 1)
        r.close()
        assert abs(scores[k] - t["logl"]) < 1e-6 + 1e-12 * abs(t["logl"]), (k, scores[k], t["logl"])
        ll_ref += t["logl"]; df_ref += p["rate"] * t["df"]; ddf_ref += p["rate"] ** 2 * t["ddf"]
        mag += abs(p["rate"] * t["df"])
    assert abs(ll - ll_ref) < 1e-10 * abs(ll_ref), (ll, ll_ref)
    assert abs(df - df_ref) < 1e-9 * max(mag, 1.0), (df, df_ref)
    assert abs(ddf - ddf_ref) < 1e-8 * abs(ddf_ref), (ddf, ddf_ref)


def test_site_specific_workload_against_the_live_reference():
    """the bench workload aa_100x20k_pmsf_g4 (one frequency profile per pattern + G4) at its full size: k_site_walk20 against the
    reference's own SITE_MODEL kernel on every pattern — per-pattern log-likelihood 1e-11, logL / derivatives 1e-9, the pulley principle"""
    from oracle import ref_engine
    if not ref_engine.available():
        pytest.skip("oracle/_ref/libiqref.so not built")
    ntaxa, nptn, mk, _ = bench.WORKLOADS["aa_100x20k_pmsf_g4"]
    model = mk()
    tree, tips, freq, _ = bench.make_inputs(ntaxa, nptn, model)
    g = LikelihoodEngine(model, ntaxa, nptn)
    g.set_tip_states(tips); g.set_pattern_freq(freq); g.set_pattern_invar(np.zeros(nptn))
    r = ref_engine.RefTree(tree, model, tips, freq, nthreads=8)
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll_r, plh_r = r.branch_logl(0)
    ll_g = g.branch_logl(ops, a, b, L)
    assert abs(ll_g - ll_r) < 1e-9 * abs(ll_r), (ll_g, ll_r)
    np.testing.assert_allclose(g.get_pattern_lh(), plh_r, rtol=1e-11, atol=0)
    np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, L), r.branch_derv(0), rtol=1e-9)
    ops2, a2, b2, L2 = plan.plan_branch(tree, len(tree.lengths) // 2)
    g.theta_reset()
    assert abs(g.branch_logl(ops2, a2, b2, L2) - ll_r) < 1e-9 * abs(ll_r)
    g.close(); r.close()


def test_mixed_branch_lengths_with_equal_classes_equal_plain_lengths_at_config3_size():
    """config 3 (protein 200 x 100 000, LG+G4) evaluated with a mixlen table whose four classes all hold the branch's own length: the
    tuned 20-state walker reads its lengths through the table and must return the plain engine's numbers bit for bit; the gradient
    over the classes sums to the scalar derivative"""
    ntaxa, nptn, mk, _ = bench.WORKLOADS["aa_200x100k_lg_g4"]
    model = mk()
    tree, tips, freq, _ = bench.make_inputs(ntaxa, nptn, model)
    plain = LikelihoodEngine(model, ntaxa, nptn)
    mixl = LikelihoodEngine(model, ntaxa, nptn)
    for e in (plain, mixl):
        e.set_tip_states(tips); e.set_pattern_freq(freq); e.set_pattern_invar(np.zeros(nptn))
    lengths = np.asarray(tree.lengths, dtype=np.float64)
    mixl.set_mixlen(np.repeat(lengths[:, None], model.ncat, axis=1))
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    ops_r, a_r, b_r, row = plan.plan_branch(tree, br, lengths=np.arange(len(lengths), dtype=np.float64))
    l1, l2 = plain.branch_logl(ops, a, b, L), mixl.branch_logl(ops_r, a_r, b_r, row)
    assert l1 == l2
    np.testing.assert_array_equal(plain.get_pattern_lh(), mixl.get_pattern_lh())
    df, ddf = plain.branch_derv(ops[:0], a, b, L)
    grad, hess, ll = mixl.branch_derv_mixlen(ops_r[:0], a_r, b_r, row)
    assert abs(ll - l1) < 1e-9 * abs(l1)
    assert abs(grad.sum() - df) < 1e-9 * max(1.0, abs(df)) + 1e-12 * abs(l1)
    assert abs(hess.sum() - ddf) < 1e-8 * max(1.0, abs(ddf))
    plain.close(); mixl.close()
