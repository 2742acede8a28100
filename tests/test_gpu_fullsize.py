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
