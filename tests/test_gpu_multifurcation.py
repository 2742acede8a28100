"""GPU: multifurcating nodes (node->degree() > 3, tree/phylokernelnew.h:1407-1631) through the C ABI — ops with
`reserved` extra children + b200lk_set_extra_children — against the CPU oracle (pinned against the compiled reference on the
same polytomous trees in test_multifurcation_cpu.py): logL, per-pattern logL, every CLV and scale count, derivatives; NORM and
SAFE, 4-state, 20-state (fragment-major CLV layout) and a generic state count; walker and levelled schedules."""
import numpy as np
import pytest

from iqtree2_b200.engine import SCHEDULE_LEVELLED, SCHEDULE_WALKER, LikelihoodEngine
from multifurc import plan_pendant
from oracle.oracle_engine import OracleEngine
from test_gpu_parity import assert_partials_close
from test_multifurcation_cpu import polytomy_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["dna", "aa", "n5mix"])
@pytest.mark.parametrize("safe", [False, True])
@pytest.mark.parametrize("schedule", [SCHEDULE_WALKER, SCHEDULE_LEVELLED])
def test_multifurcating_nodes(kind, safe, schedule):
    tree, model, pat, freq, inv, adj = polytomy_case(kind, safe, ntaxa=40, nsites=500, ncollapse=5, seed=9)
    ops, extra, a, b, L = plan_pendant(adj, tree.ntaxa, leaf=0)
    assert len(extra) >= 3
    g = LikelihoodEngine(model, tree.ntaxa, pat.shape[1], safe_numeric=safe, schedule=schedule)
    o = OracleEngine(model, tree.ntaxa, pat.shape[1], safe_numeric=safe)
    for e in (g, o):
        e.set_tip_states(pat); e.set_pattern_freq(freq); e.set_pattern_invar(inv)
        e.set_extra_children(extra)
    ll_g, ll_o = g.branch_logl(ops, a, b, L), o.branch_logl(ops, a, b, L)
    assert abs(ll_g - ll_o) < 1e-6, (ll_g, ll_o)
    np.testing.assert_allclose(g.get_pattern_lh(), o.get_pattern_lh(), rtol=1e-11, atol=0)
    for op in ops:
        s = int(op["dst_slot"])
        assert np.array_equal(g.get_scale_num(s), o.get_scale_num(s)), f"scale_num differs in slot {s}"
        assert_partials_close(g.get_partial(s), o.get_partial(s), 1e-10)
    np.testing.assert_allclose(g.branch_derv(ops[:0], a, b, L), o.branch_derv(ops[:0], a, b, L), rtol=1e-9, atol=1e-7)
    # a second call without extra children must fail loudly, not read stale ones
    with pytest.raises(Exception, match="extra children"):
        g.update_partials(ops)
    g.close()
