"""CPU: multifurcating nodes (node->degree() > 3, tree/phylokernelnew.h:1407-1631) in the oracle against the compiled
reference reading the same polytomous Newick: tree logL, per-pattern logL, derivatives; NORM and SAFE, DNA and protein."""
import numpy as np
import pytest

from cases import make_case
from multifurc import collapse, newick, plan_pendant
from oracle import ref_engine
from oracle.oracle_engine import OracleEngine


def internal_branches(tree):
    return [b for b, (u, v) in enumerate(tree.edges) if u >= tree.ntaxa and v >= tree.ntaxa]


def polytomy_case(kind, safe, ntaxa=14, nsites=300, ncollapse=3, seed=6):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=nsites, seed=seed, mean_len=0.3)
    ib = internal_branches(tree)
    # pick pairwise non-adjacent internal branches so that contractions do not interact, plus one adjacent pair -> degree 5
    chosen, used = [], set()
    for b in ib:
        u, v = map(int, tree.edges[b])
        if u in used or v in used:
            continue
        chosen.append(b); used.update((u, v))
        if len(chosen) == ncollapse:
            break
    adj = collapse(tree, chosen)
    return tree, model, pat, freq, inv, adj


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind", ["dna", "aa"])
@pytest.mark.parametrize("safe", [False, True])
def test_oracle_multifurcation_matches_reference(kind, safe):
    tree, model, pat, freq, inv, adj = polytomy_case(kind, safe)
    assert max(len(a) for a in adj.values()) >= 4
    ops, extra, a, b, L = plan_pendant(adj, tree.ntaxa, leaf=0)
    assert ops["pad"].sum() == len(extra) > 0
    o = OracleEngine(model, tree.ntaxa, pat.shape[1], safe_numeric=safe)
    o.set_tip_states(pat); o.set_pattern_freq(freq); o.set_pattern_invar(inv)
    o.set_extra_children(extra)
    ll = o.branch_logl(ops, a, b, L)
    r = ref_engine.RefTree(None, model, pat, freq, safe_numeric=safe, newick=newick(adj, tree.names), names=tree.names)
    ll_r, plh_r = r.branch_logl(0)
    assert abs(ll - ll_r) < 1e-8 * max(1.0, abs(ll_r) * 1e-3), (ll, ll_r)
    np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12, atol=0)
    df_r = r.branch_derv(0)
    np.testing.assert_allclose(o.branch_derv(ops[:0], a, b, L), df_r, rtol=1e-9, atol=1e-9)
    r.close()
