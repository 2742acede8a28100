"""CPU: the robust-phylo variants of computeLikelihoodBranch (Params::robust_phy_keep < 1: log-likelihood of the best sites;
Params::robust_median: the median site log-likelihood; tree/phylokernelnew.h:3134-3173) in the oracle against the compiled
reference.  The harness alignment registers one site per pattern, so the cases expand every pattern to its sites."""
import numpy as np
import pytest

from cases import make_case
from iqtree2_b200 import plan
from oracle import ref_engine
from oracle.oracle_engine import OracleEngine


def site_case(kind="dna", ntaxa=12, nsites=401, seed=5):
    tree, model, pat, freq, inv = make_case(kind, ntaxa=ntaxa, nsites=nsites, seed=seed, gaps=0.03, ambig=0.0)
    pat = np.ascontiguousarray(np.repeat(pat, freq.astype(int), axis=1))   # one pattern per site
    return tree, model, pat, np.ones(pat.shape[1])


@pytest.mark.skipif(not ref_engine.available(), reason="oracle/_ref/libiqref.so not built")
@pytest.mark.parametrize("kind,nsites,keep,median", [("dna", 401, 0.9, False), ("dna", 400, 0.5, False), ("aa", 150, 0.73, False),
                                                     ("dna", 401, 1.0, True), ("dna", 400, 1.0, True)])
def test_oracle_robust_phylo_matches_reference(kind, nsites, keep, median):
    tree, model, pat, freq = site_case(kind, nsites=nsites)
    n = pat.shape[1]
    assert n == nsites
    o = OracleEngine(model, tree.ntaxa, n)
    o.set_tip_states(pat); o.set_pattern_freq(freq); o.set_pattern_invar(np.zeros(n))
    ops, a, b, L = plan.plan_branch(tree, plan.pendant_branch_of(tree, 0))
    ll_plain = o.branch_logl(ops, a, b, L)
    o.set_robust_phylo(keep, median)
    ll = o.branch_logl(ops[:0], a, b, L)
    try:
        ref_engine.set_robust(keep, median)
        r = ref_engine.RefTree(tree, model, pat, freq)
        ll_r, plh_r = r.branch_logl(0)
    finally:
        ref_engine.set_robust(1.0, False)
    assert abs(ll - ll_r) < 1e-9 * max(1.0, abs(ll_r)), (ll, ll_r)
    np.testing.assert_allclose(o.get_pattern_lh(), plh_r, rtol=1e-12)
    # what it means: the sorted site log-likelihoods
    s = np.sort(plh_r)
    expect = (s[n // 2] if n & 1 else 0.5 * (s[n // 2] + s[n // 2 - 1])) if median else s[int((1.0 - keep) * n):].sum()
    assert abs(ll - expect) < 1e-9 * max(1.0, abs(expect))
    assert ll > ll_plain
    r.close()
