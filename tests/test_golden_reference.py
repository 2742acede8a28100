"""Golden vectors minted from the REAL reference kernel (tests/golden/make_golden.py, oracle/_ref/libiqref.so) checked
against the CPU oracle here and against the CUDA engine in test_gpu_reference.py.  These files travel with the repo, so
the parity chain  reference -> golden -> {oracle, CUDA}  holds on boxes where /root/reference does not exist."""
import os
import sys

import numpy as np
import pytest

from cases import load_engine, make_case
from iqtree2_b200 import plan
from oracle.oracle_engine import OracleEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
from make_golden import GOLDEN_CASES, input_digest  # noqa: E402


def load_case(name):
    kind, safe, kw = GOLDEN_CASES[name]
    tree, model, pat, freq, inv = make_case(kind, **kw)
    g = np.load(os.path.join(GOLD, f"ref_{name}.npz"))
    assert str(g["digest"]) == input_digest(tree, model, pat, freq), "seeded inputs drifted from the golden file's"
    br = plan.pendant_branch_of(tree, 0)
    ops, a, b, L = plan.plan_branch(tree, br)
    return tree, model, pat, freq, inv, safe, g, ops, a, b, L


def check_engine_against_golden(e, g, ops, a, b, L, inv):
    """shared by the oracle (CPU) and the CUDA engine (GPU): the north-star bars against the reference's outputs"""
    np.testing.assert_array_equal(inv, g["ptn_invar"])
    e.set_reference_quirks(1)
    ll = e.branch_logl(ops, a, b, L)
    assert abs(ll - float(g["logl"])) < 1e-6                                             # tree logL: 1e-6 log-units
    np.testing.assert_allclose(e.get_pattern_lh(), g["pattern_lh"], rtol=1e-11)          # per-pattern logL: 1e-11 relative
    cat_r = g["pattern_lh_cat"]
    cmax = np.abs(cat_r).max(axis=1, keepdims=True)
    np.testing.assert_allclose(e.get_pattern_lh_cat().reshape(cat_r.shape) / cmax, cat_r / cmax, rtol=1e-10, atol=1e-12)
    sc = g["scale_num"]
    for k, op in enumerate(ops):                                                          # scale counts: bit-exact
        np.testing.assert_array_equal(e.get_scale_num(int(op["dst_slot"])).reshape(sc[k].shape), sc[k])
    for key, op in (("partial_first", ops[0]), ("partial_last", ops[-1])):
        ref = g[key]
        scale = np.abs(ref).max(axis=1, keepdims=True)
        np.testing.assert_allclose(e.get_partial(int(op["dst_slot"])).reshape(ref.shape) / scale, ref / scale, rtol=0, atol=1e-11)
    df, ddf = e.branch_derv(ops[:0], a, b, L)
    assert abs(df - float(g["df"])) <= 1e-8 * max(1.0, abs(float(g["df"])))
    assert abs(ddf - float(g["ddf"])) <= 1e-8 * max(1.0, abs(float(g["ddf"])))
    np.testing.assert_array_equal(e.get_buffer_scale_all(), g["buffer_scale_all"])
    assert abs(e.logl_from_buffer(L) - float(g["logl_from_buffer"])) < 1e-6


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_against_reference_golden(name):
    tree, model, pat, freq, inv, safe, g, ops, a, b, L = load_case(name)
    o = load_engine(OracleEngine, tree, model, pat, freq, inv, safe=safe)
    check_engine_against_golden(o, g, ops, a, b, L, inv)


def test_config1_golden_tree_logic_over_oracle():
    from test_tree_logic_cpu import example_case, make_pytree
    g = np.load(os.path.join(GOLD, "ref_config1_example.npz"))
    names, pat, freq, model, newick = example_case()
    t = make_pytree(names, pat, freq, model, newick)
    assert abs(t.compute_likelihood() - float(g["logl"])) < 1e-6
    np.testing.assert_allclose(t.engine.get_pattern_lh(), g["pattern_lh"], rtol=1e-11)
    assert abs(t.optimize_all_branches(100, 0.001, 100) - float(g["opt_logl"])) < 1e-6
    np.testing.assert_allclose(t.lengths(), g["opt_lengths"], rtol=0, atol=1e-6)           # optimised branch lengths: 1e-6
