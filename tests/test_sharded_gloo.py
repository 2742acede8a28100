"""CPU, world_size 2, gloo: the pattern-sharding host logic (slice bounds, scalar all-reduce, pattern_lh gather).
The local engine here is the oracle (test infrastructure); on the GPU box bench.py runs the same class over the
CUDA engine with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import load_engine, make_case
from iqtree2_b200 import plan
from iqtree2_b200.sharded import ShardedLikelihood, shard_bounds


def test_shard_bounds_cover_everything():
    for nptn in (1, 127, 128, 129, 1000, 1_000_000):
        for world in (1, 2, 3, 8):
            covered = 0
            for r in range(world):
                lo, hi = shard_bounds(nptn, world, r)
                assert lo == covered or lo == hi == nptn
                assert r == world - 1 or (hi - lo) % 128 == 0 or hi == nptn
                covered = hi
            assert covered == nptn


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from oracle.oracle_engine import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, model, pat, freq, inv = make_case("dna_inv", ntaxa=11, nsites=900, seed=6)
    sh = ShardedLikelihood(OracleEngine, model, tree.ntaxa, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 3)
    ll = sh.branch_logl(ops, a, b, L)
    sh.theta_reset()
    df, ddf = sh.branch_derv(ops[:0], a, b, L)
    ll2 = sh.logl_from_buffer(L)
    plh = sh.gather_pattern_lh()
    prop = sh.em_posterior(2, True)          # E-step of RateFree::optimizeWithEM over both shards
    # host-driven Newton-Raphson with an all-reduce per step: both ranks must walk the same iterates
    sh.theta_reset()
    xopt, nsteps = sh.newton_branch(ops[:0], a, b, 1e-6, L, 10.0, 1e-6)
    # a shard whose derivative is not finite zeroes (df, ddf) on EVERY rank (tree/phylokernelnew.h:2629-2632)
    inv2 = inv.copy()
    inv2[5] = np.nan                          # a pattern of rank 0's slice
    sh2 = ShardedLikelihood(OracleEngine, model, tree.ntaxa, pat, freq, inv2)
    d_bad = sh2.branch_derv(ops, a, b, L)
    st_bad = sh2.last_status
    if rank == 0:
        np.save(out, np.concatenate([[ll, df, ddf, ll2], prop, [xopt, nsteps, d_bad[0], d_bad[1], st_bad], plh]))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_engine(tmp_path):
    from oracle.oracle_engine import OracleEngine
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    tree, model, pat, freq, inv = make_case("dna_inv", ntaxa=11, nsites=900, seed=6)
    assert pat.shape[1] > 128, "needs more than one tile to exercise two non-empty shards"
    o = load_engine(OracleEngine, tree, model, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 3)
    ll = o.branch_logl(ops, a, b, L)
    o.theta_reset()
    df, ddf = o.branch_derv(ops[:0], a, b, L)
    assert abs(res[0] - ll) < 1e-8 and abs(res[3] - ll) < 1e-8
    np.testing.assert_allclose(res[1:3], [df, ddf], rtol=1e-10)
    np.testing.assert_allclose(res[4:8], o.em_posterior(2, True)[0], rtol=1e-12)
    np.testing.assert_allclose(res[13:], o.get_pattern_lh(), rtol=1e-13)
    # the optimum of the sharded Newton iteration is a root of the single engine's derivative
    xopt, nsteps = res[8], res[9]
    assert nsteps >= 2
    d_at = o.branch_derv(ops[:0], a, b, float(xopt))
    d_l = o.branch_derv(ops[:0], a, b, L)
    assert abs(d_at[0]) < 1e-3 * max(1.0, abs(d_l[0])) or xopt in (1e-6, 10.0)
    assert res[10] == 0.0 and res[11] == 0.0 and res[12] == 2


def _worker_site_mixlen(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from oracle.oracle_engine import OracleEngine
    from test_mixlen_cpu import make_mixlen_case, mixlen_table, plan_rows
    from test_site_model_cpu import make_site_case
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # site-specific model: every rank takes its slice of the per-pattern eigensystems
    tree, model, pat, freq, inv = make_site_case(4, 4, ntaxa=9, nptn=300, seed=19)
    sh = ShardedLikelihood(OracleEngine, model, tree.ntaxa, pat, freq, inv)
    ops, a, b, L = plan.plan_branch(tree, 2)
    ll = sh.branch_logl(ops, a, b, L)
    df, ddf = sh.branch_derv(ops[:0], a, b, L)
    plh = sh.gather_pattern_lh()
    # mixed branch lengths: replicated table, gradient / Hessian summed over the ranks
    tree2, model2, pat2, freq2, inv2 = make_mixlen_case("dna", 4, ntaxa=9, nsites=700, seed=23)
    sm = ShardedLikelihood(OracleEngine, model2, tree2.ntaxa, pat2, freq2, inv2)
    sm.set_mixlen(mixlen_table(tree2, model2))
    ops2, a2, b2, row = plan_rows(tree2, 3)
    llm = sm.branch_logl(ops2, a2, b2, row)
    g, H, ll3 = sm.branch_derv_mixlen(ops2[:0], a2, b2, row)
    dc = sm.branch_derv_mixlen_cur(ops2[:0], a2, b2, row, 2)
    if rank == 0:
        np.save(out, np.concatenate([[ll, df, ddf, llm, ll3, dc[0], dc[1]], g, H.ravel(), plh]))
    dist.destroy_process_group()


def test_two_rank_gloo_site_specific_model_and_mixed_branch_lengths(tmp_path):
    from oracle.oracle_engine import OracleEngine
    from test_mixlen_cpu import make_mixlen_case, mixlen_engine, plan_rows
    from test_site_model_cpu import engine, make_site_case
    out = str(tmp_path / "res2.npy")
    mp.spawn(_worker_site_mixlen, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    tree, model, pat, freq, inv = make_site_case(4, 4, ntaxa=9, nptn=300, seed=19)
    o = engine(OracleEngine, tree, model, pat, freq, inv, False)
    ops, a, b, L = plan.plan_branch(tree, 2)
    ll = o.branch_logl(ops, a, b, L)
    assert abs(res[0] - ll) < 1e-9 * abs(ll)
    np.testing.assert_allclose(res[1:3], o.branch_derv(ops[:0], a, b, L), rtol=1e-10)
    np.testing.assert_allclose(res[7 + 4 + 16:], o.get_pattern_lh(), rtol=1e-13)
    tree2, model2, pat2, freq2, inv2 = make_mixlen_case("dna", 4, ntaxa=9, nsites=700, seed=23)
    assert pat2.shape[1] > 128
    m = mixlen_engine(OracleEngine, tree2, model2, pat2, freq2, inv2)
    ops2, a2, b2, row = plan_rows(tree2, 3)
    llm = m.branch_logl(ops2, a2, b2, row)
    g, H, ll3 = m.branch_derv_mixlen(ops2[:0], a2, b2, row)
    assert abs(res[3] - llm) < 1e-9 * abs(llm) and abs(res[4] - ll3) < 1e-9 * abs(ll3)
    np.testing.assert_allclose(res[5:7], m.branch_derv_mixlen_cur(ops2[:0], a2, b2, row, 2), rtol=1e-10)
    np.testing.assert_allclose(res[7:11], g, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(res[11:27].reshape(4, 4), H, rtol=1e-10, atol=1e-8)
