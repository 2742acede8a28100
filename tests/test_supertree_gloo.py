"""CPU, world_size 2, gloo: partition sharding of the edge-linked partition model (bench.py --workload super_... --gpus N):
partitions are dealt to the ranks in cost order (iqtree2_b200.sharded.deal_partitions), every rank evaluates its own
PhyloSuperTreePlen and the (logL, df, ddf) scalars are all-reduced.  The local engines here are the oracle's (test
infrastructure); the sum over ranks must be the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iqtree2_b200.sharded import deal_partitions


def test_deal_partitions_is_a_partition_of_the_partitions():
    costs = [5, 50, 7, 7, 1, 90, 33, 2, 8]
    for world in (1, 2, 3, 4):
        got = sorted(sum((deal_partitions(costs, world, r) for r in range(world)), []))
        assert got == list(range(len(costs)))
        loads = [sum(costs[i] for i in deal_partitions(costs, world, r)) for r in range(world)]
        assert max(loads) - min(loads) <= max(costs)
    assert deal_partitions(costs, 2, 0)[0] == 5 and deal_partitions(costs, 2, 1)[0] == 1   # most expensive first, alternating


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    from supercases import make_supercase
    return make_supercase(ntaxa=9, nparts=5, seed=12, sites=(60, 140))


def _worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    sys.path.insert(0, os.path.dirname(here))
    from test_supertree_cpu import make_pysuper
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, parts = _case()
    mine = deal_partitions([p["pat"].shape[1] * p["model"].nstates for p in parts], world, rank)
    st = make_pysuper(tree, [parts[i] for i in mine])
    ll = st.compute_likelihood()
    df, ddf = st.func_derv_at(2, 0.05)
    t = torch.tensor([ll, df, ddf], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_two_rank_partition_sharding_matches_single_process(tmp_path):
    from test_supertree_cpu import make_pysuper
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    tree, parts = _case()
    st = make_pysuper(tree, parts)
    ll = st.compute_likelihood()
    df, ddf = st.func_derv_at(2, 0.05)
    assert abs(res[0] - ll) < 1e-8
    np.testing.assert_allclose(res[1:], [df, ddf], rtol=1e-10)


def _worker_cpp(rank, world, port, out):
    """the same sharding with the product's C++ host mirror (b200::PhyloSuperTreePlen) in every rank, its C ABI served by the oracle
    (oracle/_build/libb200pt_over_oracle.so, test infrastructure)"""
    import ctypes
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    sys.path.insert(0, here)
    sys.path.insert(0, root)
    subprocess.check_call(["make", "-C", os.path.join(root, "oracle"), "-s"])
    lib = ctypes.CDLL(os.path.join(root, "oracle", "_build", "libb200pt_over_oracle.so"))
    import iqtree2_b200.supertree as st
    st.load_library = lambda: lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tree, parts = _case()
    mine = deal_partitions([p["pat"].shape[1] * p["model"].nstates for p in parts], world, rank)
    g = st.PhyloSuperTree(tree.names, num_threads=1)
    for i in mine:
        p = parts[i]
        g.add_partition(p["pat"], p["freq"], p["model"], part_rate=p["rate"], ptn_invar=p["inv"])
    g.read_tree(tree.to_newick(fmt="%.17g"))
    ll = g.compute_likelihood()
    df, ddf = g.compute_likelihood_derv(2, 0.05)
    t = torch.tensor([ll, df, ddf], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        np.save(out, t.numpy())
    g.close()
    dist.destroy_process_group()


def test_two_rank_partition_sharding_with_the_cpp_host_mirror(tmp_path):
    from test_supertree_cpu import make_pysuper
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker_cpp, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    tree, parts = _case()
    st = make_pysuper(tree, parts)
    ll = st.compute_likelihood()
    df, ddf = st.func_derv_at(2, 0.05)
    assert abs(res[0] - ll) < 1e-8
    np.testing.assert_allclose(res[1:], [df, ddf], rtol=1e-9)
