"""Pattern-sharded likelihood across the GPUs of one box (one process per GPU, torch.distributed).

Patterns are independent end to end (every loop of the reference kernels is per pattern; the only cross-pattern
operations are the scalar sums tree_lh / df / ddf, tree/phylokernelnew.h:2375-2379, 2747-2748), so rank r owns the
contiguous pattern slice [lo_r, hi_r) of EVERY CLV slot, scale_num, theta, ptn_freq/invar and tip-state row.  No
CLV ever crosses NVLink; each evaluation exchanges one all-reduce(sum) of (logL) or (df, ddf) — SURVEY §8(e).

The local engine is injected: the product passes iqtree2_b200.engine.LikelihoodEngine (CUDA); the CPU-only gloo
tests pass the oracle, which is how the N>1 host logic is covered without a GPU.
"""
from __future__ import annotations

import numpy as np

TILE = 128  # slice boundaries are multiples of the kernels' pattern tile


def shard_bounds(nptn: int, world: int, rank: int, tile: int = TILE):
    """Contiguous slice of rank `rank`; every slice but the last is a multiple of `tile` patterns."""
    per = (nptn + world - 1) // world
    per = (per + tile - 1) // tile * tile
    lo = min(rank * per, nptn)
    hi = min((rank + 1) * per, nptn)
    return lo, hi


class ShardedLikelihood:
    """Same call surface as the engine (branch_logl / branch_derv / logl_from_buffer / newton_branch ...), results
    summed over ranks.  Ranks whose slice is empty (more ranks than tiles) take part in the collectives with zeros."""

    def __init__(self, engine_cls, model, ntaxa, tips, ptn_freq, ptn_invar=None, group=None, device=None, **engine_kw):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nptn_total = int(tips.shape[1])
        self.lo, self.hi = shard_bounds(self.nptn_total, self.world, self.rank)
        self.local_nptn = self.hi - self.lo
        self.device = device if device is not None else torch.device("cpu")
        self.engine = None
        if self.local_nptn > 0:
            self.engine = engine_cls(model, ntaxa, self.local_nptn, **engine_kw)
            self.engine.set_tip_states(np.ascontiguousarray(tips[:, self.lo:self.hi]))
            self.engine.set_pattern_freq(np.ascontiguousarray(ptn_freq[self.lo:self.hi]))
            inv = np.zeros(self.local_nptn) if ptn_invar is None else np.ascontiguousarray(ptn_invar[self.lo:self.hi])
            self.engine.set_pattern_invar(inv)
        self._buf = torch.zeros(3, dtype=torch.float64, device=self.device)

    def _allreduce(self, vals):
        if self.world == 1:
            return list(vals)
        t = self._buf
        t.zero_()
        t[:len(vals)] = self.torch.tensor(list(vals), dtype=self.torch.float64)
        self.dist.all_reduce(t, group=self.group)
        return t[:len(vals)].tolist()

    def set_model(self, model):
        if self.engine:
            self.engine.set_model(model)

    def theta_reset(self):
        if self.engine:
            self.engine.theta_reset()

    def update_partials(self, ops):
        if self.engine:
            self.engine.update_partials(ops)

    def branch_logl(self, ops, end_a, end_b, length, save_log_value=True):
        v = self.engine.branch_logl(ops, end_a, end_b, length, save_log_value) if self.engine else 0.0
        return self._allreduce([v])[0]

    def branch_derv(self, ops, end_a, end_b, length):
        v = self.engine.branch_derv(ops, end_a, end_b, length) if self.engine else (0.0, 0.0)
        return tuple(self._allreduce(v))

    def logl_from_buffer(self, length):
        v = self.engine.logl_from_buffer(length) if self.engine else 0.0
        return self._allreduce([v])[0]

    def em_posterior(self, wsl, add_invar):
        """E-step of the EM algorithms across ranks: every rank's category sums (engine.em_posterior on its pattern slice, the
        posteriors stay on that rank's device), one all-reduce of K doubles"""
        local = None
        if self.engine:
            r = self.engine.em_posterior(wsl, add_invar)
            local = np.asarray(r[0] if isinstance(r, tuple) else r, dtype=np.float64)   # (the oracle returns the matrix as well)
        if self.world == 1:
            return local
        klen = self.torch.tensor([0 if local is None else len(local)], dtype=self.torch.int64, device=self.device)
        self.dist.all_reduce(klen, op=self.dist.ReduceOp.MAX, group=self.group)
        t = self.torch.zeros(int(klen.item()), dtype=self.torch.float64, device=self.device)
        if local is not None:
            t += self.torch.from_numpy(local).to(self.device)
        self.dist.all_reduce(t, group=self.group)
        return t.cpu().numpy()

    def gather_pattern_lh(self):
        """computePatternLikelihood across ranks: all-gather of the per-rank _pattern_lh slices (8 B/pattern)."""
        local = self.engine.get_pattern_lh() if self.engine else np.zeros(0)
        if self.world == 1:
            return local
        per = shard_bounds(self.nptn_total, self.world, 0)[1]
        pad = self.torch.zeros(per, dtype=self.torch.float64, device=self.device)
        pad[:len(local)] = self.torch.from_numpy(local)
        out = [self.torch.zeros_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad, group=self.group)
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.nptn_total, self.world, r)
            parts.append(out[r][:hi - lo].cpu().numpy())
        return np.concatenate(parts)

    def close(self):
        if self.engine:
            self.engine.close()


def deal_partitions(costs, world: int, rank: int):
    """Partition model (PartitionModel / PhyloSuperTree): whole partitions go to ranks, most expensive first (the cost order of
    PhyloSuperTree::computePartitionOrder, tree/phylosupertree.cpp:653-695: taxa x patterns x states), round-robin, so that
    every rank gets the same mix.  Returns the partition indices of `rank`.  The only exchange of a partition-sharded
    evaluation is the sum of the ranks' (logL, df, ddf)."""
    import numpy as np
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    return [int(i) for k, i in enumerate(order) if k % world == rank]
