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
    """branch_logl / branch_derv / logl_from_buffer / newton_branch / em_posterior / gather_pattern_lh of the engine with the
    results combined over ranks.  Ranks whose slice is empty (more ranks than tiles) take part in the collectives with zeros.

    * a numerical status of ANY rank is everybody's: the engine's warning code is all-reduced with MAX (`last_status`), and when
      a shard's derivative was not finite the reference's df = ddf = 0 (tree/phylokernelnew.h:2629-2632) is returned on every
      rank instead of a sum over the healthy shards;
    * newton_branch is the reference's minimizeNewton (utils/optimization.cpp:422-499) driven from the host with one
      all-reduce of (df, ddf) per step: every rank walks the same iterates (the device-resident iteration of a single engine
      would diverge between ranks);
    * Lewis' +ASC (needs the global prob_const) and robust phylo (a global sort of the site log-likelihoods) are not separable
      per rank: the constructor refuses them when world > 1."""

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
            from . import models as _models
            # a site-specific model (ModelSet / PMSF) carries one eigensystem per pattern: the rank takes its slice of them
            self.engine = engine_cls(_models.site_slice(model, self.lo, self.hi), ntaxa, self.local_nptn, **engine_kw)
            self.engine.set_tip_states(np.ascontiguousarray(tips[:, self.lo:self.hi]))
            self.engine.set_pattern_freq(np.ascontiguousarray(ptn_freq[self.lo:self.hi]))
            inv = np.zeros(self.local_nptn) if ptn_invar is None else np.ascontiguousarray(ptn_invar[self.lo:self.hi])
            self.engine.set_pattern_invar(inv)
        self._buf = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.last_status = 0
        if self.world > 1 and (engine_kw.get("asc_type", 0) == 1 or engine_kw.get("robust", False)):
            raise ValueError("Lewis' +ASC and robust phylo are not separable per pattern shard: use one engine")

    def set_ascertainment(self, asc_type, num_unobserved, nsite):
        if self.world > 1 and asc_type == 1:
            raise ValueError("Lewis' +ASC needs the global prob_const: not available with pattern sharding")
        if self.engine:
            self.engine.set_ascertainment(asc_type, num_unobserved, nsite)

    def set_robust_phylo(self, keep, median):
        if self.world > 1 and (keep < 1.0 or median):
            raise ValueError("robust phylo sorts all site log-likelihoods: not available with pattern sharding")
        if self.engine:
            self.engine.set_robust_phylo(keep, median)

    def _allreduce(self, vals):
        """sum of vals over ranks; the engines' warning status travels in the same message (MAX via a second reduce only when
        some rank reported one is not possible without knowing it: it is summed as a count instead)"""
        st = float(getattr(self.engine, "last_status", 0) or 0) if self.engine else 0.0
        if self.world == 1:
            self.last_status = int(st)
            return list(vals)
        t = self._buf
        t.zero_()
        t[:len(vals)] = self.torch.tensor(list(vals), dtype=self.torch.float64)
        t[3] = 1.0 if st > 0 else 0.0
        self.dist.all_reduce(t, group=self.group)
        self.last_status = int(st) if t[3].item() == 0 else max(int(st), 1)
        self._any_warn = t[3].item() > 0
        return t[:len(vals)].tolist()

    def set_model(self, model):
        if self.engine:
            from . import models as _models
            self.engine.set_model(_models.site_slice(model, self.lo, self.hi))

    # ---- mixed branch lengths (heterotachy): the table is replicated, the gradient / Hessian are sums over the ranks ----------
    def set_mixlen(self, table):
        self._nmixlen = 0 if table is None else int(np.asarray(table).shape[1])
        if self.engine:
            self.engine.set_mixlen(table)

    def branch_derv_mixlen(self, ops, end_a, end_b, row):
        """computeLikelihoodDerv of a mixlen tree over all shards -> (gradient [K], hessian [K][K], logL)"""
        K = self._nmixlen
        if self.engine:
            g, H, ll = self.engine.branch_derv_mixlen(ops, end_a, end_b, row)
            v = np.concatenate([g, H.ravel(), [ll]])
        else:
            v = np.zeros(K + K * K + 1)
        if self.world > 1:
            t = self.torch.from_numpy(v).to(self.device)
            self.dist.all_reduce(t, group=self.group)
            v = t.cpu().numpy()
        return v[:K].copy(), v[K:K + K * K].reshape(K, K).copy(), float(v[-1])

    def branch_derv_mixlen_cur(self, ops, end_a, end_b, row, cur):
        v = self.engine.branch_derv_mixlen_cur(ops, end_a, end_b, row, cur) if self.engine else (0.0, 0.0)
        return tuple(self._allreduce(v))

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
        out = tuple(self._allreduce(v))
        if self.world > 1 and self._any_warn:   # some shard's derivative was not finite: df = ddf = 0 everywhere (:2629-2632)
            self.last_status = 2
            return (0.0, 0.0)
        return out

    def newton_branch(self, ops, end_a, end_b, min_len, guess, max_len, tolerance, max_steps=100):
        """Optimization::minimizeNewton (utils/optimization.cpp:422-499) on f = -df, f' = -ddf summed over the ranks."""
        steps = 0

        def fd(x, first):
            nonlocal steps
            steps += 1
            df, ddf = self.branch_derv(ops if first else ops[:0], end_a, end_b, x)
            return -df, -ddf
        rts = min(max(guess, min_len), max_len)
        f, df = fd(rts, True)
        if not (np.isfinite(f) and np.isfinite(df)) or self.last_status == 2:
            return rts, steps
        if df >= 0.0 and abs(f) < tolerance:
            return rts, steps
        xl, xh = (rts, max_len) if f < 0.0 else (min_len, rts)
        for j in range(1, max_steps + 1):
            rts_old = rts
            if df <= 0.0 or ((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0:
                dx = 0.5 * (xh - xl)
                rts = xl + dx
                if xl == rts:
                    return rts, steps
            else:
                dx = f / df
                temp = rts
                rts -= dx
                if temp == rts:
                    return rts, steps
            if abs(dx) < tolerance or j == max_steps:
                return rts_old, steps
            f, df = fd(rts, False)
            if not (np.isfinite(f) and np.isfinite(df)) or self.last_status == 2:
                return rts, steps
            if df > 0.0 and abs(f) < tolerance:
                return rts, steps
            if f < 0.0:
                xl = rts
            elif f > 0.0:
                xh = rts
        return rts, steps

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
