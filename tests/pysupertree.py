"""Test infrastructure: Python restatement of PhyloSuperTreePlen (tree/phylosupertreeplen.cpp:243-412,
tree/phylosupertree.cpp:697-726) over per-partition PyTree objects (tests/pytree.py) driving the CPU oracle.
Every partition tree has the super tree's topology; its branch lengths are the super tree's times part_rate; a taxon
without data is a row of STATE_UNKNOWN.  logL = sum logL_p, df = sum rate_p df_p, ddf = sum rate_p^2 ddf_p."""
from __future__ import annotations

import numpy as np

from pytree import PyTree


class _NoEngine:
    """the super tree owns no likelihood buffers"""


class PySuperTree(PyTree):
    def __init__(self, names, newick, part_factories, part_rates, min_bl=1e-6, max_bl=10.0):
        super().__init__(lambda ntaxa, nslots: _NoEngine(), names, newick, min_bl, max_bl)
        self.parts = [PyTree(f, names, newick, min_bl, max_bl) for f in part_factories]
        self.part_rate = list(part_rates)
        self.cur_score = [0.0] * len(self.parts)
        self.map_trees()

    def map_trees(self):
        for p, r in zip(self.parts, self.part_rate):
            for (ab, ba), (sab, _) in zip(p.branches, self.branches):
                ab.length = ba.length = sab.length * r

    def _link(self, p):
        """the partition's neighbours of the super tree's current branch (same branch id, same direction)"""
        bid = self.current_it.id
        ab, ba = p.branches[bid]
        n1 = ab if ab.node.id == self.current_it.node.id else ba
        n2 = ba if n1 is ab else ab
        assert n1.node.id == self.current_it.node.id and n2.node.id == self.current_it_back.node.id
        return n1, n2

    def _set_lengths(self, value):
        self.current_it.length = self.current_it_back.length = value
        for p, r in zip(self.parts, self.part_rate):
            n1, n2 = self._link(p)
            n1.length = n2.length = value * r
            p.current_it, p.current_it_back = n1, n2

    def clear_all(self):
        for p in self.parts:
            p.clear_all()
        self.cur_score = [0.0] * len(self.parts)

    def branch_logl(self, dad_branch, dad):
        self.current_it = dad_branch
        self.current_it_back = dad_branch.node.find(dad)
        self._set_lengths(dad_branch.length)
        for k, p in enumerate(self.parts):
            self.cur_score[k] = p.branch_logl(p.current_it_back, p.current_it.node)
        return float(sum(self.cur_score))

    def from_buffer(self):
        return float(sum(self.cur_score))

    def _func_derv(self, value):
        self._set_lengths(value)
        df = ddf = 0.0
        for p, r in zip(self.parts, self.part_rate):
            d1, d2 = p.branch_derv(p.current_it_back, p.current_it.node)
            df += r * d1
            ddf += r * r * d2
        return -df, -ddf

    def func_derv_at(self, bid, value):
        ab, ba = self.branches[bid]
        if self.current_it is not ab:
            for p in self.parts:
                p.theta_computed = False
        self.current_it, self.current_it_back = ab, ba
        f, g = self._func_derv(value)
        return -f, -g

    def optimize_one_branch(self, node1, node2, clear_lh=True, max_steps=100):
        self.current_it = node1.find(node2)
        self.current_it_back = node2.find(node1)
        cur = self.current_it.length
        for p in self.parts:
            p.theta_computed = False
        super().optimize_one_branch(node1, node2, False, max_steps)
        self._set_lengths(self.current_it.length)
        for k, p in enumerate(self.parts):
            self.cur_score[k] = p.from_buffer()
        if clear_lh and cur != self.current_it.length:
            for p in self.parts:
                n1, n2 = self._link(p)
                p.clear_reverse(n1.node, n2.node)
                p.clear_reverse(n2.node, n1.node)

    def optimize_all_branches(self, iterations=100, tolerance=0.001, max_steps=100):
        self.cur_score = [0.0] * len(self.parts)
        return super().optimize_all_branches(iterations, tolerance, max_steps)
