"""Test infrastructure: a pure-Python restatement of the reference's tree-level likelihood logic (PhyloNeighbor flags,
slot re-orientation, traversal order, Newton-Raphson branch optimisation) over ANY engine with the engine interface —
in the tests, over the CPU oracle.  The C++ host mirror over the CUDA engine is checked against it.

Follows tree/phylotree.cpp:1233-1274 (computeLikelihood), :2594-2774 (optimizeOneBranch / optimizeAllBranches),
:5898-6028 (reorientPartialLh / computeTraversalInfo), tree/mtree.cpp:3040-3082, utils/optimization.cpp:422-499.
"""
from __future__ import annotations

import numpy as np

from iqtree2_b200.plan import OP_DTYPE


class Nei:
    __slots__ = ("node", "length", "id", "slot", "computed", "size")

    def __init__(self, node, length, bid):
        self.node, self.length, self.id = node, length, bid
        self.slot, self.computed, self.size = -1, False, 0


class Node:
    def __init__(self):
        self.id, self.name, self.neighbors, self.height = -1, "", [], 0.0

    def is_leaf(self):
        return len(self.neighbors) <= 1

    def find(self, other):
        for n in self.neighbors:
            if n.node is other:
                return n
        return None


class PyTree:
    def __init__(self, engine_factory, names, newick, min_bl=1e-6, max_bl=10.0):
        self.names = list(names)
        self.min_bl, self.max_bl = min_bl, max_bl
        self.nodes, self.branches = [], []
        self._parse(newick)
        self.engine = engine_factory(len(self.names), (len(self.nodes) - len(self.names)) + 2)
        self.current_it = self.current_it_back = None
        self.theta_computed = False
        self.num_partial = 0
        self._assign_slots()

    # ---- Newick -----------------------------------------------------------------------------------------------
    def _parse(self, s):
        row = {n: i for i, n in enumerate(self.names)}
        pos = [0]
        nxt = [len(self.names)]
        first_leaf = [None]

        def connect(a, b, ln):
            bid = len(self.branches)
            ab, ba = Nei(b, ln, bid), Nei(a, ln, bid)
            a.neighbors.append(ab); b.neighbors.append(ba)
            self.branches.append((ab, ba))

        def skip():
            while pos[0] < len(s) and s[pos[0]].isspace():
                pos[0] += 1

        def parse():
            skip()
            node = Node(); self.nodes.append(node)
            if s[pos[0]] == "(":
                pos[0] += 1
                kids = []
                while True:
                    kids.append(parse()); skip()
                    if s[pos[0]] == ",":
                        pos[0] += 1; continue
                    if s[pos[0]] == ")":
                        pos[0] += 1; break
                    raise ValueError("newick")
                node.id = nxt[0]; nxt[0] += 1
                for k, ln in kids:
                    connect(node, k, ln)
                skip()
                b = pos[0]
                while pos[0] < len(s) and s[pos[0]] not in ":,);":
                    pos[0] += 1
                node.name = s[b:pos[0]]
            else:
                b = pos[0]
                while pos[0] < len(s) and s[pos[0]] not in ":,);" and not s[pos[0]].isspace():
                    pos[0] += 1
                node.name = s[b:pos[0]]
                node.id = row[node.name]
                if first_leaf[0] is None:
                    first_leaf[0] = node
            skip()
            ln = -1.0
            if pos[0] < len(s) and s[pos[0]] == ":":
                pos[0] += 1
                b = pos[0]
                while pos[0] < len(s) and s[pos[0]] not in ",);":
                    pos[0] += 1
                ln = float(s[b:pos[0]])
            return node, ln

        top, _ = parse()
        assert len(top.neighbors) == 3, "test trees are written unrooted (trifurcating top node)"
        for ab, ba in self.branches:
            if ab.length < 0:
                ab.length = ba.length = self.min_bl
        # PhyloTree::readTreeString -> setRootNode(params.root == NULL): root = the leaf of alignment row 0
        # (tree/phylotree.cpp:462-472), not the leaf read first
        self.root = next(n for n in self.nodes if n.is_leaf() and n.id == 0)

    def _assign_slots(self):
        idx = 0
        stack = [(self.root, None)]
        while stack:
            node, dad = stack.pop()
            if dad is not None and not node.is_leaf():
                dad.find(node).slot = idx; idx += 1
            for it in reversed(node.neighbors):
                if it.node is not dad:
                    stack.append((it.node, node))

    # ---- invalidation -----------------------------------------------------------------------------------------
    def clear_all(self):
        for ab, ba in self.branches:
            ab.computed = ba.computed = False
            ab.size = ba.size = 0
        self.theta_computed = False
        self.current_it = self.current_it_back = None
        self.engine.theta_reset()

    @staticmethod
    def clear_reverse(node, dad):
        stack = [(node, dad)]
        while stack:
            nd, frm = stack.pop()
            for it in nd.neighbors:
                if it.node is not frm:
                    back = it.node.find(nd)
                    back.computed = False; back.size = 0
                    stack.append((it.node, nd))

    # ---- traversal --------------------------------------------------------------------------------------------
    def _size(self, nei, dad):
        if nei.size > 0:
            return nei.size
        s = 1
        for c in nei.node.neighbors:
            if c.node is not dad:
                s += self._size(c, nei.node)
        nei.size = s
        return s

    def _reorient(self, dad_branch, dad):
        if dad_branch.slot >= 0:
            return
        node = dad_branch.node
        for it in node.neighbors:
            if it.node is not dad:
                back = it.node.find(node)
                if back.slot >= 0:
                    dad_branch.slot, back.slot, back.computed = back.slot, -1, False
                    return
        raise RuntimeError("not re-oriented")

    def _traverse(self, dad_branch, dad, ops):
        import sys
        sys.setrecursionlimit(100000)
        node = dad_branch.node
        if dad_branch.computed or node.is_leaf():
            return
        kids = [it for it in node.neighbors if it.node is not dad]
        for it in kids:
            self._size(it, node)
        for it in sorted(kids, key=lambda x: -x.size):
            self._traverse(it, node, ops)
        dad_branch.computed = True
        self._reorient(dad_branch, dad)
        rec = np.zeros((), dtype=OP_DTYPE)
        rec["dst_slot"] = dad_branch.slot
        for name, it in zip(("left", "right"), kids):
            rec[name]["kind"] = 0 if it.node.is_leaf() else 1
            rec[name]["id"] = it.node.id if it.node.is_leaf() else it.slot
            rec[name]["length"] = it.length
        ops.append(rec)
        self.num_partial += 1

    def _ends(self, dad_branch, dad):
        node = dad_branch.node
        node_branch = node.find(dad)
        if node.is_leaf():
            dad, node = node, dad
            dad_branch, node_branch = node_branch, dad_branch
        ops = []
        self._traverse(dad_branch, dad, ops)
        self._traverse(node_branch, node, ops)
        ops = np.array(ops, dtype=OP_DTYPE) if ops else np.zeros(0, dtype=OP_DTYPE)
        a = (1, dad_branch.slot)
        b = (0, dad.id) if dad.is_leaf() else (1, node_branch.slot)
        return ops, a, b

    # ---- likelihood API ---------------------------------------------------------------------------------------
    def branch_logl(self, dad_branch, dad):
        length = dad_branch.length
        ops, a, b = self._ends(dad_branch, dad)
        if len(ops):
            self.theta_computed = False
        return self.engine.branch_logl(ops, a, b, length)

    def branch_derv(self, dad_branch, dad):
        length = dad_branch.length
        ops, a, b = self._ends(dad_branch, dad)
        if len(ops):
            self.theta_computed = False
        if not self.theta_computed:
            self.engine.theta_reset()
        r = self.engine.branch_derv(ops, a, b, length)
        self.theta_computed = True
        return r

    def from_buffer(self):
        return self.engine.logl_from_buffer(self.current_it.length)

    def farthest_leaf(self, node=None, dad=None):
        node = node or self.root
        if dad is not None and node.is_leaf():
            node.height = 0.0
            return node
        res = None
        node.height = 0.0
        for it in node.neighbors:
            if it.node is dad:
                continue
            leaf = self.farthest_leaf(it.node, node)
            if node.height < it.node.height + 1:
                node.height = it.node.height + 1
                res = leaf
        return res

    def compute_likelihood(self):
        if self.current_it is None:
            leaf = self.farthest_leaf()
            self.current_it = leaf.neighbors[0]
            self.current_it_back = self.current_it.node.find(leaf)
        return self.branch_logl(self.current_it, self.current_it_back.node)

    def pre_order_branches(self, node, dad, nodes, nodes2):
        if dad is not None:
            nodes.append(node); nodes2.append(dad)
        neivec = list(node.neighbors)
        for i1 in range(len(neivec)):
            for i2 in range(i1 + 1, len(neivec)):
                if neivec[i1].node.height > neivec[i2].node.height:
                    neivec[i1], neivec[i2] = neivec[i2], neivec[i1]
        for it in neivec:
            if it.node is not dad:
                self.pre_order_branches(it.node, node, nodes, nodes2)

    def _func_derv(self, value):
        self.current_it.length = self.current_it_back.length = value
        df, ddf = self.branch_derv(self.current_it, self.current_it_back.node)
        return -df, -ddf

    def minimize_newton(self, x1, xguess, x2, xacc, max_steps=100):
        rts = min(max(xguess, x1), x2)
        f, df = self._func_derv(rts)
        if df >= 0.0 and abs(f) < xacc:
            return rts
        if f < 0.0:
            xl, xh = rts, x2
        else:
            xh, xl = rts, x1
        dx = abs(xh - xl)
        for j in range(1, max_steps + 1):
            rts_old = rts
            if df <= 0.0 or ((rts - xh) * df - f) * ((rts - xl) * df - f) >= 0.0:
                dx = 0.5 * (xh - xl)
                rts = xl + dx
                if xl == rts:
                    return rts
            else:
                dx = f / df
                temp = rts
                rts -= dx
                if temp == rts:
                    return rts
            if abs(dx) < xacc or j == max_steps:
                return rts_old
            f, df = self._func_derv(rts)
            if df > 0.0 and abs(f) < xacc:
                return rts
            if f < 0.0:
                xl = rts
            elif f > 0.0:
                xh = rts
        return 0.0

    def optimize_one_branch(self, node1, node2, clear_lh=True, max_steps=100):
        self.current_it = node1.find(node2)
        self.current_it_back = node2.find(node1)
        cur = self.current_it.length
        self.theta_computed = False
        optx = self.minimize_newton(self.min_bl, cur, self.max_bl, self.min_bl, max_steps)
        if optx > self.max_bl * 0.95:
            self.current_it.length = self.current_it_back.length = optx
            opt_lh = self.from_buffer()
            self.current_it.length = self.current_it_back.length = cur
            if self.from_buffer() > opt_lh:
                optx = cur
        self.current_it.length = self.current_it_back.length = optx
        if clear_lh and cur != optx:
            self.clear_reverse(node1, node2)
            self.clear_reverse(node2, node1)

    def optimize_all_branches(self, iterations=100, tolerance=0.001, max_steps=100):
        far = self.farthest_leaf()
        self.farthest_leaf(far)
        nodes, nodes2 = [], []
        self.pre_order_branches(far, None, nodes, nodes2)
        tree_lh = self.branch_logl(nodes[0].find(nodes2[0]), nodes[0])
        for _ in range(iterations):
            saved = [ab.length for ab, _ in self.branches]
            for n1, n2 in zip(nodes, nodes2):
                self.optimize_one_branch(n1, n2, True, max_steps)
            new_lh = self.from_buffer()
            if new_lh < tree_lh - tolerance * 0.1:
                self.clear_all()
                for (ab, ba), ln in zip(self.branches, saved):
                    ab.length = ba.length = ln
                return self.compute_likelihood()
            if tree_lh <= new_lh <= tree_lh + tolerance:
                return new_lh
            tree_lh = new_lh
        return tree_lh

    def lengths(self):
        return np.array([ab.length for ab, _ in self.branches])
