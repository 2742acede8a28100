"""Test infrastructure: a pure-Python restatement of the reference's tree-level likelihood logic (PhyloNeighbor flags,
slot re-orientation, traversal order, Newton-Raphson branch optimisation) over ANY engine with the engine interface —
in the tests, over the CPU oracle.  The C++ host mirror over the CUDA engine is checked against it.

Follows tree/phylotree.cpp:1233-1274 (computeLikelihood), :2594-2774 (optimizeOneBranch / optimizeAllBranches),
:3975-4067 (doNNI / changeNNIBrans), :4069-4322 (getBestNNIForBran), :5898-6028 (reorientPartialLh /
computeTraversalInfo), tree/mtree.cpp:3040-3082, utils/optimization.cpp:422-499.
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

    # ---- tree-length scaling (tree/phylotree.cpp:2455-2528; Brent: utils/optimization.cpp:185-373) ---------------------
    def scale_length(self, norm):
        for ab, ba in self.branches:
            ab.length *= norm
            ba.length *= norm

    def _scaling_function(self, value):
        if value != self.current_scaling:
            self.scale_length(value / self.current_scaling)
            self.current_scaling = value
            self.clear_all()
        return -self.compute_likelihood()

    @staticmethod
    def brent_opt(func, ax, bx, cx, tol, fax, fbx, fcx):
        ITMAX, CGOLD, ZEPS = 100, 0.3819660, 1.0e-10
        sign = lambda a, b: abs(a) if b >= 0.0 else -abs(a)
        d = e = 0.0
        a, b = min(ax, cx), max(ax, cx)
        x, fx = bx, fbx
        if fax < fcx:
            w, fw, v, fv = ax, fax, cx, fcx
        else:
            w, fw, v, fv = cx, fcx, ax, fax
        for _ in range(ITMAX):
            xm = 0.5 * (a + b)
            tol1 = tol * abs(x) + ZEPS
            tol2 = 2.0 * tol1
            if abs(x - xm) <= (tol2 - 0.5 * (b - a)):
                return x, fx
            if abs(e) > tol1:
                r = (x - w) * (fx - fv)
                q = (x - v) * (fx - fw)
                p = (x - v) * q - (x - w) * r
                q = 2.0 * (q - r)
                if q > 0.0:
                    p = -p
                q = abs(q)
                etemp, e = e, d
                if abs(p) >= abs(0.5 * q * etemp) or p <= q * (a - x) or p >= q * (b - x):
                    e = a - x if x >= xm else b - x
                    d = CGOLD * e
                else:
                    d = p / q
                    u = x + d
                    if u - a < tol2 or b - u < tol2:
                        d = sign(tol1, xm - x)
            else:
                e = a - x if x >= xm else b - x
                d = CGOLD * e
            u = x + d if abs(d) >= tol1 else x + sign(tol1, d)
            fu = func(u)
            if fu <= fx:
                if u >= x:
                    a = x
                else:
                    b = x
                v, w, x = w, x, u
                fv, fw, fx = fw, fx, fu
            else:
                if u < x:
                    a = u
                else:
                    b = u
                if fu <= fw or w == x:
                    v, w, fv, fw = w, u, fw, fu
                elif fu <= fv or v == x or v == w:
                    v, fv = u, fu
        return x, fx

    def minimize_one_dimen(self, func, xmin, xguess, xmax, tolerance):
        xguess = min(max(xguess, xmin), xmax)
        eps = xguess * tolerance * 50.0
        ax = max(xguess - eps, xmin)
        bx = xguess
        cx = min(xguess + eps, xmax)
        fb = func(bx)
        fa = func(ax) if ax < bx else fb
        fc = func(cx) if bx < cx else fb
        if fa < fb or fc < fb:
            if ax != xmin:
                fa = func(xmin)
            if cx != xmax:
                fc = func(xmax)
            ax, cx = xmin, xmax
        optx, fx = self.brent_opt(func, ax, bx, cx, tolerance, fa, fb, fc)
        if fx > fb:
            func(bx)
            return bx
        return optx

    def optimize_tree_length_scaling(self, min_scaling, scaling, max_scaling, gradient_epsilon):
        """returns (logL, scaling); every branch length ends up multiplied by scaling / (the scaling passed in)"""
        self.current_scaling = scaling
        lens = self.lengths()
        max_brlen = max(self.min_bl, lens.max())
        min_brlen = min(self.max_bl, lens.min())
        if min_brlen <= 0.0:
            min_brlen = self.min_bl
        max_scaling = min(max_scaling, 10.0 * self.max_bl / max_brlen)
        min_scaling = max(min_scaling, 0.1 * self.min_bl / min_brlen)
        new = self.minimize_one_dimen(self._scaling_function, min(scaling, min_scaling), scaling, max(max_scaling, scaling),
                                      max(0.001, gradient_epsilon))
        if new != self.current_scaling:
            self.scale_length(new / self.current_scaling)
            self.current_scaling = new
            self.clear_all()
        return self.compute_likelihood(), new

    # ---- NNI (tree/phylotree.cpp:3975-4322) ---------------------------------------------------------------------
    NNI_MAX_NR_STEP = 10   # utils/pllnni.cpp:42

    @staticmethod
    def _copy_nei(nei, slot):
        """PhyloNeighbor::newNeighbor (tree/phylonode.h:78-95): same node / length / id / size, nothing computed"""
        c = Nei(nei.node, nei.length, nei.id)
        c.size, c.slot = nei.size, slot
        return c

    def get_best_nni_for_bran(self, node1, node2, nni5=True, nni5_num_eval=1):
        """returns two moves: dict(score, new_len[5 or 1], swap=(node1 side node, node2 side node), outer=[4 nodes in
        newLen[1..4] order]); the tree, its lengths and every flag of the original neighbours are as before on return"""
        assert not node1.is_leaf() and not node2.is_leaf() and len(node1.neighbors) == 3 and len(node2.neighbors) == 3
        # the six (nni5) or two neighbour objects that the evaluation may overwrite are replaced by copies on the two spare slots
        saved = [(node1, node1.neighbors.index(node1.find(node2))), (node2, node2.neighbors.index(node2.find(node1)))]
        if nni5:
            for it in node1.neighbors:
                if it.node is not node2:
                    saved.append((it.node, it.node.neighbors.index(it.node.find(node1))))
            for it in node2.neighbors:
                if it.node is not node1:
                    saved.append((it.node, it.node.neighbors.index(it.node.find(node2))))
        else:
            self._reorient(node1.find(node2), node1)
            self._reorient(node2.find(node1), node2)
        nslots = sum(1 for ab, ba in self.branches for x in (ab, ba) if x.slot >= 0)   # the T-2 regular slots
        saved_nei, mem_id = [], 0
        for owner, k in saved:
            old = owner.neighbors[k]
            saved_nei.append(old)
            slot = -1
            if old.slot >= 0:
                slot = nslots + mem_id; mem_id += 1
            owner.neighbors[k] = self._copy_nei(old, slot)
        assert mem_id == 2
        node12, node21 = node1.find(node2), node2.find(node1)
        k1 = next(k for k, it in enumerate(node1.neighbors) if it.node is not node2)
        k2s = [k for k, it in enumerate(node2.neighbors) if it.node is not node1]
        moves = []
        for k2 in k2s:
            node1_nei, node2_nei = node1.neighbors[k1], node2.neighbors[k2]
            self._reorient(node12, node1)
            self._reorient(node21, node2)
            # the swap (Node::updateNeighbor)
            node1.neighbors[k1] = node2_nei
            node2_nei.node.find(node2).node = node1
            node2.neighbors[k2] = node1_nei
            node1_nei.node.find(node1).node = node2
            new_len = [0.0] * 5
            for _ in range(nni5_num_eval):
                node12.computed = node21.computed = False
                i = 1
                if nni5:
                    for it in list(node1.neighbors):
                        if it.node is node2:
                            continue
                        it.node.find(node1).computed = False
                        self.optimize_one_branch(node1, it.node, False, self.NNI_MAX_NR_STEP)
                        new_len[i] = node1.find(it.node).length; i += 1
                    node21.computed = False
                self.optimize_one_branch(node1, node2, False, self.NNI_MAX_NR_STEP)
                new_len[0] = node12.length
                if nni5:
                    for it in list(node2.neighbors):
                        if it.node is node1:
                            continue
                        it.node.find(node2).computed = False
                        self.optimize_one_branch(node2, it.node, False, self.NNI_MAX_NR_STEP)
                        new_len[i] = node2.find(it.node).length; i += 1
                    node12.computed = False
            score = self.from_buffer()
            outer = [it.node for it in node1.neighbors if it.node is not node2] + [it.node for it in node2.neighbors if it.node is not node1]
            moves.append(dict(score=score, new_len=new_len if nni5 else new_len[:1], swap=(node1_nei.node, node2_nei.node), outer=outer,
                              k1=k1, k2=k2))
            self._reorient(node12, node1)
            self._reorient(node21, node2)
            # swap back
            node1.neighbors[k1] = node1_nei
            node1_nei.node.find(node2).node = node1
            node2.neighbors[k2] = node2_nei
            node2_nei.node.find(node1).node = node2
        for (owner, k), old in reversed(list(zip(saved, saved_nei))):
            if owner.neighbors[k] is self.current_it:
                self.current_it = old
            if owner.neighbors[k] is self.current_it_back:
                self.current_it_back = old
            owner.neighbors[k] = old
        # lengths of the four outer branches: the outward neighbours were optimised, their saved partners hold the old values
        for a, b in ((node1, node2), (node2, node1)):
            for it in a.neighbors:
                if it.node is not b:
                    it.length = it.node.find(a).length
        return moves

    def do_nni(self, node1, node2, move, clear_lh=True):
        """PhyloTree::doNNI (:3975-4036) with the neighbour positions of a move from get_best_nni_for_bran, then
        changeNNIBrans (:4038-4067) when the move carries lengths"""
        k1, k2 = move["k1"], move["k2"]
        self._reorient(node1.find(node2), node1)
        self._reorient(node2.find(node1), node2)
        node1_nei, node2_nei = node1.neighbors[k1], node2.neighbors[k2]
        node1.neighbors[k1] = node2_nei
        node2_nei.node.find(node2).node = node1
        node2.neighbors[k2] = node1_nei
        node1_nei.node.find(node1).node = node2
        nei12, nei21 = node1.find(node2), node2.find(node1)
        if clear_lh:
            nei12.computed = nei21.computed = False
            nei12.size = nei21.size = 0
            self.clear_reverse(node2, node1)
            self.clear_reverse(node1, node2)
        new_len = move.get("new_len")
        if new_len:
            nei12.length = nei21.length = new_len[0]
            if len(new_len) == 5:
                i = 1
                for a, b in ((node1, node2), (node2, node1)):
                    for it in a.neighbors:
                        if it.node is not b:
                            it.length = it.node.find(a).length = new_len[i]; i += 1
