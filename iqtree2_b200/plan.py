"""Host-side traversal planner (numpy): tree + evaluation branch -> post-order op list.

Mirrors what PhyloTree::computeTraversalInfo (tree/phylotree.cpp:5918-6028) hands to the kernels:
one entry per stale directed partial, children before parents, children visited in descending
subtree-size order (:5931-5942).  Slot policy is the reference's default LM_PER_NODE
(one CLV slot per internal node, tree/phylotree.cpp:1151-1157): internal node k owns slot k - ntaxa.

The op record layout is the C ABI's `b200lk_op` (include/b200lk.h).
"""
from __future__ import annotations

import numpy as np

CHILD_DTYPE = np.dtype([("kind", "<i4"), ("id", "<i4"), ("length", "<f8")])
OP_DTYPE = np.dtype([("dst_slot", "<i4"), ("pad", "<i4"), ("left", CHILD_DTYPE), ("right", CHILD_DTYPE)])
assert OP_DTYPE.itemsize == 40

KIND_TIP, KIND_SLOT = 0, 1
KIND_ROOT = 2          # non-reversible kernel: the virtual root leaf of a rooted tree (MTree::isRootLeaf, tree/mtree.h:135)
CHILD_TRANSPOSED = 4   # flag or-ed into a child's kind: the child's branch points TOWARD_ROOT (tree/phylokernelnew.h:939-951)


def subtree_sizes(tree):
    """size[(node, dad)] = number of nodes in the subtree of `node` seen from `dad` (PhyloNeighbor::size)."""
    adj = tree.adjacency()
    size = {}
    import sys
    sys.setrecursionlimit(max(10000, 4 * tree.nnodes))

    def rec(node, dad):
        key = (node, dad)
        if key in size:
            return size[key]
        s = 1
        for v, _ in adj[node]:
            if v != dad:
                s += rec(v, node)
        size[key] = s
        return s

    for u, v in tree.edges:
        rec(int(u), int(v))
        rec(int(v), int(u))
    return size


def plan_branch(tree, branch: int, lengths=None, sort_by_size: bool = True):
    """Full traversal towards `branch` = (u, v).

    Returns (ops, endA, endB, length): ops is a structured array (OP_DTYPE) in post-order covering both
    sides; endA/endB are (kind, id) of the two ends, with endA always a slot (the reference swaps so that
    a leaf end becomes `dad`, tree/phylokernelnew.h:2654-2661).
    """
    lengths = tree.lengths if lengths is None else lengths
    adj = tree.adjacency()
    nt = tree.ntaxa
    sizes = subtree_sizes(tree) if sort_by_size else None
    ops = []

    def visit(node, dad):
        # iterative post-order
        stack = [(node, dad, False)]
        while stack:
            nd, dd, done = stack.pop()
            if nd < nt:
                continue
            kids = [(v, b) for v, b in adj[nd] if v != dd]
            if done:
                rec = np.zeros((), dtype=OP_DTYPE)
                rec["dst_slot"] = nd - nt
                for name, (v, b) in zip(("left", "right"), kids):
                    rec[name]["kind"] = KIND_TIP if v < nt else KIND_SLOT
                    rec[name]["id"] = v if v < nt else v - nt
                    rec[name]["length"] = lengths[b]
                ops.append(rec)
                continue
            stack.append((nd, dd, True))
            order = kids
            if sizes is not None:
                order = sorted(kids, key=lambda vb: -sizes[(vb[0], nd)])
            # first-visited child must be popped first -> push in reverse
            for v, b in reversed(order):
                stack.append((v, nd, False))

    u, v = int(tree.edges[branch][0]), int(tree.edges[branch][1])
    visit(u, v)
    visit(v, u)
    ops = np.array(ops, dtype=OP_DTYPE) if ops else np.zeros(0, dtype=OP_DTYPE)

    def end(x):
        return (KIND_TIP, x) if x < nt else (KIND_SLOT, x - nt)

    a, b = end(u), end(v)
    if a[0] == KIND_TIP:
        a, b = b, a
    assert a[0] == KIND_SLOT, "a tree with a single branch between two leaves is not supported"
    return ops, a, b, float(lengths[branch])


def plan_branch_nonrev(tree, branch: int, lengths=None, rooted: bool = True):
    """plan_branch for the NON-REVERSIBLE kernel (tree/phylokernelnonrev.h).  `tree` has ntaxa leaves of which the LAST one is the
    virtual root leaf the reference attaches to the root of a rooted tree (MTree::readTree, tree/mtree.cpp:742-754: id = leafNum-1);
    the alignment has ntaxa - 1 rows.  Children whose branch points towards the root carry CHILD_TRANSPOSED, the root leaf is
    KIND_ROOT.  Returns (ops, endA, endB, length) with endA = dad_branch->partial_lh after the reference's swap
    (phylokernelnonrev.h:1071-1078: the root-ward end, or the inner end of a pendant branch) and endB = the other end (tip, root
    leaf or slot).  rooted=False: an unrooted tree evaluated with the non-reversible kernel (no root leaf, no direction)."""
    ops, a, b, L = plan_branch(tree, branch, lengths)
    if not rooted:
        return ops, a, b, L
    nt = tree.ntaxa
    rootleaf = nt - 1
    adj = tree.adjacency()
    # parent of every node in the orientation away from the root leaf
    parent = {rootleaf: -1}
    stack = [rootleaf]
    while stack:
        u = stack.pop()
        for v, _ in adj[u]:
            if v not in parent:
                parent[v] = u
                stack.append(v)
    for op in ops:
        nd = int(op["dst_slot"]) + nt
        for name in ("left", "right"):
            kind, cid = int(op[name]["kind"]), int(op[name]["id"])
            v = cid if kind == KIND_TIP else cid + nt
            if kind == KIND_TIP and cid == rootleaf:
                kind, cid = KIND_ROOT, 0
            if parent[nd] == v:
                kind |= CHILD_TRANSPOSED
            op[name]["kind"], op[name]["id"] = kind, cid
    u, v = int(tree.edges[branch][0]), int(tree.edges[branch][1])
    # (node, dad) after the swap: dad is a leaf if there is one, else the leaf-ward end
    if u == rootleaf or v == rootleaf:
        inner = v if u == rootleaf else u
        return ops, (KIND_SLOT, inner - nt), (KIND_ROOT, 0), L
    if u < nt or v < nt:
        leaf, inner = (u, v) if u < nt else (v, u)
        return ops, (KIND_SLOT, inner - nt), (KIND_TIP, leaf), L
    rootward, leafward = (u, v) if parent[v] == u else (v, u)
    return ops, (KIND_SLOT, rootward - nt), (KIND_SLOT, leafward - nt), L


def op_levels(ops):
    """Depth level of each op (0 = depends on tips only). Ops in one level are mutually independent;
    the levelled launch schedule runs level by level."""
    level_of_slot = {}
    levels = np.zeros(len(ops), dtype=np.int32)
    for k, op in enumerate(ops):
        lv = 0
        for name in ("left", "right"):
            if op[name]["kind"] == KIND_SLOT:
                lv = max(lv, level_of_slot[int(op[name]["id"])] + 1)
        levels[k] = lv
        level_of_slot[int(op["dst_slot"])] = lv
    return levels


def pendant_branch_of(tree, leaf: int = 0) -> int:
    for b, (u, v) in enumerate(tree.edges):
        if int(u) == leaf or int(v) == leaf:
            return b
    raise ValueError(leaf)


def count_node_cases(ops):
    """(N_ii, N_ti, N_tt) of SURVEY §8(d)'s flop formula."""
    lt = ops["left"]["kind"] == KIND_TIP
    rt = ops["right"]["kind"] == KIND_TIP
    ntt = int(np.sum(lt & rt))
    nti = int(np.sum(lt ^ rt))
    return len(ops) - ntt - nti, nti, ntt
