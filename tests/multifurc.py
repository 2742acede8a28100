"""Test helper: collapse internal branches of a binary treesim.Tree into multifurcations and plan the traversal towards a
pendant branch with multi-child ops (b200lk_op.reserved = number of extra children, extra children in op order)."""
from __future__ import annotations

import numpy as np

from iqtree2_b200 import plan


def collapse(tree, branches):
    """adjacency of the tree with the given INTERNAL branches contracted: {node: [(neighbour, length)]} over surviving nodes;
    a contracted branch merges its inner end into its outer end (node ids of the survivors are kept)."""
    nt = tree.ntaxa
    adj = {u: [] for u in range(tree.nnodes)}
    for b, (u, v) in enumerate(tree.edges):
        adj[int(u)].append((int(v), float(tree.lengths[b])))
        adj[int(v)].append((int(u), float(tree.lengths[b])))
    for b in branches:
        u, v = int(tree.edges[b][0]), int(tree.edges[b][1])
        assert u >= nt and v >= nt, "only internal branches can be contracted"
        # find current representatives (earlier contractions may have removed u or v)
        while u not in adj or v not in adj:
            raise ValueError("contract disjoint branches only")
        keep, gone = u, v
        adj[keep] = [(w, l) for w, l in adj[keep] if w != gone]
        for w, l in adj[gone]:
            if w == keep:
                continue
            adj[keep].append((w, l))
            adj[w] = [(keep if x == gone else x, ll) for x, ll in adj[w]]
        del adj[gone]
    return adj


def newick(adj, names, root_leaf=0):
    def rec(node, dad):
        if len(adj[node]) == 1:
            return names[node]
        return "(" + ",".join(rec(w, node) + ":%.17g" % l for w, l in adj[node] if w != dad) + ")"
    (inner, l0), = adj[root_leaf]
    kids = [names[root_leaf] + ":%.17g" % l0] + [rec(w, inner) + ":%.17g" % l for w, l in adj[inner] if w != root_leaf]
    return "(" + ",".join(kids) + ");"


def plan_pendant(adj, ntaxa, leaf=0):
    """(ops, extra_children, endA, endB, length): full traversal towards the pendant branch of `leaf`; internal node k keeps
    CLV slot k - ntaxa; children in adjacency order."""
    (inner, length), = adj[leaf]
    ops, extra = [], []

    def child_rec(w, l):
        c = np.zeros((), dtype=plan.CHILD_DTYPE)
        c["kind"] = plan.KIND_TIP if w < ntaxa else plan.KIND_SLOT
        c["id"] = w if w < ntaxa else w - ntaxa
        c["length"] = l
        return c

    def visit(node, dad):
        kids = [(w, l) for w, l in adj[node] if w != dad]
        for w, l in kids:
            if w >= ntaxa:
                visit(w, node)
        rec = np.zeros((), dtype=plan.OP_DTYPE)
        rec["dst_slot"] = node - ntaxa
        rec["left"] = child_rec(*kids[0])
        rec["right"] = child_rec(*kids[1])
        rec["pad"] = len(kids) - 2
        for w, l in kids[2:]:
            extra.append(child_rec(w, l))
        ops.append(rec)

    visit(inner, leaf)
    ops = np.array(ops, dtype=plan.OP_DTYPE)
    extra = np.array(extra, dtype=plan.CHILD_DTYPE) if extra else np.zeros(0, dtype=plan.CHILD_DTYPE)
    return ops, extra, (plan.KIND_SLOT, inner - ntaxa), (plan.KIND_TIP, leaf), length
