"""Test helpers: key every directed CLV / branch by the leaf set it covers, so that slots of the engine interface,
nodes of the Python tree mirror and PhyloNeighbors of the compiled reference (oracle/ref_engine.py) can be matched."""
from __future__ import annotations


def slot_leafsets(tree, branch):
    """{slot: frozenset(leaves)} of the CLVs a full traversal towards `branch` computes (slot = node - ntaxa)."""
    adj = tree.adjacency()
    nt = tree.ntaxa
    out = {}

    def rec(node, dad):
        if node < nt:
            return frozenset([node])
        s = frozenset()
        for v, _ in adj[node]:
            if v != dad:
                s |= rec(v, node)
        out[node - nt] = s
        return s

    import sys
    sys.setrecursionlimit(max(10000, 4 * tree.nnodes))
    u, v = int(tree.edges[branch][0]), int(tree.edges[branch][1])
    rec(u, v)
    rec(v, u)
    return out


def branch_splits(tree):
    """[frozenset(side of branch b not containing taxon 0)] for every branch of a treesim.Tree."""
    adj = tree.adjacency()
    nt = tree.ntaxa
    allt = frozenset(range(nt))

    def side(node, dad):
        if node < nt:
            return frozenset([node])
        s = frozenset()
        for v, _ in adj[node]:
            if v != dad:
                s |= side(v, node)
        return s

    res = []
    for u, v in tree.edges:
        s = side(int(v), int(u))
        res.append(allt - s if 0 in s else s)
    return res


def pytree_splits(t):
    """Same for a tests/pytree.PyTree: one frozenset per entry of t.branches."""
    allt = frozenset(range(len(t.names)))

    def side(node, dad):
        if node.is_leaf():
            return frozenset([node.id])
        s = frozenset()
        for n in node.neighbors:
            if n.node is not dad:
                s |= side(n.node, node)
        return s

    res = []
    for ab, ba in t.branches:
        s = side(ab.node, ba.node)
        res.append(allt - s if 0 in s else s)
    return res


def pytree_leafset(node, dad):
    """leaf set of the subtree of `node` seen from `dad` in a tests/pytree.PyTree"""
    if node.is_leaf():
        return frozenset([node.id])
    s = frozenset()
    for n in node.neighbors:
        if n.node is not dad:
            s |= pytree_leafset(n.node, node)
    return s
