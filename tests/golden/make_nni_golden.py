#!/usr/bin/env python3
"""Mints tests/golden/ref_config1_nni.json from the REAL reference (oracle/_ref/libiqref.so): PhyloTree::getBestNNIForBran
(tree/phylotree.cpp:4069-4322, nni5 and central-branch-only) on every inner branch of BASELINE config 1 (example.phy, the
committed NJ tree, GTR+G4).  Per branch (keyed by its split = the side without taxon 0) and per move (keyed by the two swapped
subtrees): NNIMove::newloglh and NNIMove::newLen with the subtree each outer length leads to.  Run here, where /root/reference
exists; the JSON travels with the repo."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from oracle import ref_engine
    from refmap import pytree_splits
    from test_tree_logic_cpu import example_case, make_pytree
    names, pat, freq, model, newick = example_case()
    t = make_pytree(names, pat, freq, model, newick)          # only for the list of inner splits
    spl = pytree_splits(t)
    inner = [spl[ab.id] for ab, ba in t.branches if not ab.node.is_leaf() and not ba.node.is_leaf()]
    out = {"tree": newick, "cases": []}
    for nni5 in (True, False):
        r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
        logl = r.compute_likelihood()[0]
        for s in inner:
            moves = r.best_nni(s, nni5=nni5)
            out["cases"].append({"nni5": nni5, "split": sorted(s), "logl_before": logl,
                                 "moves": [{"swap": [sorted(m["swap"][0]), sorted(m["swap"][1])], "score": m["score"],
                                            "new_len": m["new_len"], "outer": [sorted(x) for x in m["outer"]]} for m in moves]})
        assert abs(r.compute_likelihood()[0] - logl) < 1e-9
        r.close()
    with open(os.path.join(HERE, "ref_config1_nni.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(out["cases"]), "branch evaluations written")


if __name__ == "__main__":
    main()
