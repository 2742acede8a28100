#!/usr/bin/env python3
"""Generates tests/golden/ref_*.npz from the REAL reference kernel (oracle/_ref/libiqref.so = IQ-TREE 2.4.0's own TUs
compiled in place by oracle/build_ref.sh).  Run where /root/reference is mounted:

    python oracle/build_ref.sh && python tests/golden/make_golden.py

Each file holds the OUTPUTS of the reference for a seeded input that tests/cases.py regenerates (plus a digest of the
inputs, so that generator drift is detected instead of silently comparing different alignments):
  logl, df, ddf            computeLikelihoodBranch / computeLikelihoodDerv on the pendant branch of taxon 0
  pattern_lh               _pattern_lh
  scale_num                scale_num of every CLV of the traversal, stacked in planner (post-order) op order
  partial_digest           per-op max|partial_lh| per pattern is too large to commit; the first/last CLV are kept whole
  (config 1 only) opt_logl, opt_lengths (keyed by the split order of tests/refmap.pytree_splits)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from cases import make_case  # noqa: E402
from iqtree2_b200 import plan  # noqa: E402
from oracle import ref_engine  # noqa: E402
from refmap import pytree_splits, slot_leafsets  # noqa: E402

# name -> (kind, safe, kwargs of make_case)
GOLDEN_CASES = {
    "dna_norm": ("dna", False, dict(ntaxa=14, nsites=400)),
    "dna_safe": ("dna", True, dict(ntaxa=14, nsites=400)),
    "dna_inv": ("dna_inv", False, dict(ntaxa=14, nsites=400)),
    "dna_g8": ("dna_g8", False, dict(ntaxa=14, nsites=400)),
    "aa_norm": ("aa", False, dict(ntaxa=14, nsites=400)),
    "aa_safe": ("aa", True, dict(ntaxa=14, nsites=400)),
    "aa_c4mix": ("aa_c4", False, dict(ntaxa=12, nsites=200)),
    "aa_fused": ("aa_fused", False, dict(ntaxa=12, nsites=200)),
    "aa_c4mix_safe": ("aa_c4", True, dict(ntaxa=10, nsites=150)),
    "bin_safe": ("bin", True, dict(ntaxa=14, nsites=400)),
    "codon_safe": ("codon", True, dict(ntaxa=10, nsites=120)),
    "n5mix_safe": ("n5mix", True, dict(ntaxa=14, nsites=400)),
    "dna_deep_norm": ("dna", False, dict(ntaxa=500, nsites=40, tree_kind="caterpillar", mean_len=1.5, gaps=0.0, ambig=0.0)),
    "dna_deep_safe": ("dna", True, dict(ntaxa=500, nsites=40, tree_kind="caterpillar", mean_len=1.5, gaps=0.0, ambig=0.0)),
    "aa_deep_norm": ("aa", False, dict(ntaxa=200, nsites=40, tree_kind="caterpillar", mean_len=1.5, gaps=0.0, ambig=0.0)),
}


def input_digest(tree, model, pat, freq):
    h = hashlib.sha256()
    for a in (tree.edges, tree.lengths, pat, freq, model.eval, model.evec, model.inv_evec, model.cat_rate, model.cat_prop):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    assert ref_engine.available(), "build oracle/_ref/libiqref.so first (oracle/build_ref.sh)"
    for name, (kind, safe, kw) in GOLDEN_CASES.items():
        tree, model, pat, freq, inv = make_case(kind, **kw)
        r = ref_engine.RefTree(tree, model, pat, freq, safe_numeric=safe)
        br = plan.pendant_branch_of(tree, 0)
        ops, a, b, L = plan.plan_branch(tree, br)
        ll, plh, cat = r.branch_logl(0, want_cat=True)
        ref = r.directed_partials()
        sets = slot_leafsets(tree, br)
        scale = np.stack([ref[sets[int(op["dst_slot"])]][1] for op in ops])
        first, last = ref[sets[int(ops[0]["dst_slot"])]][0], ref[sets[int(ops[-1]["dst_slot"])]][0]
        df, ddf = r.branch_derv(0)
        th, bs = r.theta()
        out = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(out, digest=input_digest(tree, model, pat, freq), logl=ll, df=df, ddf=ddf, pattern_lh=plh,
                            pattern_lh_cat=cat, scale_num=scale, partial_first=first, partial_last=last,
                            buffer_scale_all=bs, ptn_invar=r.ptn_invar(), logl_from_buffer=r.logl_from_buffer())
        print(f"{name}: logL={ll:.6f} df={df:.6e} max scale={int(scale.max())} -> {os.path.getsize(out)} bytes")
    # BASELINE config 1: example.phy + committed NJ tree, GTR+G4; fixed-tree logL and optimizeAllBranches
    from test_tree_logic_cpu import example_case, make_pytree
    names, pat, freq, model, newick = example_case()
    r = ref_engine.RefTree(None, model, pat, freq, newick=newick, names=names)
    ll, plh = r.compute_likelihood()
    opt = r.optimize_all_branches(100, 0.001, 100)
    ref_len = r.branches()
    t = make_pytree(names, pat, freq, model, newick)
    lens = np.array([ref_len[s] for s in pytree_splits(t)])
    out = os.path.join(HERE, "ref_config1_example.npz")
    np.savez_compressed(out, logl=ll, pattern_lh=plh, opt_logl=opt, opt_lengths=lens)
    print(f"config1: logL={ll:.6f} optimised logL={opt:.6f} -> {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main()
