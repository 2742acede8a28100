"""Shared seeded test cases: (tree, model, tip patterns, frequencies, invariant-site terms)."""
from __future__ import annotations

import numpy as np

from iqtree2_b200 import models, treesim


def ptn_invar_for(model, pat):
    """computePtnInvar (tree/phylotreesse.cpp:549-612) with Pattern::const_char from Alignment::computeConst; +I uses
    the mixture-averaged state frequencies (getStateFrequency(.., -1))."""
    from iqtree2_b200 import alignment
    freq = (model.mix_weight[:, None] * model.state_freq.reshape(model.nmix, model.nstates)).sum(axis=0)
    return alignment.ptn_invar(pat, freq, model.pinvar, model.nstates, model.state_unknown, model.seq_type)


def make_case(kind="dna", ntaxa=12, nsites=300, seed=1, gaps=0.05, ambig=0.03, tree_kind="yule", mean_len=0.1, model=None):
    if model is not None:
        pass   # the caller's model
    elif kind == "dna":
        model = models.make_gtr_gamma()
    elif kind == "dna_inv":
        model = models.make_gtr_gamma(pinvar=0.2)
    elif kind == "dna_g1":
        model = models.make_gtr_gamma(ncat=1)
    elif kind == "dna_g8":
        model = models.make_gtr_gamma(ncat=8, alpha=0.3)
    elif kind == "aa":
        model = models.make_lg_gamma()
    elif kind == "aa_c4":  # small LG+C60-style mixture: 4 classes x 4 rates
        model = models.make_lg_c60_gamma(nclasses=4)
    elif kind == "aa_c60":  # the full LG+C60+G4 of BASELINE config 4: 60 classes x 4 rates = 240 categories, block 4800
        model = models.make_lg_c60_gamma()
    elif kind == "aa_fused":  # 4 classes fused with the 4 rate categories (LG4X-style; ModelFactory::fused_mix_rate)
        model = models.make_lg_c60_gamma(nclasses=4)
        model.fused_mix_rate = True
    elif kind == "bin":
        model = models.make_random_reversible(2, ncat=4, seed=seed)
    elif kind == "n5mix":
        model = models.make_random_reversible(5, ncat=3, nmix=2, seed=seed)
    elif kind == "codon":
        model = models.make_random_reversible(61, ncat=1, seed=seed)
    elif kind == "n7":
        model = models.make_random_reversible(7, ncat=2, seed=seed)
    elif kind == "codon_g4":   # 61 states x 4 rate categories: the FP64-MMA walker for general state counts, 8 state tiles
        model = models.make_random_reversible(61, ncat=4, seed=seed)
    elif kind == "n12mix":     # 12 states, 2 classes x 3 rates: 2 state tiles, mixture
        model = models.make_random_reversible(12, ncat=3, nmix=2, seed=seed)
    elif kind == "n23":        # 23 states: 3 state tiles (odd tile count), one category
        model = models.make_random_reversible(23, ncat=1, seed=seed)
    elif kind == "n36g2":      # 36 states x 2 categories: 5 state tiles
        model = models.make_random_reversible(36, ncat=2, seed=seed)
    else:
        raise ValueError(kind)
    if tree_kind == "yule":
        tree = treesim.yule_tree(ntaxa, seed=seed, mean_len=mean_len)
    else:
        tree = treesim.caterpillar_tree(ntaxa, seed=seed, mean_len=mean_len)
    aln = treesim.simulate_alignment(tree, model, nsites, seed=seed + 1)
    pat, freq = treesim.compress_patterns(aln)
    rng = np.random.default_rng(seed + 2)
    pat = pat.copy()
    if gaps > 0:
        pat[rng.random(pat.shape) < gaps] = model.state_unknown
    if ambig > 0 and model.seq_type == "DNA":
        m = rng.random(pat.shape) < ambig
        pat[m] = rng.integers(4, 18, size=int(m.sum()))
    if ambig > 0 and model.seq_type == "AA":
        m = rng.random(pat.shape) < ambig
        pat[m] = rng.integers(20, 23, size=int(m.sum()))
    inv = ptn_invar_for(model, pat)
    return tree, model, pat, freq, inv


def load_engine(cls, tree, model, pat, freq, inv, safe=False, **kw):
    e = cls(model, tree.ntaxa, pat.shape[1], safe_numeric=safe, **kw)
    e.set_tip_states(pat)
    e.set_pattern_freq(freq)
    e.set_pattern_invar(inv)
    return e
