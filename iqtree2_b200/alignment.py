"""Alignment front end (host side): PHYLIP/FASTA reading, state encoding, site -> pattern compression.

Follows alignment/alignment.cpp: convertState (:1415-1500), pattern order = first occurrence (:1092-1178),
ptn_freq = pattern multiplicity, const_char for +I (computePtnInvar, tree/phylotreesse.cpp:549-612).
"""
from __future__ import annotations

import numpy as np

from .models import DNA_STATE_UNKNOWN, PROTEIN_STATE_UNKNOWN, PROTEIN_SYMBOLS

_DNA = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3, "R": 1 + 4 + 3, "Y": 2 + 8 + 3, "W": 1 + 8 + 3, "S": 2 + 4 + 3, "M": 1 + 2 + 3,
        "K": 4 + 8 + 3, "B": 2 + 4 + 8 + 3, "H": 1 + 2 + 8 + 3, "D": 1 + 4 + 8 + 3, "V": 1 + 2 + 4 + 3}


def convert_state(ch: str, seq_type: str) -> int:
    ch = ch.upper()
    if ch in "?-.~":
        return DNA_STATE_UNKNOWN if seq_type == "DNA" else PROTEIN_STATE_UNKNOWN
    if seq_type == "DNA":
        if ch in "ONX":
            return DNA_STATE_UNKNOWN
        if ch in _DNA:
            return _DNA[ch]
        raise ValueError(f"invalid DNA character {ch!r}")
    if ch == "B":
        return 20
    if ch == "Z":
        return 21
    if ch == "J":
        return 22
    if ch in "*UOX":
        return PROTEIN_STATE_UNKNOWN
    k = PROTEIN_SYMBOLS.find(ch)
    if k < 0:
        raise ValueError(f"invalid amino-acid character {ch!r}")
    return k


def read_phylip(path: str):
    """Sequential or interleaved PHYLIP. Returns (names, list of sequence strings)."""
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    ntaxa, nsites = (int(x) for x in lines[0].split()[:2])
    names, seqs = [], []
    body = lines[1:]
    for ln in body[:ntaxa]:
        parts = ln.split(None, 1)
        names.append(parts[0])
        seqs.append(parts[1].replace(" ", "") if len(parts) > 1 else "")
    k = 0
    for ln in body[ntaxa:]:
        seqs[k % ntaxa] += ln.replace(" ", "")
        k += 1
    assert all(len(s) == nsites for s in seqs), "sequence length mismatch"
    return names, seqs


def encode(seqs, seq_type="DNA") -> np.ndarray:
    lut = np.full(256, 255, dtype=np.uint8)
    for c in range(256):
        try:
            lut[c] = convert_state(chr(c), seq_type)
        except ValueError:
            pass
    arr = np.stack([lut[np.frombuffer(s.encode("ascii"), dtype=np.uint8)] for s in seqs])
    if (arr == 255).any():
        raise ValueError("alignment contains invalid characters")
    return arr


def compress(states: np.ndarray):
    """Distinct columns in first-occurrence order. Returns (patterns [ntaxa][nptn], ptn_freq)."""
    from .treesim import compress_patterns
    return compress_patterns(states)


def ptn_invar(patterns: np.ndarray, state_freq: np.ndarray, pinvar: float, nstates: int, state_unknown: int, seq_type="DNA"):
    """computePtnInvar (tree/phylotreesse.cpp:549-612)."""
    nptn = patterns.shape[1]
    out = np.zeros(nptn)
    if pinvar == 0.0:
        return out
    for p in range(nptn):
        col = patterns[:, p]
        # Pattern::const_char: the single state all non-gap characters share (ambiguity codes by intersection are
        # not reproduced here: patterns with ambiguity codes are treated as non-constant unless all entries agree)
        known = col[col != state_unknown]
        if known.size == 0:
            out[p] = pinvar
        elif (known == known[0]).all():
            s = int(known[0])
            if s < nstates:
                out[p] = pinvar * state_freq[s]
            elif seq_type == "DNA":
                c = s - nstates + 1
                out[p] = pinvar * sum(state_freq[x] for x in range(nstates) if c & (1 << x))
    return out
