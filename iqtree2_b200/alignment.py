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


_AMBI_AA = (4 + 8, 32 + 64, 512 + 1024)   # B = N|D, Z = Q|E, J = I|L as state bitmasks


def state_appearance(nstates: int, state_unknown: int, seq_type: str) -> np.ndarray:
    """Alignment::getAppearance (alignment/alignment.cpp:4769-4810) as a table: bitmask of states per tip code."""
    app = np.zeros(state_unknown + 1, dtype=np.uint64)
    full = (1 << nstates) - 1
    for s in range(state_unknown + 1):
        if s == state_unknown:
            app[s] = full
        elif s < nstates:
            app[s] = 1 << s
        elif seq_type == "DNA":
            app[s] = (s - (nstates - 1)) & full
        elif seq_type == "AA":
            app[s] = _AMBI_AA[s - 20] if s < 23 else 0
        else:
            app[s] = full
    return app


def const_char(patterns: np.ndarray, nstates: int, state_unknown: int, seq_type="DNA") -> np.ndarray:
    """Pattern::const_char as Alignment::computeConst sets it (alignment/alignment.cpp:925-1013): the intersection of
    the state sets of all characters of a pattern; STATE_UNKNOWN+1 = not constant, STATE_UNKNOWN = all gaps, a state,
    or the ambiguity code of the intersection."""
    app = state_appearance(nstates, state_unknown, seq_type)
    inter = np.bitwise_and.reduce(app[patterns], axis=0)
    nptn = patterns.shape[1]
    out = np.full(nptn, state_unknown + 1, dtype=np.int64)
    full = (1 << nstates) - 1
    for p in np.nonzero(inter)[0]:
        m = int(inter[p])
        if m == full:
            out[p] = state_unknown
        elif m & (m - 1) == 0:
            out[p] = m.bit_length() - 1
        elif seq_type == "DNA":
            out[p] = nstates - 1 + m
        elif seq_type == "AA":
            out[p] = nstates + (0 if (m >> 2) & 3 == 3 else 1 if (m >> 5) & 3 == 3 else 2)
    return out


def ptn_invar(patterns: np.ndarray, state_freq: np.ndarray, pinvar: float, nstates: int, state_unknown: int, seq_type="DNA"):
    """computePtnInvar (tree/phylotreesse.cpp:549-612): p_invar x frequency mass of the pattern's constant state set."""
    nptn = patterns.shape[1]
    out = np.zeros(nptn)
    if pinvar == 0.0:
        return out
    cc = const_char(patterns, nstates, state_unknown, seq_type)
    for p in np.nonzero(cc <= state_unknown)[0]:
        c = int(cc[p])
        if c == state_unknown:
            out[p] = pinvar
        elif c < nstates:
            out[p] = pinvar * state_freq[c]
        elif seq_type == "DNA":
            m = c - nstates + 1
            out[p] = pinvar * sum(state_freq[x] for x in range(nstates) if m & (1 << x))
        elif seq_type == "AA":
            m = _AMBI_AA[c - nstates]
            out[p] = pinvar * sum(state_freq[x] for x in range(11) if m & (1 << x))
    return out
