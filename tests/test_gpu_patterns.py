"""GPU: site-pattern compression on the device (b200lk_compress_patterns) against the first-occurrence hash-map loop of
Alignment::addPattern (alignment/alignment.cpp:1092-1133) restated in numpy/python: same patterns in the same order, same
frequencies, same site -> pattern map (bit-exact); random alignments with heavy duplication, all-gap columns, a single
site, and a 1000-taxon x 200 000-site alignment where essentially every column is its own pattern."""
import numpy as np
import pytest

from iqtree2_b200 import models, treesim
from iqtree2_b200.engine import compress_patterns

pytestmark = pytest.mark.gpu


def reference_compress(states):
    """the reference's loop: a dict keyed by the column, patterns appended on first sight"""
    index, pats, freq = {}, [], []
    sp = np.zeros(states.shape[1], dtype=np.int32)
    cols = np.ascontiguousarray(states.T)
    for s in range(cols.shape[0]):
        key = cols[s].tobytes()
        k = index.get(key)
        if k is None:
            k = index[key] = len(pats)
            pats.append(cols[s]); freq.append(0)
        freq[k] += 1
        sp[s] = k
    return np.ascontiguousarray(np.array(pats).T), np.array(freq, dtype=np.float64), sp


@pytest.mark.parametrize("ntaxa,nsite,nstates,seed", [(5, 1, 4, 1), (4, 3000, 2, 2), (12, 20000, 4, 3), (30, 5000, 20, 4), (7, 777, 4, 5)])
def test_compress_matches_reference_loop(ntaxa, nsite, nstates, seed):
    rng = np.random.default_rng(seed)
    st = rng.integers(0, nstates, size=(ntaxa, nsite), dtype=np.uint8)
    st[:, rng.random(nsite) < 0.05] = 18          # all-gap columns stay patterns
    if nsite > 10:
        st[:, nsite // 2:] = st[:, rng.integers(0, nsite // 2, size=nsite - nsite // 2)]   # heavy duplication
    pat, freq, sp = compress_patterns(st)
    rp, rf, rs = reference_compress(st)
    assert pat.shape == rp.shape
    np.testing.assert_array_equal(pat, rp)
    np.testing.assert_array_equal(freq, rf)
    np.testing.assert_array_equal(sp, rs)
    np.testing.assert_array_equal(pat[:, sp], st)     # decompression reproduces the alignment
    assert freq.sum() == nsite


def test_compress_large_tree_evolved_alignment():
    tree = treesim.yule_tree(1000, seed=1)
    st = treesim.evolve_patterns_fast(tree, 4, 200_000, seed=2)
    st = np.concatenate([st, st[:, :50_000]], axis=1)     # a quarter of the sites repeat earlier columns
    pat, freq, sp = compress_patterns(st)
    np.testing.assert_array_equal(pat[:, sp], st)
    assert freq.sum() == st.shape[1] and pat.shape[1] <= 200_000
    first = np.full(pat.shape[1], st.shape[1], dtype=np.int64)
    np.minimum.at(first, sp, np.arange(st.shape[1]))
    assert np.all(np.diff(first) > 0)                     # first-occurrence order
    assert np.all(freq[sp[200_000:]] >= 2)
