"""PhyloSuperTreePlen — Python face of the edge-linked partition model of the C++ host mirror (include/b200pt.h, b200st_*;
reference: tree/phylosupertreeplen.cpp, tree/phylosupertree.cpp).  One device engine per partition, all partitions
evaluated concurrently on their own CUDA streams.  No CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import SEQ_TYPES, EngineError, load_library


def _dp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class PhyloSuperTree:
    def __init__(self, names, num_threads=0):
        self.lib = load_library()
        L = self.lib
        L.b200st_create.restype = C.c_void_p
        L.b200st_last_error.restype = C.c_char_p
        L.b200st_last_error.argtypes = [C.c_void_p]
        L.b200st_destroy.argtypes = [C.c_void_p]
        L.b200st_launch_count.restype = C.c_int64
        L.b200st_launch_count.argtypes = [C.c_void_p]
        L.b200st_device_bytes.restype = C.c_int64
        L.b200st_device_bytes.argtypes = [C.c_void_p]
        self.names = list(names)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        h = L.b200st_create(len(names), arr, num_threads)
        if not h:
            raise EngineError(L.b200st_last_error(None).decode())
        self._h = C.c_void_p(h)
        self.nparts = 0

    def _check(self, rc):
        if rc < 0:
            raise EngineError(self.lib.b200st_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200st_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_partition(self, patterns, ptn_freq, model, part_rate=1.0, ptn_invar=None, device=0, safe_numeric=False, schedule=0):
        """patterns [ntaxa][nptn] in the super tree's taxon order (a taxon without data: a row of model.state_unknown)."""
        pat = np.ascontiguousarray(patterns, dtype=np.uint8)
        assert pat.shape[0] == len(self.names)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        freq = f(ptn_freq)
        inv = f(ptn_invar) if ptn_invar is not None else None
        arrs = [f(model.eval), f(model.evec), f(model.inv_evec), f(model.mix_weight), f(model.cat_rate), f(model.cat_prop)]
        rc = self.lib.b200st_add_partition(self._h, device, int(safe_numeric), schedule, _dp(pat), C.c_int64(pat.shape[1]), _dp(freq),
                                           _dp(inv), model.nstates, SEQ_TYPES[model.seq_type], model.state_unknown, model.ncat,
                                           model.nmix, int(model.fused_mix_rate), *[_dp(a) for a in arrs], C.c_double(part_rate))
        self._check(rc)
        self.nparts += 1
        return rc

    def read_tree(self, newick: str):
        self._check(self.lib.b200st_read_tree(self._h, newick.encode()))

    def get_tree(self) -> str:
        buf = C.create_string_buffer(1 << 22)
        self._check(self.lib.b200st_get_tree(self._h, buf, C.c_int64(len(buf))))
        return buf.value.decode()

    def branches(self):
        n = self.lib.b200st_num_branches(self._h)
        a = np.zeros(n, dtype=np.int32); b = np.zeros(n, dtype=np.int32); ln = np.zeros(n)
        self._check(self.lib.b200st_get_branches(self._h, _dp(a), _dp(b), _dp(ln)))
        return a, b, ln

    def set_branch_length(self, branch, length, clear_partials=True):
        self._check(self.lib.b200st_set_branch_length(self._h, branch, C.c_double(length), int(clear_partials)))

    def set_part_rate(self, part, rate):
        self._check(self.lib.b200st_set_part_rate(self._h, part, C.c_double(rate)))

    def clear_all_partial_lh(self):
        self._check(self.lib.b200st_clear_all_partial_lh(self._h))

    def compute_likelihood(self):
        out = C.c_double()
        self._check(self.lib.b200st_compute_likelihood(self._h, C.byref(out)))
        return out.value

    def compute_likelihood_branch(self, branch):
        out = C.c_double()
        self._check(self.lib.b200st_compute_likelihood_branch(self._h, branch, C.byref(out)))
        return out.value

    def compute_likelihood_derv(self, branch, length):
        df, ddf = C.c_double(), C.c_double()
        self._check(self.lib.b200st_compute_likelihood_derv(self._h, branch, C.c_double(length), C.byref(df), C.byref(ddf)))
        return df.value, ddf.value

    def optimize_one_branch(self, branch, clear_lh=True, max_nr_step=100):
        self._check(self.lib.b200st_optimize_one_branch(self._h, branch, int(clear_lh), max_nr_step))

    def optimize_all_branches(self, iterations=100, tolerance=0.001, max_nr_step=100):
        out = C.c_double()
        self._check(self.lib.b200st_optimize_all_branches(self._h, iterations, C.c_double(tolerance), max_nr_step, C.byref(out)))
        return out.value

    def get_best_nni_for_bran(self, branch, nni5=True, reverse=False):
        """NNI evaluation for the linked model: two dicts as PhyloTree.get_best_nni_for_bran (score, new_len, swap, outer)"""
        score, new_len = np.zeros(2), np.zeros((2, 5))
        sw, outer = np.zeros((2, 2), dtype=np.int32), np.zeros((2, 4), dtype=np.int32)
        self._check(self.lib.b200st_get_best_nni_for_bran(self._h, C.c_int32(int(branch)), C.c_int32(1 if reverse else 0),
                                                          C.c_int32(1 if nni5 else 0), _dp(score), _dp(new_len), _dp(sw), _dp(outer)))
        return [dict(move=m, reverse=bool(reverse), score=float(score[m]), new_len=new_len[m].tolist() if nni5 else new_len[m, :1].tolist(),
                     swap=(int(sw[m, 0]), int(sw[m, 1])), outer=[int(x) for x in outer[m]]) for m in range(2)]

    def do_nni(self, branch, move, new_len=None, clear_lh=True, reverse=False):
        nni5 = new_len is not None and len(new_len) == 5
        buf = None
        if new_len is not None:
            buf = np.zeros(5)
            buf[:len(new_len)] = new_len
        self._check(self.lib.b200st_do_nni(self._h, C.c_int32(int(branch)), C.c_int32(1 if reverse else 0), C.c_int32(int(move)),
                                           C.c_int32(1 if nni5 else 0), _dp(buf), C.c_int32(1 if clear_lh else 0)))

    def part_scores(self):
        s = np.zeros(self.nparts)
        self._check(self.lib.b200st_get_part_scores(self._h, _dp(s)))
        return s

    @property
    def launch_count(self):
        return int(self.lib.b200st_launch_count(self._h))

    @property
    def device_bytes(self):
        return int(self.lib.b200st_device_bytes(self._h))


class PhyloSuperTreeUnlinked:
    """The edge-unlinked partition model (`-Q` / `-sp`; the reference's PhyloSuperTree, tree/phylosupertree.cpp:653-791;
    include/b200pt.h, b200ut_*): a driver over independent iqtree2_b200.phylotree.PhyloTree objects, each with its own taxa,
    alignment, model, tree and branch lengths.  The partition trees are borrowed and kept alive here."""

    def __init__(self, num_threads=0):
        self.lib = load_library()
        L = self.lib
        L.b200ut_create.restype = C.c_void_p
        L.b200ut_last_error.restype = C.c_char_p
        L.b200ut_last_error.argtypes = [C.c_void_p]
        L.b200ut_destroy.argtypes = [C.c_void_p]
        L.b200ut_num_patterns.restype = C.c_int64
        L.b200ut_num_patterns.argtypes = [C.c_void_p]
        self._h = C.c_void_p(L.b200ut_create(num_threads))
        self.parts = []

    def _check(self, rc):
        if rc < 0:
            raise EngineError(self.lib.b200ut_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200ut_destroy(self._h)
            self._h = None
        self.parts = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_partition(self, tree):
        self._check(self.lib.b200ut_add_partition(self._h, tree._h))
        self.parts.append(tree)
        return len(self.parts) - 1

    def clear_all_partial_lh(self):
        self._check(self.lib.b200ut_clear_all_partial_lh(self._h))

    def compute_likelihood(self, want_pattern_lh=False):
        out = C.c_double()
        plh = np.zeros(self.lib.b200ut_num_patterns(self._h)) if want_pattern_lh else None
        self._check(self.lib.b200ut_compute_likelihood(self._h, C.byref(out), _dp(plh)))
        return (out.value, plh) if want_pattern_lh else out.value

    def compute_pattern_likelihood(self):
        out = C.c_double()
        plh = np.zeros(self.lib.b200ut_num_patterns(self._h))
        self._check(self.lib.b200ut_compute_pattern_likelihood(self._h, _dp(plh), C.byref(out)))
        return plh, out.value

    def optimize_all_branches(self, iterations=100, tolerance=0.001, max_nr_step=100):
        out = C.c_double()
        self._check(self.lib.b200ut_optimize_all_branches(self._h, C.c_int32(iterations), C.c_double(tolerance), C.c_int32(max_nr_step),
                                                          C.byref(out)))
        return out.value

    def part_scores(self):
        s = np.zeros(len(self.parts))
        self._check(self.lib.b200ut_get_part_scores(self._h, _dp(s)))
        return s
