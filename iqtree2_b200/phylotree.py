"""PhyloTree — Python face of the C++ host mirror (include/b200pt.h): the tree-level likelihood API the reference's
callers use, over the CUDA engine.  No fallback: constructing a tree on a box without the CUDA library/GPU fails
at the first likelihood call."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .engine import SEQ_TYPES, EngineError, load_library


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class PhyloTree:
    def __init__(self, device=0, safe_numeric=False, force_norm=False, schedule=0, fused_newton=True):
        self.lib = load_library()
        self.lib.b200pt_create.restype = C.c_void_p
        self.lib.b200pt_last_error.restype = C.c_char_p
        self.lib.b200pt_last_error.argtypes = [C.c_void_p]
        self.lib.b200pt_destroy.argtypes = [C.c_void_p]
        self.lib.b200pt_num_partial_lh_computations.restype = C.c_int64
        self.lib.b200pt_num_partial_lh_computations.argtypes = [C.c_void_p]
        self._h = C.c_void_p(self.lib.b200pt_create(device, int(safe_numeric), int(force_norm), schedule, int(fused_newton)))
        self.nptn = 0
        self.ncat_mix = 0

    def _check(self, rc):
        if rc < 0:
            raise EngineError(self.lib.b200pt_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200pt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs -----------------------------------------------------------------------------------------------
    def set_alignment(self, names, patterns, ptn_freq, model):
        pat = np.ascontiguousarray(patterns, dtype=np.uint8)
        freq = np.ascontiguousarray(ptn_freq, dtype=np.float64)
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self.nptn = pat.shape[1]
        self._check(self.lib.b200pt_set_alignment(self._h, len(names), arr, _dp(pat), C.c_int64(self.nptn), _dp(freq),
                                                  model.nstates, SEQ_TYPES[model.seq_type], model.state_unknown))

    def set_pattern_invar(self, inv):
        inv = np.ascontiguousarray(inv, dtype=np.float64)
        self._check(self.lib.b200pt_set_pattern_invar(self._h, _dp(inv)))

    def set_ascertainment(self, asc_type, num_unobserved, nsite):
        self._check(self.lib.b200pt_set_ascertainment(self._h, asc_type, C.c_int64(num_unobserved), C.c_int64(nsite)))

    def set_robust_phylo(self, keep=1.0, median=False):
        self.lib.b200pt_set_robust_phylo.argtypes = [C.c_void_p, C.c_double, C.c_int32]
        self._check(self.lib.b200pt_set_robust_phylo(self._h, float(keep), int(bool(median))))

    def set_model(self, model):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        arrs = [f(model.eval), f(model.evec), f(model.inv_evec), f(model.mix_weight), f(model.cat_rate), f(model.cat_prop)]
        self.ncat_mix = model.ncat_mix
        self._check(self.lib.b200pt_set_model(self._h, model.ncat, model.nmix, int(model.fused_mix_rate), *[_dp(a) for a in arrs]))

    def read_tree(self, newick: str):
        self._check(self.lib.b200pt_read_tree(self._h, newick.encode()))

    def get_tree(self) -> str:
        buf = C.create_string_buffer(1 << 22)
        self._check(self.lib.b200pt_get_tree(self._h, buf, C.c_int64(len(buf))))
        return buf.value.decode()

    def branches(self):
        n = self.lib.b200pt_num_branches(self._h)
        a = np.zeros(n, dtype=np.int32); b = np.zeros(n, dtype=np.int32); ln = np.zeros(n)
        self._check(self.lib.b200pt_get_branches(self._h, _dp(a), _dp(b), _dp(ln)))
        return a, b, ln

    def set_branch_length(self, branch, length, clear_partials=True):
        self._check(self.lib.b200pt_set_branch_length(self._h, branch, C.c_double(length), int(clear_partials)))

    # ---- likelihood API ---------------------------------------------------------------------------------------
    def clear_all_partial_lh(self):
        self._check(self.lib.b200pt_clear_all_partial_lh(self._h))

    def compute_likelihood(self):
        out = C.c_double()
        self._check(self.lib.b200pt_compute_likelihood(self._h, C.byref(out)))
        return out.value

    def compute_likelihood_branch(self, branch):
        out = C.c_double()
        self._check(self.lib.b200pt_compute_likelihood_branch(self._h, branch, C.byref(out)))
        return out.value

    def compute_likelihood_derv(self, branch):
        df, ddf = C.c_double(), C.c_double()
        self._check(self.lib.b200pt_compute_likelihood_derv(self._h, branch, C.byref(df), C.byref(ddf)))
        return df.value, ddf.value

    def compute_likelihood_from_buffer(self):
        out = C.c_double()
        self._check(self.lib.b200pt_compute_likelihood_from_buffer(self._h, C.byref(out)))
        return out.value

    def compute_pattern_likelihood(self, with_cat=False):
        lh = np.zeros(self.nptn)
        cat = np.zeros((self.nptn, self.ncat_mix)) if with_cat else None
        self._check(self.lib.b200pt_compute_pattern_likelihood(self._h, _dp(lh), _dp(cat) if with_cat else None))
        return (lh, cat) if with_cat else lh

    def optimize_one_branch(self, branch, clear_lh=True, max_nr_step=100):
        self._check(self.lib.b200pt_optimize_one_branch(self._h, branch, int(clear_lh), max_nr_step))

    def get_best_nni_for_bran(self, branch, nni5=True, want_pattern_lh=False, reverse=False):
        """PhyloTree::getBestNNIForBran: the two NNIs around inner branch `branch`, evaluated on resident CLVs.  Two dicts:
        score, new_len (5 or 1), swap = (node id on node1's side, on node2's side), outer = node ids of new_len[1..4].
        node1 / node2 = ends a / b of branches(), or b / a with reverse=True"""
        score, new_len = np.zeros(2), np.zeros((2, 5))
        sw, outer = np.zeros((2, 2), dtype=np.int32), np.zeros((2, 4), dtype=np.int32)
        plh = np.zeros((2, self.nptn)) if want_pattern_lh else None
        self._check(self.lib.b200pt_get_best_nni_for_bran(self._h, C.c_int32(int(branch)), C.c_int32(1 if reverse else 0),
                                                           C.c_int32(1 if nni5 else 0), _dp(score),
                                                           _dp(new_len), sw.ctypes.data_as(C.c_void_p), outer.ctypes.data_as(C.c_void_p),
                                                           _dp(plh) if plh is not None else None))
        moves = [dict(move=m, reverse=bool(reverse), score=float(score[m]), new_len=new_len[m].tolist() if nni5 else new_len[m, :1].tolist(),
                      swap=(int(sw[m, 0]), int(sw[m, 1])), outer=[int(x) for x in outer[m]]) for m in range(2)]
        if plh is not None:
            for m in range(2):
                moves[m]["pattern_lh"] = plh[m]
        return moves

    def do_nni(self, branch, move, new_len=None, clear_lh=True, reverse=False):
        """PhyloTree::doNNI (+ changeNNIBrans when new_len is given: 5 lengths, or 1 = central branch only)"""
        nni5 = new_len is not None and len(new_len) == 5
        buf = None
        if new_len is not None:
            buf = np.zeros(5)
            buf[:len(new_len)] = new_len
        self._check(self.lib.b200pt_do_nni(self._h, C.c_int32(int(branch)), C.c_int32(1 if reverse else 0), C.c_int32(int(move)),
                                            C.c_int32(1 if nni5 else 0),
                                            _dp(buf) if buf is not None else None, C.c_int32(1 if clear_lh else 0)))

    def optimize_tree_length_scaling(self, min_scaling, scaling, max_scaling, gradient_epsilon=0.001):
        """PhyloTree::optimizeTreeLengthScaling: (logL, scaling)"""
        sc, ll = C.c_double(float(scaling)), C.c_double()
        self._check(self.lib.b200pt_optimize_tree_length_scaling(self._h, C.c_double(float(min_scaling)), C.byref(sc),
                                                                 C.c_double(float(max_scaling)), C.c_double(float(gradient_epsilon)),
                                                                 C.byref(ll)))
        return ll.value, sc.value

    def scale_length(self, norm):
        self._check(self.lib.b200pt_scale_length(self._h, C.c_double(float(norm))))

    def set_pipelined_sweep(self, on=True):
        """queue every optimizeAllBranches sweep as a whole on the device (one synchronisation per sweep)"""
        self._check(self.lib.b200pt_set_pipelined_sweep(self._h, int(on)))

    def optimize_all_branches(self, iterations=100, tolerance=0.001, max_nr_step=100):
        out = C.c_double()
        self._check(self.lib.b200pt_optimize_all_branches(self._h, iterations, C.c_double(tolerance), max_nr_step, C.byref(out)))
        return out.value

    # ---- -mem (LM_MEM_SAVE) and planning without execution ------------------------------------------------------
    def set_mem_save(self, on=True, max_mem_size=0.0):
        """Params::lh_mem_save / max_mem_size (tree/phylotree.cpp:999-1021): recycle a bounded number of CLV slots."""
        self._check(self.lib.b200pt_set_mem_save(self._h, int(on), C.c_double(max_mem_size)))

    def set_plan_only(self, on=True):
        self._check(self.lib.b200pt_set_plan_only(self._h, int(on)))

    @property
    def max_lh_slots(self):
        n = self.lib.b200pt_max_lh_slots(self._h)
        self._check(n)
        return n

    def plan_branch(self, branch):
        """(ops, (kind, id) of the node end, (kind, id) of the dad end, length): the traversal of
        computeLikelihoodBranch(branch); flags and slots are updated as if it had been executed."""
        from . import plan
        cap = 4 * max(self.lib.b200pt_num_branches(self._h), 4)
        ops = np.zeros(cap, dtype=plan.OP_DTYPE)
        nops, ns, dk, di = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        ln = C.c_double()
        self._check(self.lib.b200pt_plan_branch(self._h, branch, _dp(ops), cap, C.byref(nops), C.byref(ns), C.byref(dk),
                                                C.byref(di), C.byref(ln)))
        return ops[:nops.value].copy(), (plan.KIND_SLOT, ns.value), (dk.value, di.value), ln.value

    @property
    def num_partial_lh_computations(self):
        return int(self.lib.b200pt_num_partial_lh_computations(self._h))
