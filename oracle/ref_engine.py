"""ctypes wrapper around oracle/_ref/libiqref.so — TEST INFRASTRUCTURE.

The library is the UNMODIFIED reference likelihood path (IQ-TREE 2.4.0 TUs compiled in place by oracle/build_ref.sh)
behind oracle/ref_harness.cpp.  Used to pin oracle/lk_oracle.c and the CUDA engine against the real kernel (tests/)
and as bench.py's CPU baseline ("kind": "reference").  Never imported by the product package.  /root/reference is NOT
needed at run time: the prebuilt .so travels with the gpurun snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libiqref.so")
SEQ_TYPES = {"GENERIC": 0, "DNA": 1, "AA": 2}


class _Desc(C.Structure):
    _fields_ = [("nstates", C.c_int32), ("ncat", C.c_int32), ("nmix", C.c_int32), ("fused_mix_rate", C.c_int32),
                ("safe_numeric", C.c_int32), ("seq_type", C.c_int32), ("state_unknown", C.c_int32), ("ntaxa", C.c_int32),
                ("nptn", C.c_int64), ("nthreads", C.c_int32), ("reserved", C.c_int32),
                ("states", C.c_void_p), ("ptn_freq", C.c_void_p), ("newick", C.c_char_p),
                ("eval", C.c_void_p), ("evec", C.c_void_p), ("inv_evec", C.c_void_p), ("mix_weight", C.c_void_p),
                ("state_freq", C.c_void_p), ("cat_rate", C.c_void_p), ("cat_prop", C.c_void_p), ("pinvar", C.c_double)]


def available() -> bool:
    return os.path.exists(LIB)


def build() -> bool:
    """(Re)build when /root/reference is mounted; otherwise keep whatever prebuilt library is there."""
    subprocess.call(["bash", os.path.join(_HERE, "build_ref.sh")])
    return available()


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(LIB + " (run oracle/build_ref.sh where /root/reference is mounted)")
        _lib = C.CDLL(LIB)
        _lib.iqref_create.restype = C.c_void_p
        for f in ("iqref_optimize_one_branch", "iqref_branch_logl", "iqref_compute_likelihood", "iqref_logl_from_buffer", "iqref_time_eval",
                  "iqref_optimize_all_branches"):
            getattr(_lib, f).restype = C.c_double
    return _lib


LIB_B200 = os.path.join(os.path.dirname(LIB), "libiqref_b200.so")
_lib_b200 = None


def available_b200() -> bool:
    return os.path.exists(LIB_B200)


def lib_b200():
    """the reference + oracle/ref_b200_shim.cpp linked against iqtree2_b200/libb200lk.so (own copy of the reference's globals)"""
    global _lib_b200
    if _lib_b200 is None:
        if not available_b200():
            raise FileNotFoundError(LIB_B200 + " (run oracle/build_ref.sh after building iqtree2_b200/libb200lk.so)")
        _lib_b200 = C.CDLL(LIB_B200)
        _lib_b200.iqref_create.restype = C.c_void_p
        for f in ("iqref_optimize_one_branch", "iqref_branch_logl", "iqref_compute_likelihood", "iqref_logl_from_buffer", "iqref_time_eval",
                  "iqref_optimize_all_branches"):
            getattr(_lib_b200, f).restype = C.c_double
    return _lib_b200


def set_mem_save(on: bool, max_mem_size: float = 0.0):
    """Params::lh_mem_save / max_mem_size (-mem) for the trees created afterwards."""
    lib().iqref_set_mem_save(int(bool(on)), C.c_double(float(max_mem_size)))


def set_asc_next(asc_type: int, unobserved, b200=False):
    """+ASC for the next RefTree (of the B200-bound library with b200=True): ModelFactory::ASC_type (1 = ASC_VARIANT, Lewis;
    2 = ASC_VARIANT_MISSING, Holder) and unobserved_ptns [nunobs][ntaxa]"""
    u = np.ascontiguousarray(unobserved, dtype=np.uint8)
    (lib_b200() if b200 else lib()).iqref_set_asc_next(C.c_int(asc_type), u.ctypes.data_as(C.c_void_p), C.c_int(u.shape[0]), C.c_int(u.shape[1]))


def set_robust(keep: float = 1.0, median: bool = False):
    """Params::robust_phy_keep / robust_median for the trees created afterwards; (1.0, False) = off"""
    L = lib()
    L.iqref_set_robust.argtypes = [C.c_double, C.c_int]
    L.iqref_set_robust.restype = None
    L.iqref_set_robust(float(keep), int(bool(median)))


def max_threads() -> int:
    return int(lib().iqref_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def rooted_newick(tree, fmt="%.17g"):
    """Newick of a treesim.Tree whose LAST leaf is the virtual root leaf: written from the node the root leaf hangs on (two
    children = a rooted tree for MTree::readTree), leaves named by their alignment rows"""
    import sys
    sys.setrecursionlimit(max(10000, 4 * tree.nnodes))
    adj = tree.adjacency()
    root = tree.ntaxa - 1
    top = adj[root][0][0]

    def rec(node, dad):
        if node < tree.ntaxa:
            return "%d" % node
        return "(" + ",".join(rec(v, node) + ":" + (fmt % tree.lengths[b]) for v, b in adj[node] if v != dad) + ")"

    return "(" + ",".join(rec(v, top) + ":" + (fmt % tree.lengths[b]) for v, b in adj[top] if v != root) + ");"


class RefTree:
    """One reference PhyloTree (alignment patterns + tree + harness-supplied eigensystem)."""

    def __init__(self, tree, model, pat, freq, safe_numeric=False, nthreads=1, newick=None, names=None, b200=False):
        """tree: treesim.Tree, or None with `newick` + `names` (leaf names replaced by their alignment rows).
        b200=True: the same reference tree with its kernel pointers bound to the product's C ABI (oracle/ref_b200_shim.cpp,
        libiqref_b200.so; needs a GPU): every member above the kernels is still the reference's own code."""
        self._b200 = bool(b200)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.model, self.tree = model, tree
        self.ntaxa, self.nptn = int(pat.shape[0]), int(pat.shape[1])
        self.n, self.Cmix = model.nstates, model.ncat_mix
        self.B = self.n * self.Cmix
        self.safe = bool(safe_numeric)
        self._keep = [np.ascontiguousarray(pat, dtype=np.uint8), np.ascontiguousarray(np.rint(freq), dtype=np.int32),
                      f64(model.eval), f64(model.evec), f64(model.inv_evec), f64(model.mix_weight), f64(model.state_freq),
                      f64(model.cat_rate), f64(model.cat_prop)]
        assert np.allclose(self._keep[1], freq), "the reference stores integer pattern frequencies"
        if newick is not None:
            import re
            row = {nm: str(i) for i, nm in enumerate(names)}
            nwk = re.sub(r"(?<=[(,])\s*([^():,;\s]+)(?=\s*[:,)])", lambda m: row[m.group(1)], newick)
        elif getattr(model, "nonrev", None) and model.nonrev["rooted"]:
            # rooted tree: `tree` has ntaxa + 1 leaves, the last one the virtual root leaf; the Newick is written from the node it
            # hangs on, with a bifurcating root, and MTree::readTree attaches the root leaf itself (tree/mtree.cpp:742-754)
            assert tree.ntaxa == self.ntaxa + 1
            nwk = rooted_newick(tree)
        else:
            old = tree.names
            tree.names = ["%d" % i for i in range(self.ntaxa)]   # readTreeString takes alignment row ids as leaf names
            try:
                nwk = tree.to_newick(fmt="%.17g")
            finally:
                tree.names = old
        st = getattr(model, "site", None)
        if st:   # site-specific model: one eigensystem per pattern
            b = [f64(st["eval"]), f64(st["evec"]), f64(st["inv_evec"])]
            self._L().iqref_set_site_next(_p(b[0]), _p(b[1]), _p(b[2]), C.c_int64(self.nptn), C.c_int(self.n))
        ml = getattr(model, "mixlen_factor", None)
        if ml is not None:   # mixed branch lengths: a MixlenTree whose branches get lengths[c] = length * factor[c]
            fct = f64(ml)
            assert fct.size == model.ncat
            self._nmixlen = int(fct.size)
            self._L().iqref_set_mixlen_next(_p(fct), C.c_int(fct.size))
        nr = getattr(model, "nonrev", None)
        if nr:
            a = [f64(nr["ceval"]), f64(nr["cevec"]), f64(nr["cinv_evec"]), f64(nr["rate_matrix"])]
            self._L().iqref_set_nonrev_next(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), C.c_int(self.n), C.c_int(model.nmix))
        k = self._keep
        d = _Desc(self.n, model.ncat, model.nmix, int(model.fused_mix_rate), int(self.safe), SEQ_TYPES[model.seq_type],
                  model.state_unknown, self.ntaxa, self.nptn, int(nthreads), 0, _p(k[0]), _p(k[1]), nwk.encode(),
                  _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]), _p(k[6]), _p(k[7]), _p(k[8]), float(model.pinvar))
        if self._b200:
            self._L().iqref_b200_enable(1)
        self._h = C.c_void_p(self._L().iqref_create(C.byref(d)))
        if not self._h:
            raise RuntimeError("iqref_create failed")
        if self._b200:
            L = self._L()
            L.iqref_b200_enable(0)
            L.iqref_tree_ptr.restype = C.c_void_p
            self._tree_ptr = C.c_void_p(L.iqref_tree_ptr(self._h))
            if L.iqref_b200_attach(self._tree_ptr) != 0:
                raise RuntimeError("iqref_b200_attach failed")

    def _L(self):
        return lib_b200() if self._b200 else lib()

    def b200_set_fused_newton(self, on=True):
        """optimizeOneBranch of the bound tree through b200lk_newton_branch (one C call per branch) instead of the reference's
        host-side minimizeNewton loop over the derivative kernel pointer"""
        if self._L().iqref_b200_set_fused_newton(self._tree_ptr, int(bool(on))) != 0:
            raise RuntimeError("not a B200-bound tree")

    def b200_kernel_calls(self):
        L = self._L()
        L.iqref_b200_kernel_calls.restype = C.c_longlong
        L.iqref_b200_launches.restype = C.c_longlong
        return int(L.iqref_b200_kernel_calls(self._tree_ptr)), int(L.iqref_b200_launches(self._tree_ptr))

    def close(self):
        if getattr(self, "_h", None):
            self._L().iqref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:   # interpreter shutdown: module globals may already be gone
            pass

    def info(self):
        out = (C.c_int32 * 6)()
        self._L().iqref_info(self._h, out)
        return dict(vector_size=out[0], safe_numeric=bool(out[1]), num_threads=out[2], num_packets=out[3], kernel=out[4],
                    partial_computations=out[5])

    def clear_all_partial_lh(self):
        self._L().iqref_clear_all_partial_lh(self._h)

    def branch_logl(self, leaf=0, want_cat=False):
        """computeLikelihoodBranch on the pendant branch of alignment row `leaf` -> (logL, pattern_lh[, pattern_lh_cat])."""
        plh = np.empty(self.nptn)
        cat = np.empty((self.nptn, self.Cmix)) if want_cat else None
        ll = self._L().iqref_branch_logl(self._h, int(leaf), _p(plh), _p(cat) if want_cat else None)
        return (ll, plh, cat) if want_cat else (ll, plh)

    def compute_likelihood(self):
        plh = np.empty(self.nptn)
        return self._L().iqref_compute_likelihood(self._h, _p(plh)), plh

    def branch_derv(self, leaf=0):
        df, ddf = C.c_double(), C.c_double()
        self._L().iqref_branch_derv(self._h, int(leaf), C.byref(df), C.byref(ddf))
        return df.value, ddf.value

    def branch_derv_mixlen(self, leaf=0):
        """computeLikelihoodDerv of a mixlen tree -> (gradient [K], hessian [K][K], logL)"""
        K = self._nmixlen
        df, ddf = np.zeros(K + 1), np.zeros((K, K))
        self._L().iqref_branch_derv_mixlen(self._h, int(leaf), _p(df), _p(ddf))
        return df[:K].copy(), ddf, float(df[K])

    def branch_derv_mixlen_cur(self, leaf, cur):
        df, ddf = C.c_double(), C.c_double()
        self._L().iqref_branch_derv_mixlen_cur(self._h, int(leaf), int(cur), C.byref(df), C.byref(ddf))
        return df.value, ddf.value

    def logl_from_buffer(self):
        return self._L().iqref_logl_from_buffer(self._h)

    def theta(self):
        th, bs = np.empty((self.nptn, self.B)), np.empty(self.nptn)
        self._L().iqref_get_theta(self._h, _p(th), _p(bs))
        return th, bs

    def tip_partial_lh(self):
        out = np.empty((self.model.state_unknown + 1, self.model.nmix, self.n))
        self._L().iqref_get_tip_partial_lh(self._h, _p(out))
        return out

    def ptn_invar(self):
        out = np.empty(self.nptn)
        self._L().iqref_get_ptn_invar(self._h, _p(out))
        return out

    def directed_partials(self, want_partial=True):
        """{frozenset(leaves of the subtree): (partial_lh [nptn][B] or None, scale_num)} for every computed directed CLV."""
        res = {}
        S = self.Cmix if self.safe else 1
        for k in range(self._L().iqref_num_directed(self._h)):
            mask = np.zeros(self.ntaxa, dtype=np.uint8)
            part = np.empty((self.nptn, self.B)) if want_partial else None
            sc = np.empty((self.nptn, S), dtype=np.uint16)
            if self._L().iqref_get_directed(self._h, k, _p(mask), _p(part) if want_partial else None, _p(sc)):
                res[frozenset(np.nonzero(mask)[0].tolist())] = (part, sc)
        return res

    def optimize_all_branches(self, iterations=100, tolerance=0.001, max_nr_step=100):
        return self._L().iqref_optimize_all_branches(self._h, C.c_int(int(iterations)), C.c_double(float(tolerance)), C.c_int(int(max_nr_step)))

    def optimize_one_branch(self, leaf, max_nr_step=100):
        return self._L().iqref_optimize_one_branch(self._h, C.c_int(int(leaf)), C.c_int(int(max_nr_step)))

    def best_traversal(self):
        """the branch order optimizeAllBranches walks, as splits (side not containing taxon 0)"""
        nb = 2 * self.ntaxa - 3
        masks = np.zeros((nb, self.ntaxa), dtype=np.uint8)
        k = self._L().iqref_best_traversal(self._h, _p(masks))
        return [frozenset(np.nonzero(masks[b])[0].tolist()) for b in range(k)]

    def branches(self):
        """{frozenset(side not containing taxon 0): length}"""
        nb = 2 * self.ntaxa   # (2T - 3 branches; a rooted tree has the root leaf's branch on top: 2T - 1)
        masks, lens = np.zeros((nb, self.ntaxa), dtype=np.uint8), np.zeros(nb)
        k = self._L().iqref_get_branches(self._h, _p(masks), _p(lens))
        return {frozenset(np.nonzero(masks[b])[0].tolist()): float(lens[b]) for b in range(k)}

    def pattern_lh_cat_wsl(self, wsl, leaf=0):
        """PhyloTree::computePatternLhCat(wsl) at the pendant branch of `leaf`: [nptn][getNumLhCat(wsl)]"""
        L = self._L()
        K = L.iqref_pattern_lh_cat_wsl(self._h, C.c_int(int(leaf)), C.c_int(int(wsl)), None)
        out = np.zeros((self.nptn, K))
        L.iqref_pattern_lh_cat_wsl(self._h, C.c_int(int(leaf)), C.c_int(int(wsl)), _p(out))
        return out

    def optimize_tree_length_scaling(self, min_scaling, scaling, max_scaling, gradient_epsilon):
        """PhyloTree::optimizeTreeLengthScaling: (logL, scaling)"""
        L = self._L()
        L.iqref_optimize_tree_length_scaling.restype = C.c_double
        sc = C.c_double(float(scaling))
        ll = L.iqref_optimize_tree_length_scaling(self._h, C.c_double(float(min_scaling)), C.byref(sc), C.c_double(float(max_scaling)),
                                                  C.c_double(float(gradient_epsilon)))
        return ll, sc.value

    def best_nni(self, split, nni5=True):
        """PhyloTree::getBestNNIForBran on the inner branch with that split (side not containing taxon 0): two dicts with
        score, new_len, swap = (leaf sets of the two swapped subtrees), outer = leaf sets of the branches of new_len[1..4]"""
        nt = self.ntaxa
        mask = np.zeros(nt, dtype=np.uint8)
        mask[sorted(split)] = 1
        score, new_len = np.zeros(2), np.zeros((2, 5))
        sw, outer = np.zeros((2, 2, nt), dtype=np.uint8), np.zeros((2, 4, nt), dtype=np.uint8)
        L = self._L()
        L.iqref_best_nni.restype = C.c_int
        rc = L.iqref_best_nni(self._h, _p(mask), C.c_int(1 if nni5 else 0), _p(score), _p(new_len), _p(sw), _p(outer))
        if rc != 0:
            raise ValueError("no inner branch with that split")
        fs = lambda m: frozenset(np.nonzero(m)[0].tolist())
        return [dict(score=float(score[m]), new_len=new_len[m].tolist() if nni5 else new_len[m, :1].tolist(),
                     swap=(fs(sw[m, 0]), fs(sw[m, 1])), outer=[fs(outer[m, k]) for k in range(4)]) for m in range(2)]

    def rell(self, pattern_lh, boot):
        """IQTree::saveCurrentTree's RELL scores through the tree's own dotProduct member (dotProductSIMD<float, Vec8f>)"""
        plh = np.ascontiguousarray(pattern_lh, dtype=np.float64)
        bs = np.ascontiguousarray(boot, dtype=np.float32)
        out = np.zeros(bs.shape[0], dtype=np.float32)
        rc = self._L().iqref_rell(self._h, _p(plh), _p(bs), C.c_int(plh.size), C.c_int(bs.shape[0]), _p(out))
        if rc != 0:
            raise RuntimeError("the reference tree has no dotProduct kernel bound")
        return out

    def time_eval(self, leaf=0, reps=3):
        ll, df, ddf, mean = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        best = self._L().iqref_time_eval(self._h, int(leaf), int(reps), C.byref(ll), C.byref(df), C.byref(ddf), C.byref(mean))
        return dict(best_sec=best, mean_sec=mean.value, logl=ll.value, df=df.value, ddf=ddf.value)

    @property
    def max_lh_slots(self):
        return int(self._L().iqref_max_lh_slots(self._h))

    def last_traversal_size(self):
        """partial updates the last branch_logl / branch_derv call scheduled (traversal_info.size())"""
        return int(self._L().iqref_last_traversal_size(self._h))
