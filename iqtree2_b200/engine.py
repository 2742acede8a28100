"""ctypes binding of the C ABI in include/b200lk.h (libb200lk.so) — the product path.

There is no fallback: if the CUDA library is missing or no GPU is visible, constructing a
LikelihoodEngine raises.  Method names follow the C ABI, which follows the reference's PhyloTree
kernel entry points (tree/phylotree.h:899-1048, 1340-1346).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

ABI_SYMBOLS = [
    "b200lk_abi_version", "b200lk_create", "b200lk_destroy", "b200lk_last_error", "b200lk_device_bytes",
    "b200lk_set_stream", "b200lk_set_schedule", "b200lk_set_tip_states", "b200lk_set_pattern_freq",
    "b200lk_set_pattern_invar", "b200lk_set_model", "b200lk_set_rates", "b200lk_update_partials",
    "b200lk_branch_logl", "b200lk_branch_derv", "b200lk_theta_reset", "b200lk_logl_from_buffer",
    "b200lk_newton_branch", "b200lk_get_pattern_lh", "b200lk_get_pattern_scaling", "b200lk_get_pattern_lh_cat",
    "b200lk_get_theta", "b200lk_get_buffer_scale_all", "b200lk_get_scale_num", "b200lk_get_partial",
    "b200lk_get_tip_partial_lh", "b200lk_pattern_lh_dev", "b200lk_set_profiling", "b200lk_last_timing",
    "b200lk_launch_count", "b200lk_set_reference_quirks",
    "b200lk_enqueue_branch_logl", "b200lk_enqueue_branch_derv", "b200lk_enqueue_logl_from_buffer", "b200lk_finish",
    "b200lk_set_boot_samples", "b200lk_rell", "b200lk_set_extra_children", "b200lk_set_ascertainment", "b200lk_set_robust_phylo", "b200lk_em_posterior",
    "b200lk_set_pattern_freq_from_posterior",
    "b200lk_enqueue_newton_branch", "b200lk_sweep_finish", "b200lk_compress_patterns",
    "b200lk_group_create", "b200lk_group_destroy", "b200lk_group_last_error", "b200lk_group_enqueue_branch_logl",
    "b200lk_group_enqueue_branch_derv", "b200lk_group_enqueue_logl_from_buffer", "b200lk_group_finish", "b200lk_group_launch_count",
    "b200lk_enqueue_results_dev", "b200lk_boot_reset", "b200lk_rell_update", "b200lk_get_boot_state", "b200lk_set_model_nonrev", "b200lk_set_site_model", "b200lk_set_mixlen", "b200lk_branch_derv_mixlen", "b200lk_branch_derv_mixlen_cur",
]

SCHEDULE_AUTO, SCHEDULE_LEVELLED, SCHEDULE_WALKER = 0, 1, 2
SEQ_TYPES = {"GENERIC": 0, "DNA": 1, "AA": 2}


class Config(C.Structure):
    _fields_ = [("nstates", C.c_int32), ("ncat", C.c_int32), ("nmix", C.c_int32), ("fused_mix_rate", C.c_int32),
                ("safe_numeric", C.c_int32), ("seq_type", C.c_int32), ("state_unknown", C.c_int32),
                ("ntaxa", C.c_int32), ("nptn", C.c_int64), ("nslots", C.c_int32), ("device", C.c_int32)]


_lib = None


def load_library(path: str | None = None):
    """Load libb200lk.so (building it in-tree when sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _build.LIB
    if path is None:
        try:
            _build.build()
        except Exception:
            if not os.path.exists(p):
                raise
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -m iqtree2_b200.build` (needs nvcc); "
                           "the likelihood engine has no CPU fallback")
    lib = C.CDLL(p)
    lib.b200lk_last_error.restype = C.c_char_p
    lib.b200lk_last_error.argtypes = [C.c_void_p]
    lib.b200lk_device_bytes.restype = C.c_int64
    lib.b200lk_device_bytes.argtypes = [C.c_void_p]
    lib.b200lk_launch_count.restype = C.c_int64
    lib.b200lk_launch_count.argtypes = [C.c_void_p]
    lib.b200lk_pattern_lh_dev.restype = C.c_void_p
    lib.b200lk_pattern_lh_dev.argtypes = [C.c_void_p]
    lib.b200lk_destroy.restype = None
    lib.b200lk_destroy.argtypes = [C.c_void_p]
    if path is None:
        _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class EngineError(RuntimeError):
    pass


class LikelihoodEngine:
    """One PhyloTree's likelihood buffers on one GPU for one pattern range (see include/b200lk.h)."""

    def __init__(self, model, ntaxa, nptn, nslots=None, safe_numeric=False, device=0, schedule=SCHEDULE_AUTO):
        self._h = None
        self.lib = load_library()
        self.model = model
        self.n, self.ncat, self.nmix = model.nstates, model.ncat, model.nmix
        self.Cmix = model.ncat_mix
        self.B = self.n * self.Cmix
        self.ntaxa, self.nptn = int(ntaxa), int(nptn)
        self.nslots = int(nslots if nslots is not None else max(ntaxa - 2, 1) + 2)
        self.safe = bool(safe_numeric)
        self.device = int(device)
        cfg = Config(self.n, self.ncat, self.nmix, int(model.fused_mix_rate), int(self.safe), SEQ_TYPES[model.seq_type],
                     model.state_unknown, self.ntaxa, self.nptn, self.nslots, self.device)
        h = C.c_void_p()
        rc = self.lib.b200lk_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EngineError(f"b200lk_create failed ({rc}): {self.lib.b200lk_last_error(None).decode()}")
        self._h = h
        self.last_status = 0
        if schedule:
            self.set_schedule(schedule)
        self.set_model(model)

    # ---- plumbing -----------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise EngineError(f"b200lk error {rc}: {self.lib.b200lk_last_error(self._h).decode()}")
        self.last_status = rc
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self.lib.b200lk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self):
        return int(self.lib.b200lk_device_bytes(self._h))

    @property
    def launch_count(self):
        return int(self.lib.b200lk_launch_count(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.b200lk_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_schedule(self, schedule):
        self._check(self.lib.b200lk_set_schedule(self._h, C.c_int(schedule)))

    def set_reference_quirks(self, mask):
        self._check(self.lib.b200lk_set_reference_quirks(self._h, C.c_int32(int(mask))))

    def set_profiling(self, on=True):
        self._check(self.lib.b200lk_set_profiling(self._h, C.c_int32(int(on))))

    def last_timing(self):
        ms = (C.c_float * 5)()
        launches = (C.c_int32 * 3)()
        self.lib.b200lk_last_timing(self._h, ms, launches)
        return {"partials_ms": ms[0], "root_ms": ms[1], "derv_ms": ms[2], "total_ms": ms[3], "partials_kernel_ms": ms[4],
                "launches": [launches[0], launches[1], launches[2]]}

    # ---- inputs -------------------------------------------------------------------------------
    def set_model(self, model):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        ev, U, Ui, w = f(model.eval), f(model.evec), f(model.inv_evec), f(model.mix_weight)
        assert ev.size == self.nmix * self.n and U.size == self.nmix * self.n * self.n
        self._check(self.lib.b200lk_set_model(self._h, _dp(ev), _dp(U), _dp(Ui), _dp(w)))
        r, p = f(model.cat_rate), f(model.cat_prop)
        self._check(self.lib.b200lk_set_rates(self._h, _dp(r), _dp(p)))
        st = getattr(model, "site", None)
        if st:   # site-specific model (ModelSet / PMSF): one eigensystem per pattern
            b = [f(st["eval"]), f(st["evec"]), f(st["inv_evec"])]
            assert b[0].shape == (self.nptn, self.n) and b[1].shape == (self.nptn, self.n, self.n) and b[2].shape == b[1].shape
            self._check(self.lib.b200lk_set_site_model(self._h, _dp(b[0]), _dp(b[1]), _dp(b[2])))
        nr = getattr(model, "nonrev", None)
        if nr:   # non-reversible kernel (tree/phylokernelnonrev.h): complex eigensystem, rate matrices, state frequencies
            a = [f(nr["ceval"]), f(nr["cevec"]), f(nr["cinv_evec"]), f(nr["rate_matrix"]), f(model.state_freq)]
            self._check(self.lib.b200lk_set_model_nonrev(self._h, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), _dp(w), C.c_int32(int(nr["rooted"]))))

    def set_tip_states(self, states):
        s = np.ascontiguousarray(states, dtype=np.uint8)
        assert s.shape == (self.ntaxa, self.nptn), (s.shape, self.ntaxa, self.nptn)
        self._check(self.lib.b200lk_set_tip_states(self._h, _dp(s)))

    def set_pattern_freq(self, freq):
        f = np.ascontiguousarray(freq, dtype=np.float64)
        assert f.shape == (self.nptn,)
        self._check(self.lib.b200lk_set_pattern_freq(self._h, _dp(f)))

    def set_pattern_invar(self, invar):
        f = np.ascontiguousarray(invar, dtype=np.float64)
        assert f.shape == (self.nptn,)
        self._check(self.lib.b200lk_set_pattern_invar(self._h, _dp(f)))

    # ---- the kernel entry points --------------------------------------------------------------
    @staticmethod
    def _ops(ops):
        ops = np.ascontiguousarray(ops)
        assert ops.dtype.itemsize == 40 or len(ops) == 0
        return ops

    def update_partials(self, ops):
        ops = self._ops(ops)
        self._check(self.lib.b200lk_update_partials(self._h, _dp(ops), C.c_int32(len(ops))))

    def branch_logl(self, ops, end_a, end_b, length, save_log_value=True):
        ops = self._ops(ops)
        assert end_a[0] == 1, "the first branch end must be a CLV slot"
        out = C.c_double()
        self._check(self.lib.b200lk_branch_logl(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]),
                                                C.c_int32(end_b[0]), C.c_int32(end_b[1]), C.c_double(length),
                                                C.c_int32(int(save_log_value)), C.byref(out)))
        return out.value

    # ---- asynchronous forms: queue several calls, wait once (include/b200lk.h) ------------------------------
    def enqueue_branch_logl(self, ops, end_a, end_b, length, save_log_value=True):
        ops = self._ops(ops)
        self._check(self.lib.b200lk_enqueue_branch_logl(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]), C.c_int32(end_b[0]),
                                                        C.c_int32(end_b[1]), C.c_double(length), C.c_int32(int(save_log_value))))

    def enqueue_branch_derv(self, ops, end_a, end_b, length):
        ops = self._ops(ops)
        self._check(self.lib.b200lk_enqueue_branch_derv(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]), C.c_int32(end_b[0]),
                                                        C.c_int32(end_b[1]), C.c_double(length)))

    def enqueue_results_dev(self, dev_ptr):
        """queue a copy of (logL, df, ddf) into three doubles of device memory (e.g. a torch CUDA tensor's data_ptr())"""
        self._check(self.lib.b200lk_enqueue_results_dev(self._h, C.c_void_p(int(dev_ptr))))

    def finish(self):
        ll, df, ddf = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.b200lk_finish(self._h, C.byref(ll), C.byref(df), C.byref(ddf)))
        return ll.value, df.value, ddf.value

    def theta_reset(self):
        self._check(self.lib.b200lk_theta_reset(self._h))

    def branch_derv(self, ops, end_a, end_b, length):
        ops = self._ops(ops)
        df, ddf = C.c_double(), C.c_double()
        self._check(self.lib.b200lk_branch_derv(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]),
                                                C.c_int32(end_b[0]), C.c_int32(end_b[1]), C.c_double(length),
                                                C.byref(df), C.byref(ddf)))
        return df.value, ddf.value

    # ---- mixed branch lengths (heterotachy) ---------------------------------------------------------
    def set_mixlen(self, table):
        """table [nbranches][ncat] of per-class branch lengths (None / empty: off); afterwards every `length` is a ROW of it"""
        t = np.ascontiguousarray(table if table is not None else np.zeros((0, 1)), dtype=np.float64)
        self._check(self.lib.b200lk_set_mixlen(self._h, _dp(t), C.c_int32(t.shape[0])))
        self._nmixlen = t.shape[1] if t.shape[0] else 0

    def branch_derv_mixlen(self, ops, end_a, end_b, row):
        """computeLikelihoodDerv of a mixlen tree -> (gradient [K], hessian [K][K], logL)"""
        ops = self._ops(ops)
        K = self._nmixlen
        df, ddf = np.zeros(K + 1), np.zeros((K, K))
        self._check(self.lib.b200lk_branch_derv_mixlen(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]), C.c_int32(end_b[0]),
                                                       C.c_int32(end_b[1]), C.c_double(row), _dp(df), _dp(ddf)))
        return df[:K].copy(), ddf, float(df[K])

    def branch_derv_mixlen_cur(self, ops, end_a, end_b, row, cur):
        """computeLikelihoodDervMixlen for class `cur` -> (df, ddf)"""
        ops = self._ops(ops)
        df, ddf = C.c_double(), C.c_double()
        self._check(self.lib.b200lk_branch_derv_mixlen_cur(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]), C.c_int32(end_b[0]),
                                                           C.c_int32(end_b[1]), C.c_double(row), C.c_int32(cur), C.byref(df), C.byref(ddf)))
        return df.value, ddf.value

    def logl_from_buffer(self, length):
        out = C.c_double()
        self._check(self.lib.b200lk_logl_from_buffer(self._h, C.c_double(length), C.byref(out)))
        return out.value

    def newton_branch(self, ops, end_a, end_b, min_len, guess, max_len, tolerance, max_steps=100):
        ops = self._ops(ops)
        out = C.c_double()
        steps = C.c_int32()
        self._check(self.lib.b200lk_newton_branch(self._h, _dp(ops), C.c_int32(len(ops)), C.c_int32(end_a[1]),
                                                  C.c_int32(end_b[0]), C.c_int32(end_b[1]), C.c_double(min_len),
                                                  C.c_double(guess), C.c_double(max_len), C.c_double(tolerance),
                                                  C.c_int32(max_steps), C.byref(out), C.byref(steps)))
        return out.value, steps.value

    # ---- read-backs (reference V=1 layouts) ---------------------------------------------------
    def _get(self, fn, n, dtype=np.float64, *args):
        out = np.empty(n, dtype=dtype)
        self._check(getattr(self.lib, fn)(self._h, *args, _dp(out)))
        return out

    def get_pattern_lh(self):
        return self._get("b200lk_get_pattern_lh", self.nptn)

    def set_ascertainment(self, asc_type, num_unobserved, nsite):
        """+ASC (1 = Lewis, 2 = Holder's +ASC_MIS): the last num_unobserved patterns are the unobserved constant patterns
        (Holder: nstates groups of num_unobserved/nstates), nsite = aln->getNSite()"""
        self._check(self.lib.b200lk_set_ascertainment(self._h, C.c_int32(asc_type), C.c_int64(num_unobserved), C.c_int64(nsite)))

    def em_posterior(self, wsl, add_invar, want_posterior=False):
        """E-step of RateFree / ModelMixture::optimizeWithEM on the categories of computePatternLhCat(wsl) of the last
        branch_logl (2 = RATECAT, 3 = MIXTURE, 4 = MIXTURE_RATECAT): new_prop[K] (and posterior[nptn][K])"""
        K = C.c_int32()
        self._check(self.lib.b200lk_em_posterior(self._h, C.c_int32(wsl), C.c_int32(int(bool(add_invar))), None, None, C.byref(K)))
        prop = np.zeros(K.value)
        post = np.zeros((self.nptn, K.value)) if want_posterior else None
        self._check(self.lib.b200lk_em_posterior(self._h, C.c_int32(wsl), C.c_int32(int(bool(add_invar))), _dp(prop),
                                                 _dp(post) if post is not None else None, C.byref(K)))
        return (prop, post) if want_posterior else prop

    def set_pattern_freq_from_posterior(self, src, k):
        """ptn_freq := column k of src's posteriors (device to device; the M-step trees of the EM algorithms)"""
        self._check(self.lib.b200lk_set_pattern_freq_from_posterior(self._h, src._h, C.c_int32(int(k))))

    def set_robust_phylo(self, keep=1.0, median=False):
        """Params::robust_phy_keep / robust_median: branch_logl returns the log-likelihood of the best sites / the median site"""
        self.lib.b200lk_set_robust_phylo.argtypes = [C.c_void_p, C.c_double, C.c_int32]
        self._check(self.lib.b200lk_set_robust_phylo(self._h, float(keep), int(bool(median))))

    def set_extra_children(self, children):
        """third, fourth ... children of the multifurcating ops (op['pad'] = their number) of the NEXT call, in op order"""
        ch = np.ascontiguousarray(children)
        self._check(self.lib.b200lk_set_extra_children(self._h, ch.ctypes.data_as(C.c_void_p), C.c_int32(len(ch))))

    def set_boot_samples(self, boot_samples):
        """UFBoot: resampled pattern frequencies [nboot][nptn] (float32), kept resident on the device."""
        b = np.ascontiguousarray(boot_samples, dtype=np.float32)
        assert b.ndim == 2 and b.shape[1] == self.nptn
        self._nboot = b.shape[0]
        self._check(self.lib.b200lk_set_boot_samples(self._h, b.ctypes.data_as(C.c_void_p), C.c_int32(b.shape[0])))

    def rell(self):
        """RELL log-likelihood of every bootstrap replicate for the tree of the last branch_logl / logl_from_buffer."""
        out = np.zeros(self._nboot, dtype=np.float32)
        self._check(self.lib.b200lk_rell(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def boot_reset(self):
        self._check(self.lib.b200lk_boot_reset(self._h))

    def rell_update(self, cur_logl, epsilon=0.5, uniform=None):
        """UFBoot bookkeeping for the last evaluation's _pattern_lh: (updated flags [nboot] uint8, number of updated replicates)"""
        u = None if uniform is None else np.ascontiguousarray(uniform, dtype=np.float64)
        upd = np.zeros(self._nboot, dtype=np.uint8)
        n = C.c_int32()
        self._check(self.lib.b200lk_rell_update(self._h, C.c_double(cur_logl), C.c_double(epsilon), _dp(u) if u is not None else None, _dp(upd), C.byref(n)))
        return upd, n.value

    def get_boot_state(self):
        bl, bc, bo = np.zeros(self._nboot), np.zeros(self._nboot, dtype=np.int32), np.zeros(self._nboot)
        self._check(self.lib.b200lk_get_boot_state(self._h, _dp(bl), _dp(bc), _dp(bo)))
        return bl, bc, bo

    def get_pattern_scaling(self):
        return self._get("b200lk_get_pattern_scaling", self.nptn)

    def get_pattern_lh_cat(self):
        return self._get("b200lk_get_pattern_lh_cat", self.nptn * self.Cmix).reshape(self.nptn, self.Cmix)

    def get_theta(self):
        return self._get("b200lk_get_theta", self.nptn * self.B).reshape(self.nptn, self.Cmix, self.n)

    def get_buffer_scale_all(self):
        return self._get("b200lk_get_buffer_scale_all", self.nptn)

    def get_scale_num(self, slot):
        S = self.Cmix if self.safe else 1
        return self._get("b200lk_get_scale_num", self.nptn * S, np.uint16, C.c_int32(slot)).reshape(self.nptn, S)

    def get_partial(self, slot):
        return self._get("b200lk_get_partial", self.nptn * self.B, np.float64, C.c_int32(slot)).reshape(self.nptn, self.Cmix, self.n)

    def get_tip_partial_lh(self):
        return self._get("b200lk_get_tip_partial_lh", (self.model.state_unknown + 1) * self.nmix * self.n)


def compress_patterns(states, device=0):
    """Site-pattern compression on the device (b200lk_compress_patterns): distinct columns of states [ntaxa][nsite] in
    first-occurrence order -> (patterns [ntaxa][nptn], ptn_freq [nptn], site_pattern [nsite])."""
    lib = load_library()
    st = np.ascontiguousarray(states, dtype=np.uint8)
    ntaxa, nsite = st.shape
    pat = np.zeros(ntaxa * nsite, dtype=np.uint8)
    freq = np.zeros(nsite)
    sp = np.zeros(nsite, dtype=np.int32)
    nptn = C.c_int64()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.b200lk_compress_patterns(C.c_int32(device), p(st), C.c_int32(ntaxa), C.c_int64(nsite), p(pat), p(freq), p(sp), C.byref(nptn))
    if rc != 0:
        raise EngineError(f"b200lk_compress_patterns failed ({rc}); there is no CPU fallback")
    n = nptn.value
    return pat[: ntaxa * n].reshape(ntaxa, n).copy(), freq[:n].copy(), sp


def to_simd_layout(partial_v1: np.ndarray, V: int) -> np.ndarray:
    """[nptn][B] (V=1) -> the reference's SIMD-interleaved layout for vector width V
    (partial_lh[(p/V)*B*V + k*V + p%V], tree/phylokernelnew.h:1914-1916); nptn is padded with zeros."""
    nptn = partial_v1.shape[0]
    flat = partial_v1.reshape(nptn, -1)
    Bk = flat.shape[1]
    pad = (-nptn) % V
    if pad:
        flat = np.concatenate([flat, np.zeros((pad, Bk), flat.dtype)])
    return flat.reshape(-1, V, Bk).transpose(0, 2, 1).reshape(-1)
