"""ctypes wrapper around oracle/_build/liblkoracle.so — TEST INFRASTRUCTURE (see lk_oracle.c header).

Exposes the same method names as iqtree2_b200.engine.LikelihoodEngine so parity tests drive both with the
same code.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liblkoracle.so")


class _Config(C.Structure):
    _fields_ = [("nstates", C.c_int32), ("ncat", C.c_int32), ("nmix", C.c_int32), ("fused_mix_rate", C.c_int32),
                ("safe_numeric", C.c_int32), ("seq_type", C.c_int32), ("state_unknown", C.c_int32),
                ("ntaxa", C.c_int32), ("nptn", C.c_int64), ("nslots", C.c_int32), ("device", C.c_int32)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lk_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_last_error.restype = C.c_char_p
    return _lib


SEQ_TYPES = {"GENERIC": 0, "DNA": 1, "AA": 2}


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleEngine:
    """CPU oracle with the engine interface (set_* / update_partials / branch_logl / branch_derv / ...)."""

    def __init__(self, model, ntaxa, nptn, nslots=None, safe_numeric=False):
        self.model = model
        self.n, self.ncat, self.nmix = model.nstates, model.ncat, model.nmix
        self.Cmix = model.ncat_mix
        self.B = self.n * self.Cmix
        self.ntaxa, self.nptn = int(ntaxa), int(nptn)
        self.nslots = int(nslots if nslots is not None else max(ntaxa - 2, 1) + 2)
        self.safe = bool(safe_numeric)
        cfg = _Config(self.n, self.ncat, self.nmix, int(model.fused_mix_rate), int(self.safe), SEQ_TYPES[model.seq_type],
                      model.state_unknown, self.ntaxa, self.nptn, self.nslots, 0)
        self._h = C.c_void_p(lib().orc_create(C.byref(cfg)))
        if not self._h:
            raise MemoryError("orc_create failed")
        self.set_model(model)

    def close(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    __del__ = close

    # ---- inputs -------------------------------------------------------------------------------
    def set_model(self, model):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        ev, U, Ui, w = f(model.eval), f(model.evec), f(model.inv_evec), f(model.mix_weight)
        lib().orc_set_model(self._h, _dp(ev), _dp(U), _dp(Ui), _dp(w))
        r, p = f(model.cat_rate), f(model.cat_prop)
        lib().orc_set_rates(self._h, _dp(r), _dp(p))
        st = getattr(model, "site", None)
        if st:   # site-specific model (lk_oracle_site.h)
            b = [f(st["eval"]), f(st["evec"]), f(st["inv_evec"])]
            assert b[0].shape == (self.nptn, self.n)
            if lib().orc_set_site_model(self._h, _dp(b[0]), _dp(b[1]), _dp(b[2])) < 0:
                raise RuntimeError(lib().orc_last_error(self._h).decode())
        nr = getattr(model, "nonrev", None)
        if nr:   # non-reversible kernel (lk_oracle_nonrev.h)
            a = [f(nr["ceval"]), f(nr["cevec"]), f(nr["cinv_evec"]), f(nr["rate_matrix"]), f(model.state_freq)]
            rc = lib().orc_set_model_nonrev(self._h, _dp(a[0]), _dp(a[1]), _dp(a[2]), _dp(a[3]), _dp(a[4]), _dp(w), C.c_int(int(nr["rooted"])))
            if rc < 0:
                raise RuntimeError(lib().orc_last_error(self._h).decode())

    def set_tip_states(self, states):
        s = np.ascontiguousarray(states, dtype=np.uint8)
        assert s.shape == (self.ntaxa, self.nptn)
        lib().orc_set_tip_states(self._h, _dp(s))

    def set_pattern_freq(self, freq):
        f = np.ascontiguousarray(freq, dtype=np.float64)
        assert f.shape == (self.nptn,)
        lib().orc_set_pattern_freq(self._h, _dp(f))

    def set_pattern_invar(self, invar):
        f = np.ascontiguousarray(invar, dtype=np.float64)
        assert f.shape == (self.nptn,)
        lib().orc_set_pattern_invar(self._h, _dp(f))

    # ---- compute ------------------------------------------------------------------------------
    def set_ascertainment(self, asc_type, num_unobserved, nsite):
        """+ASC (1 = Lewis): the last num_unobserved patterns are the unobserved constant patterns, nsite = aln->getNSite()"""
        rc = lib().orc_set_ascertainment(self._h, C.c_int(asc_type), C.c_int64(num_unobserved), C.c_int64(nsite))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())

    def set_extra_children(self, children):
        """third, fourth ... children of the multifurcating ops (op['pad'] = their number) of the NEXT call, in op order"""
        ch = np.ascontiguousarray(children)
        lib().orc_set_extra_children(self._h, _dp(ch), C.c_int(len(ch)))

    def update_partials(self, ops):
        ops = np.ascontiguousarray(ops)
        rc = lib().orc_update_partials(self._h, _dp(ops), C.c_int(len(ops)))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())

    def branch_logl(self, ops, end_a, end_b, length, save_log_value=True):
        ops = np.ascontiguousarray(ops)
        out = C.c_double()
        assert end_a[0] == 1
        rc = lib().orc_branch_logl(self._h, _dp(ops), C.c_int(len(ops)), C.c_int(end_a[1]), C.c_int(end_b[0]),
                                   C.c_int(end_b[1]), C.c_double(length), C.c_int(int(save_log_value)), C.byref(out))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        self.last_status = rc
        return out.value

    def theta_reset(self):
        lib().orc_theta_reset(self._h)

    def set_robust_phylo(self, keep=1.0, median=False):
        """Params::robust_phy_keep / robust_median: branch_logl returns the log-likelihood of the best sites / the median site"""
        L = lib()
        L.orc_set_robust_phylo.argtypes = [C.c_void_p, C.c_double, C.c_int]
        if L.orc_set_robust_phylo(self._h, float(keep), int(bool(median))) < 0:
            raise RuntimeError(L.orc_last_error(self._h).decode())

    def set_reference_quirks(self, mask):
        lib().orc_set_reference_quirks(self._h, C.c_int(int(mask)))

    def branch_derv(self, ops, end_a, end_b, length):
        ops = np.ascontiguousarray(ops)
        df, ddf = C.c_double(), C.c_double()
        rc = lib().orc_branch_derv(self._h, _dp(ops), C.c_int(len(ops)), C.c_int(end_a[1]), C.c_int(end_b[0]),
                                   C.c_int(end_b[1]), C.c_double(length), C.byref(df), C.byref(ddf))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        self.last_status = rc
        return df.value, ddf.value

    # ---- mixed branch lengths (lk_oracle_mixlen.h) --------------------------------------------------
    def set_mixlen(self, table):
        """table [nbranches][ncat] of per-class branch lengths (None / empty: off); afterwards every `length` is a ROW of it"""
        t = np.ascontiguousarray(table if table is not None else np.zeros((0, 1)), dtype=np.float64)
        if lib().orc_set_mixlen(self._h, _dp(t), C.c_int(t.shape[0])) < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        self._nmixlen = t.shape[1] if t.shape[0] else 0

    def branch_derv_mixlen(self, ops, end_a, end_b, row):
        """computeLikelihoodDerv of a mixlen tree -> (gradient [K], hessian [K][K], logL)"""
        ops = np.ascontiguousarray(ops)
        K = self._nmixlen
        df, ddf = np.zeros(K + 1), np.zeros((K, K))
        rc = lib().orc_branch_derv_mixlen(self._h, _dp(ops), C.c_int(len(ops)), C.c_int(end_a[1]), C.c_int(end_b[0]),
                                          C.c_int(end_b[1]), C.c_double(row), _dp(df), _dp(ddf))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        return df[:K].copy(), ddf, float(df[K])

    def branch_derv_mixlen_cur(self, ops, end_a, end_b, row, cur):
        """computeLikelihoodDervMixlen for class `cur` -> (df, ddf)"""
        ops = np.ascontiguousarray(ops)
        df, ddf = C.c_double(), C.c_double()
        rc = lib().orc_branch_derv_mixlen_cur(self._h, _dp(ops), C.c_int(len(ops)), C.c_int(end_a[1]), C.c_int(end_b[0]),
                                              C.c_int(end_b[1]), C.c_double(row), C.c_int(cur), C.byref(df), C.byref(ddf))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        self.last_status = rc
        return df.value, ddf.value

    def logl_from_buffer(self, length):
        out = C.c_double()
        rc = lib().orc_logl_from_buffer(self._h, C.c_double(length), C.byref(out))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error(self._h).decode())
        self.last_status = rc
        return out.value

    # ---- read-backs ---------------------------------------------------------------------------
    def _get(self, fn, n, dtype=np.float64, *args):
        out = np.empty(n, dtype=dtype)
        getattr(lib(), fn)(self._h, *args, _dp(out))
        return out

    def get_pattern_lh(self):
        return self._get("orc_get_pattern_lh", self.nptn)

    def get_pattern_scaling(self):
        return self._get("orc_get_pattern_scaling", self.nptn)

    def get_pattern_lh_cat(self):
        return self._get("orc_get_pattern_lh_cat", self.nptn * self.Cmix).reshape(self.nptn, self.Cmix)

    def pattern_lh_cat_wsl(self, wsl):
        """computePatternLhCat(wsl) of the last branch_logl: [nptn][K] (2 = RATECAT, 3 = MIXTURE, 4 = MIXTURE_RATECAT)"""
        L = lib()
        K = L.orc_pattern_lh_cat_wsl(self._h, C.c_int(wsl), None)
        if K < 0:
            raise RuntimeError(L.orc_last_error(self._h).decode())
        out = np.zeros((self.nptn, K))
        L.orc_pattern_lh_cat_wsl(self._h, C.c_int(wsl), _dp(out))
        return out

    def em_posterior(self, wsl, add_invar):
        """E-step of RateFree / ModelMixture::optimizeWithEM: (new_prop[K], posterior[nptn][K])"""
        L = lib()
        K = L.orc_pattern_lh_cat_wsl(self._h, C.c_int(wsl), None)
        if K < 0:
            raise RuntimeError(L.orc_last_error(self._h).decode())
        prop, post = np.zeros(K), np.zeros((self.nptn, K))
        L.orc_em_posterior(self._h, C.c_int(wsl), C.c_int(int(bool(add_invar))), _dp(prop), _dp(post))
        return prop, post

    def get_theta(self):
        return self._get("orc_get_theta", self.nptn * self.B).reshape(self.nptn, self.Cmix, self.n)

    def get_buffer_scale_all(self):
        return self._get("orc_get_buffer_scale_all", self.nptn)

    def get_scale_num(self, slot):
        S = self.Cmix if self.safe else 1
        return self._get("orc_get_scale_num", self.nptn * S, np.uint16, C.c_int(slot)).reshape(self.nptn, S)

    def get_partial(self, slot):
        """[nptn][ncat_mix][nstates] (the reference's V=1 layout)."""
        return self._get("orc_get_partial", self.nptn * self.B, np.float64, C.c_int(slot)).reshape(self.nptn, self.Cmix, self.n)

    def get_tip_partial_lh(self):
        return self._get("orc_get_tip_partial_lh", (self.model.state_unknown + 1) * self.nmix * self.n)

    @staticmethod
    def num_threads():
        return int(lib().orc_num_threads())
