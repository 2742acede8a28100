"""Kernel INPUTS: substitution-model eigensystems and rate categories (host side, numpy).

The likelihood engine never derives a model; it consumes (eval, evec, inv_evec, state_freq, mixture
weights, category rates/proportions) exactly as the reference kernels do ("Layout crib").
This module produces those arrays for tests and for bench.py's synthetic workloads, with the
reference's normalisation so that partials live in the same eigen-coordinates:

    evec = diag(1/sqrt(pi)) . V        inv_evec = V^T . diag(sqrt(pi))
    (reference: model/modelmarkov.cpp:1535-1562 Eigen3 path, utils/eigendecomposition.cpp:218-240)

Eigenvector sign/order is solver dependent, which is why the C ABI takes the eigensystem as an input
(SURVEY §8c "Third-party arithmetic on the path").
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

DNA_STATE_UNKNOWN = 18      # alignment/alignment.cpp:738
PROTEIN_STATE_UNKNOWN = 23  # alignment/alignment.cpp:739
PROTEIN_SYMBOLS = "ARNDCQEGHILKMFPSTWYV"  # alignment/alignment.cpp:33


@dataclass
class ModelInputs:
    """Everything computePartialInfo / computeLikelihood* read from model + site_rate."""
    nstates: int
    nmix: int
    eval: np.ndarray        # [nmix*n]
    evec: np.ndarray        # [nmix*n*n]  row-major  evec[m][x][i]
    inv_evec: np.ndarray    # [nmix*n*n]  row-major  inv_evec[m][i][x]
    state_freq: np.ndarray  # [nmix*n]
    mix_weight: np.ndarray  # [nmix]
    cat_rate: np.ndarray    # [ncat]
    cat_prop: np.ndarray    # [ncat]
    pinvar: float = 0.0
    state_unknown: int = 0  # STATE_UNKNOWN of the alignment type
    seq_type: str = "DNA"
    fused_mix_rate: bool = False
    Q: list = field(default_factory=list)  # normalised rate matrices per mixture class (for the brute-force checker)
    # non-reversible models (tree/phylokernelnonrev.h): None, or a dict with the complex eigensystem the reference's default
    # technique uses (ModelMarkov::computeTransMatrixNonrev, model/modelmarkov.cpp:477-486) — "ceval" [nmix][n][2] (re, im),
    # "cevec" / "cinv_evec" [nmix][n][n][2] row-major, "rate_matrix" [nmix][n][n] (for the derivatives Q.P, Q.Q.P) and "rooted"
    nonrev: dict = None
    # site-specific models (ModelSet: `-fs` site-frequency files / PMSF; the SITE_MODEL branch of tree/phylokernelnew.h): None, or a
    # dict with one eigensystem per PATTERN — "eval" [nptn][n], "evec" [nptn][n][n], "inv_evec" [nptn][n][n], "Q" (list, for checkers)
    site: dict = None
    # mixed branch lengths (heterotachy; PhyloTreeMixlen): None, or [ncat] factors — the test and bench trees give every branch the
    # per-class lengths  length * factor[c]  (the engines themselves take an arbitrary [nbranches][ncat] table)
    mixlen_factor: np.ndarray = None

    @property
    def ncat(self) -> int:
        return len(self.cat_rate)

    @property
    def ncat_mix(self) -> int:
        return self.ncat if self.fused_mix_rate else self.ncat * self.nmix


def discrete_gamma_rates(alpha: float, ncat: int, median: bool = False) -> np.ndarray:
    """Mean-of-category discrete Gamma (Yang 1994), as model/rategamma.cpp:149-163 computes it."""
    from scipy.stats import gamma as gdist
    from scipy.special import gammainc
    if ncat == 1:
        return np.ones(1)
    if median:
        probs = (2.0 * np.arange(ncat) + 1.0) / (2.0 * ncat)
        r = gdist.ppf(probs, alpha, scale=1.0 / alpha)
        return r * ncat / r.sum()
    cuts = gdist.ppf(np.arange(1, ncat) / ncat, alpha, scale=1.0 / alpha)
    fk = gammainc(alpha + 1.0, cuts * alpha)
    fk = np.concatenate([[0.0], fk, [1.0]])
    return (fk[1:] - fk[:-1]) * ncat


def reversible_eigensystem(exch: np.ndarray, pi: np.ndarray):
    """Eigendecomposition of Q = S.diag(pi) normalised to one expected substitution per unit time.

    Returns (eval[n], evec[n][n], inv_evec[n][n], Q[n][n]) with the reference's normalisation.
    """
    n = len(pi)
    pi = np.asarray(pi, dtype=np.float64)
    pi = pi / pi.sum()
    S = np.array(exch, dtype=np.float64)
    S = 0.5 * (S + S.T)
    np.fill_diagonal(S, 0.0)
    Q = S * pi[None, :]
    np.fill_diagonal(Q, -Q.sum(axis=1))
    scale = -(pi * np.diag(Q)).sum()
    Q = Q / scale
    sq = np.sqrt(pi)
    A = (sq[:, None] * Q) / sq[None, :]
    A = 0.5 * (A + A.T)
    w, V = np.linalg.eigh(A)
    evec = V / sq[:, None]
    inv_evec = V.T * sq[None, :]
    return w, evec, inv_evec, Q


def _pack(nstates, systems, weights, cat_rate, cat_prop, pinvar, state_unknown, seq_type):
    nmix = len(systems)
    return ModelInputs(
        nstates=nstates, nmix=nmix,
        eval=np.concatenate([s[0] for s in systems]).astype(np.float64),
        evec=np.concatenate([s[1].ravel() for s in systems]).astype(np.float64),
        inv_evec=np.concatenate([s[2].ravel() for s in systems]).astype(np.float64),
        state_freq=np.concatenate([s[4] for s in systems]).astype(np.float64),
        mix_weight=np.asarray(weights, dtype=np.float64),
        cat_rate=np.asarray(cat_rate, dtype=np.float64),
        cat_prop=np.asarray(cat_prop, dtype=np.float64),
        pinvar=float(pinvar), state_unknown=state_unknown, seq_type=seq_type,
        Q=[s[3] for s in systems])


def gtr_exchangeabilities(rates6):
    """GTR rate order AC, AG, AT, CG, CT, GT (model/modeldna.cpp parameter order)."""
    a, b, c, d, e, f = rates6
    return np.array([[0, a, b, c], [a, 0, d, e], [b, d, 0, f], [c, e, f, 0]], dtype=np.float64)


def make_gtr_gamma(rates6=(1.0, 2.0, 1.0, 1.0, 2.0, 1.0), pi=(0.3, 0.2, 0.2, 0.3), alpha=0.5, ncat=4,
                   pinvar=0.0) -> ModelInputs:
    pi = np.asarray(pi, dtype=np.float64)
    pi = pi / pi.sum()
    w, U, Ui, Q = reversible_eigensystem(gtr_exchangeabilities(rates6), pi)
    rates = discrete_gamma_rates(alpha, ncat)
    props = np.full(ncat, (1.0 - pinvar) / ncat)
    return _pack(4, [(w, U, Ui, Q, pi)], [1.0], rates, props, pinvar, DNA_STATE_UNKNOWN, "DNA")


def load_lg():
    """LG exchangeabilities/frequencies (PAML lower-triangle file extracted by scripts/extract_model_data.py)."""
    rows = [ln.split() for ln in open(os.path.join(_DATA, "LG.paml")) if ln.strip()]
    S = np.zeros((20, 20))
    for i in range(1, 20):
        for j in range(i):
            S[i, j] = S[j, i] = float(rows[i - 1][j])
    pi = np.array([float(x) for x in rows[19]])
    return S, pi / pi.sum()


def make_lg_gamma(alpha=1.0, ncat=4, pinvar=0.0) -> ModelInputs:
    S, pi = load_lg()
    w, U, Ui, Q = reversible_eigensystem(S, pi)
    rates = discrete_gamma_rates(alpha, ncat)
    props = np.full(ncat, (1.0 - pinvar) / ncat)
    return _pack(20, [(w, U, Ui, Q, pi)], [1.0], rates, props, pinvar, PROTEIN_STATE_UNKNOWN, "AA")


def make_lg_c60_gamma(alpha=1.0, ncat=4, nclasses=60) -> ModelInputs:
    """LG+C60+G4: 60 frequency classes sharing LG exchangeabilities (model/modelmixture.cpp:915-977)."""
    S, _ = load_lg()
    spec = json.load(open(os.path.join(_DATA, "C60.json")))["classes"][:nclasses]
    systems, weights = [], []
    for cl in spec:
        pi = np.array(cl["freq"], dtype=np.float64)
        pi = pi / pi.sum()
        w, U, Ui, Q = reversible_eigensystem(S, pi)
        systems.append((w, U, Ui, Q, pi))
        weights.append(cl["weight"])
    weights = np.array(weights)
    weights = weights / weights.sum()
    rates = discrete_gamma_rates(alpha, ncat)
    props = np.full(ncat, 1.0 / ncat)
    return _pack(20, systems, weights, rates, props, 0.0, PROTEIN_STATE_UNKNOWN, "AA")


def make_random_reversible(nstates, ncat=4, nmix=1, alpha=0.7, seed=0, pinvar=0.0, state_unknown=None) -> ModelInputs:
    """Random reversible model with arbitrary state count (binary=2, codon=61, ...) for generic-n parity tests."""
    rng = np.random.default_rng(seed)
    systems, weights = [], []
    for _ in range(nmix):
        S = rng.gamma(1.0, 1.0, size=(nstates, nstates)) + 0.05
        pi = rng.dirichlet(np.full(nstates, 5.0))
        w, U, Ui, Q = reversible_eigensystem(S, pi)
        systems.append((w, U, Ui, Q, pi))
        weights.append(rng.uniform(0.5, 1.5))
    weights = np.array(weights) / np.sum(weights)
    rates = discrete_gamma_rates(alpha, ncat)
    props = np.full(ncat, (1.0 - pinvar) / ncat)
    su = nstates if state_unknown is None else state_unknown
    seq_type = "DNA" if (nstates == 4 and su == DNA_STATE_UNKNOWN) else ("AA" if (nstates == 20 and su == PROTEIN_STATE_UNKNOWN) else "GENERIC")
    return _pack(nstates, systems, weights, rates, props, pinvar, su, seq_type)


def make_site_specific(nstates, nptn, ncat=4, alpha=0.7, seed=0, pinvar=0.0, state_unknown=None, exch=None) -> ModelInputs:
    """PMSF-like site-specific model: ONE set of exchangeabilities, a frequency vector per pattern (Dirichlet), hence one
    eigensystem per pattern with the reference's normalisation.  The class-level arrays (eval / evec / inv_evec / state_freq) hold
    the mean-frequency model and are not used by the site-specific kernels."""
    rng = np.random.default_rng(seed)
    n = nstates
    S = (rng.gamma(1.0, 1.0, size=(n, n)) + 0.05) if exch is None else exch
    S = (S + S.T) / 2
    pis = rng.dirichlet(np.full(n, 3.0), size=nptn) * 0.9 + 0.1 / n
    ev, U, Ui, Qs = [], [], [], []
    for p in range(nptn):
        w, u, ui, Q = reversible_eigensystem(S, pis[p])
        ev.append(w); U.append(u); Ui.append(ui); Qs.append(Q)
    mean_pi = pis.mean(axis=0)
    w, u, ui, Q = reversible_eigensystem(S, mean_pi)
    rates = discrete_gamma_rates(alpha, ncat) if ncat > 1 else np.ones(1)
    props = np.full(ncat, (1.0 - pinvar) / ncat)
    su = nstates if state_unknown is None else state_unknown
    seq_type = "DNA" if (nstates == 4 and su == DNA_STATE_UNKNOWN) else ("AA" if (nstates == 20 and su == PROTEIN_STATE_UNKNOWN) else "GENERIC")
    m = _pack(n, [(w, u, ui, Q, mean_pi)], [1.0], rates, props, pinvar, su, seq_type)
    m.site = dict(eval=np.ascontiguousarray(np.stack(ev)), evec=np.ascontiguousarray(np.stack(U)), inv_evec=np.ascontiguousarray(np.stack(Ui)),
                  Q=Qs, pi=pis)
    return m


def site_slice(model: ModelInputs, lo: int, hi: int) -> ModelInputs:
    """the model of the pattern range [lo, hi) of a site-specific model (a pattern shard, a CPU sample); any other model as it is"""
    if getattr(model, "site", None) is None:
        return model
    import dataclasses
    st = model.site
    cut = {k: (v[lo:hi] if k in ("eval", "evec", "inv_evec", "pi") else v) for k, v in st.items() if k != "Q"}
    cut["Q"] = st["Q"][lo:hi] if "Q" in st else []
    return dataclasses.replace(model, site={k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in cut.items()})


def make_site_specific_fast(nstates, nptn, ncat=4, alpha=0.7, seed=0, nprofiles=256) -> ModelInputs:
    """make_site_specific for large pattern counts: `nprofiles` distinct frequency profiles, assigned to the patterns at random
    (the arrays are still one eigensystem per pattern, as ModelSet holds them)"""
    base = make_site_specific(nstates, nprofiles, ncat=ncat, alpha=alpha, seed=seed,
                              state_unknown=DNA_STATE_UNKNOWN if nstates == 4 else (PROTEIN_STATE_UNKNOWN if nstates == 20 else None))
    idx = np.random.default_rng(seed + 1).integers(0, nprofiles, size=nptn)
    st = base.site
    base.site = dict(eval=np.ascontiguousarray(st["eval"][idx]), evec=np.ascontiguousarray(st["evec"][idx]),
                     inv_evec=np.ascontiguousarray(st["inv_evec"][idx]), pi=st["pi"][idx], Q=[])
    return base


def make_nonreversible(nstates=4, ncat=4, nmix=1, alpha=0.7, seed=0, pinvar=0.0, state_unknown=None, rooted=True) -> ModelInputs:
    """Random NON-reversible model (an UNREST-like general Markov rate matrix per class): Q with arbitrary positive off-diagonals,
    normalised to one expected substitution per unit time at its stationary distribution; complex eigensystem from numpy.  The
    reversible-kernel fields (eval / evec / inv_evec) are filled with the real parts and are not used by the non-reversible path."""
    rng = np.random.default_rng(seed)
    n = nstates
    ceval, cevec, cinv, qs, freqs, weights = [], [], [], [], [], []
    for _ in range(nmix):
        Q = rng.gamma(1.0, 1.0, size=(n, n)) + 0.05
        np.fill_diagonal(Q, 0.0)
        np.fill_diagonal(Q, -Q.sum(axis=1))
        # stationary distribution: left null vector
        w, vl = np.linalg.eig(Q.T)
        pi = np.real(vl[:, np.argmin(np.abs(w))])
        pi = pi / pi.sum()
        Q = Q / (-(pi * np.diag(Q)).sum())
        lam, V = np.linalg.eig(Q)
        Vi = np.linalg.inv(V)
        ceval.append(np.stack([lam.real, lam.imag], axis=-1))
        cevec.append(np.stack([V.real, V.imag], axis=-1))
        cinv.append(np.stack([Vi.real, Vi.imag], axis=-1))
        qs.append(Q)
        freqs.append(pi)
        weights.append(rng.uniform(0.5, 1.5))
    weights = np.array(weights) / np.sum(weights)
    rates = discrete_gamma_rates(alpha, ncat) if ncat > 1 else np.ones(1)
    props = np.full(ncat, (1.0 - pinvar) / ncat)
    su = nstates if state_unknown is None else state_unknown
    seq_type = "DNA" if (nstates == 4 and su == DNA_STATE_UNKNOWN) else ("AA" if (nstates == 20 and su == PROTEIN_STATE_UNKNOWN) else "GENERIC")
    f64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return ModelInputs(
        nstates=n, nmix=nmix, eval=f64(np.concatenate([c[:, 0] for c in ceval])), evec=f64(np.concatenate([c[..., 0].ravel() for c in cevec])),
        inv_evec=f64(np.concatenate([c[..., 0].ravel() for c in cinv])), state_freq=f64(np.concatenate(freqs)), mix_weight=f64(weights),
        cat_rate=f64(rates), cat_prop=f64(props), pinvar=float(pinvar), state_unknown=su, seq_type=seq_type, Q=qs,
        nonrev=dict(ceval=f64(np.stack(ceval)), cevec=f64(np.stack(cevec)), cinv_evec=f64(np.stack(cinv)), rate_matrix=f64(np.stack(qs)),
                    rooted=bool(rooted)))


def state_indicator(state: int, nstates: int, state_unknown: int, seq_type: str) -> np.ndarray:
    """0/1 tip likelihood of an observed (possibly ambiguous) state.

    model/modelsubst.cpp:197-207, model/modeldna.cpp:462-475, model/modelprotein.cpp:1149-1166.
    """
    v = np.zeros(nstates)
    if state < nstates:
        v[state] = 1.0
    elif seq_type == "DNA" and state < 18:
        c = state - nstates + 1
        for i in range(nstates):
            if c & (1 << i):
                v[i] = 1.0
    elif seq_type == "AA" and state < nstates + 3:
        amb = [(2, 3), (5, 6), (9, 10)][state - nstates]
        v[amb[0]] = v[amb[1]] = 1.0
    else:
        v[:] = 1.0
    return v


def tip_partial_lh(model: ModelInputs) -> np.ndarray:
    """tip_partial_lh[state][m][i] = sum_j inv_evec_m[i][j] * ind_state[j]  (phylotreesse.cpp:367-375)."""
    n, nmix, su = model.nstates, model.nmix, model.state_unknown
    out = np.zeros((su + 1, nmix, n))
    inv = model.inv_evec.reshape(nmix, n, n)
    for s in range(su + 1):
        ind = state_indicator(s, n, su, model.seq_type)
        for m in range(nmix):
            # same accumulation order as ModelSubst::multiplyWithInvEigenvector (j ascending, plain adds)
            acc = np.zeros(n)
            for j in range(n):
                acc = acc + inv[m][:, j] * ind[j]
            out[s, m] = acc
    return out.ravel()
