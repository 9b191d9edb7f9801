"""Per-partition model constants in ExaML's conventions, generated on the host with numpy.

The engine consumes exactly the arrays ExaML keeps in `pInfo` (axml.h:533-629):
EIGN[S], EV[S*S], EI[S*S], tipVector[codes*S], gammaRates[4].  In a drop-in build they are
produced by the reference's own models.c and merely passed through; this module regenerates
them for synthetic benchmarks and tests where the reference is absent.

Mirrors (behaviour, not code):
  initGeneric     models.c:3234-3457   symmetrised rate matrix -> eigen-system, EIGN/EV/EI, tipVector
  makeGammaCats   models.c:3795-3849   mean (or median) rates of 4 equiprobable Gamma(alpha) categories
  bitVectorAA     globalVariables.h:95-98
An eigen-basis is not unique (signs; any rotation inside a degenerate eigen-space), so the
arrays agree with the reference's only up to that freedom; every likelihood computed from them
is identical (tests/test_model.py).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

MAX_TIP_EV = 0.999999999        # axml.h:88

DNA_CODES = 16
AA_CODES = 23
BITVECTOR_AA = [1 << i for i in range(20)] + [12, 96, (1 << 20) - 1]


@dataclass
class ModelConstants:
    states: int
    EIGN: np.ndarray        # [S]
    EV: np.ndarray          # [S*S]  row-major EV[j*S+m]
    EI: np.ndarray          # [S*S]  row-major EI[j*S+k]
    tipVector: np.ndarray   # [codes*S]
    gammaRates: np.ndarray  # [4]
    frequencies: np.ndarray

    def copy(self) -> "ModelConstants":
        return ModelConstants(self.states, self.EIGN.copy(), self.EV.copy(), self.EI.copy(),
                              self.tipVector.copy(), self.gammaRates.copy(), self.frequencies.copy())


def gamma_rates(alpha: float, use_median: bool = False, K: int = 4) -> np.ndarray:
    """Discrete Gamma(alpha, alpha) category rates (models.c:3795-3849)."""
    from scipy.special import gammainc
    from scipy.stats import gamma as gamma_dist
    if use_median:
        mid = (2.0 * np.arange(K) + 1.0) / (2.0 * K)
        r = gamma_dist.ppf(mid, alpha, scale=1.0 / alpha)
        return r * K / r.sum()
    cuts = gamma_dist.ppf((np.arange(1, K)) / K, alpha, scale=1.0 / alpha)
    probs = gammainc(alpha + 1.0, cuts * alpha)
    r = np.empty(K)
    r[0] = probs[0] * K
    r[K - 1] = (1.0 - probs[K - 2]) * K
    for i in range(1, K - 1):
        r[i] = (probs[i] - probs[i - 1]) * K
    return r


def tip_bit_vectors(states: int):
    if states == 4:
        return list(range(DNA_CODES))
    if states == 20:
        return list(BITVECTOR_AA)
    raise ValueError("states must be 4 or 20")


def make_constants(rates: np.ndarray, freqs: np.ndarray, alpha: float = 1.0,
                   use_median: bool = False) -> ModelConstants:
    """rates: upper triangle of the symmetric exchangeability matrix, row by row
    (S*(S-1)/2 values: for DNA AC AG AT CG CT GT); freqs: stationary frequencies."""
    f = np.asarray(freqs, dtype=np.float64)
    S = f.size
    r = np.zeros((S, S))
    iu = np.triu_indices(S, 1)
    r[iu] = np.asarray(rates, dtype=np.float64)
    r = r + r.T
    fracchange = float(f @ r @ f)
    a = r * np.sqrt(np.outer(f, f))
    np.fill_diagonal(a, 0.0)
    np.fill_diagonal(a, -(r * f[None, :]).sum(axis=1))
    w, U = np.linalg.eigh(a)                 # ascending; the zero eigenvalue is the largest
    order = np.argsort(-w)                   # zero first, as the reference arranges it
    w, U = w[order], U[:, order]
    EIGV = U * np.sqrt(f)[:, None]           # EIGV[state, eigen]
    EIGV[:, 0] /= EIGV[:, 0].sum()           # first column becomes the frequencies
    EIGN = -w / fracchange
    EIGN[0] = 0.0
    EV = EIGV.copy()
    EI = EV / EIGV[:, [0]]
    EI[:, 0] = 1.0
    bits = tip_bit_vectors(S)
    tip = np.zeros((len(bits), S))
    for code, value in enumerate(bits):
        for j in range(S):
            if (value >> j) & 1:
                tip[code] += EIGV[j]
    tip = np.minimum(tip, MAX_TIP_EV)
    return ModelConstants(S, EIGN, EV.reshape(-1).copy(), EI.reshape(-1).copy(), tip.reshape(-1).copy(),
                          gamma_rates(alpha, use_median), f.copy())


def gtr_dna(rates=(1.0, 1.0, 1.0, 1.0, 1.0, 1.0), freqs=(0.25, 0.25, 0.25, 0.25), alpha=1.0) -> ModelConstants:
    return make_constants(np.array(rates), np.array(freqs), alpha)


def poisson_aa(alpha: float = 1.0, freqs=None, seed: int | None = None) -> ModelConstants:
    """20-state reversible model.  With seed=None: equal exchangeabilities (Poisson); with a seed:
    random positive exchangeabilities and frequencies (a generic GTR-like AA matrix, so that no
    eigen-value is degenerate -- the shape the engine sees for LG/WAG/...)."""
    S = 20
    if seed is None:
        ex = np.ones(S * (S - 1) // 2)
        f = np.full(S, 1.0 / S) if freqs is None else np.asarray(freqs, float)
    else:
        rng = np.random.default_rng(seed)
        ex = rng.gamma(2.0, 0.5, S * (S - 1) // 2) + 0.05
        f = rng.dirichlet(np.full(S, 20.0)) if freqs is None else np.asarray(freqs, float)
    return make_constants(ex, f, alpha)


def transition_matrix(mc: ModelConstants, t_z: float, rate: float = 1.0) -> np.ndarray:
    """P(z)[i, l] = sum_k EI[i,k] exp(EIGN[k] rate log z) EV[l,k] -- the state-space transition
    matrix implied by the constants (used by tests to cross-check against expm)."""
    S = mc.states
    EI = mc.EI.reshape(S, S)
    EV = mc.EV.reshape(S, S)
    d = np.exp(mc.EIGN * rate * np.log(t_z))
    return (EI * d[None, :]) @ EV.T
