"""Synthetic PLINK data: numpy twin of csrc/synth.cuh (identical bytes) and response simulators.

Distributions follow the reference's `simulate_random_snparray` (src/simulate_utilities.jl:32-50) and
`simulate_random_response` (src/simulate_utilities.jl:207-242): maf_j = clip(0.5 U, 0.01, 0.5), genotype =
Bern(maf) + Bern(maf); k causal SNPs with N(0,1) effects (N(0, 0.3^2) for count traits); eta = X beta (+ Z gamma),
clamped to +-20 for non-Normal traits; y ~ d(linkinv(eta)).  Everything is keyed by explicit seeds.
"""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def mix64(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30); x *= _M1
        x ^= x >> np.uint64(27); x *= _M2
        x ^= x >> np.uint64(31)
    return x


def col_key(seed: int, j):
    with np.errstate(over="ignore"):
        return mix64(np.uint64(seed) ^ ((np.asarray(j, dtype=np.uint64) + np.uint64(1)) * _GOLD))


def maf_threshold(key):
    h = mix64(key ^ np.uint64(0xA5A5A5A5A5A5A5A5))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    maf = np.clip(0.5 * u, 0.01, 0.5)
    return (maf * 4294967296.0).astype(np.uint64), maf


def codes(seed: int, n: int, cols, missing_rate: float = 0.0) -> np.ndarray:
    """PLINK 2-bit codes [n, len(cols)] (uint8) of the given global columns."""
    cols = np.asarray(cols, dtype=np.uint64)
    key = col_key(seed, cols)                                   # [c]
    thr, _ = maf_threshold(key)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = mix64(key[None, :] ^ (i[:, None] * _GOLD))          # [n, c]
    a1 = (h & np.uint64(0xFFFFFFFF)) < thr[None, :]
    a2 = (h >> np.uint64(32)) < thr[None, :]
    g = a1.astype(np.uint8) + a2.astype(np.uint8)
    code = np.where(g > 0, g + 1, 0).astype(np.uint8)
    miss_thr = np.uint64(int(missing_rate * 4294967296.0))
    if miss_thr:
        h2 = mix64(h ^ np.uint64(0x5851F42D4C957F2D))
        code[(h2 & np.uint64(0xFFFFFFFF)) < miss_thr] = 1
    return code


def packed_columns(seed: int, n: int, cols, missing_rate: float = 0.0) -> np.ndarray:
    """Packed .bed columns [len(cols), ceil(n/4)] (uint8)."""
    c = codes(seed, n, cols, missing_rate)
    nb = (n + 3) // 4
    pad = np.zeros((nb * 4, c.shape[1]), dtype=np.uint8)
    pad[:n] = c
    pad = pad.T.reshape(c.shape[1], nb, 4)
    return (pad[:, :, 0] | (pad[:, :, 1] << 2) | (pad[:, :, 2] << 4) | (pad[:, :, 3] << 6)).astype(np.uint8)


def ternary_tiles(bed: np.ndarray, n: int) -> np.ndarray:
    """Host twin of the ternary copy of a genotype handle (geno.cu k_make_tern): the byte stream the FAST / PAIR sweeps read.
    bed: packed PLINK columns [p, ceil(n/4)].  Slabs of 640 samples; per slab, columns in groups of four; the 16 bytes at word
    position w of a group hold the words of its four columns, component i = column 4 q + (i ^ (w & 3)); a word packs samples
    20 w .. 20 w + 19 of the slab, a byte five dosages in base 3 (missing counts as 0)."""
    p = bed.shape[0]
    p4 = (p + 3) // 4 * 4
    slabs = -(-n // 640)
    codes4 = np.stack([(bed >> (2 * s)) & 3 for s in range(4)], axis=2).reshape(p, -1)[:, :n]      # [p, n] PLINK codes
    dos = np.array([0, 0, 1, 2], dtype=np.uint32)[codes4]                                          # 01 (missing) -> 0
    pad = np.zeros((p4, slabs * 640), dtype=np.uint32)
    pad[:p, :n] = dos
    d = pad.reshape(p4, slabs, 32, 4, 5)                                                           # column, slab, w, byte, digit
    byte = (d * np.array([1, 3, 9, 27, 81], dtype=np.uint32)).sum(axis=4).astype(np.uint8)         # [p4, slabs, 32, 4]
    out = np.zeros((slabs, p4 // 4, 32, 4, 4), dtype=np.uint8)                                     # slab, quad, w, component, byte
    for w in range(32):
        for i in range(4):
            cols = np.arange(p4 // 4) * 4 + (i ^ (w & 3))
            out[:, :, w, i, :] = byte[cols, :, w, :].transpose(1, 0, 2)
    return out.reshape(-1)


def standardized_columns(seed: int, n: int, cols, missing_rate: float = 0.0) -> np.ndarray:
    """x[:, cols] with SnpLinAlg(center, scale, impute) semantics, float64 [n, len(cols)]."""
    c = codes(seed, n, cols, missing_rate)
    dos = np.array([0.0, np.nan, 1.0, 2.0])[c]
    miss = np.isnan(dos)
    nobs = n - miss.sum(axis=0)
    mu = np.nansum(dos, axis=0) / nobs
    s = np.sqrt(mu * (1 - mu / 2))
    sinv = np.where(s > 0, 1.0 / np.where(s > 0, s, 1.0), 1.0)
    x = np.where(miss, mu[None, :], dos)
    return (x - mu[None, :]) * sinv[None, :]


def simulate_response(seed: int, n: int, p: int, k: int, d: str = "Normal", n_cov: int = 0, geno_seed: int = None,
                      missing_rate: float = 0.0, nb_r: float = 10.0):
    """Returns (y, z, true_idx, true_beta, true_c).  z is n x (1 + n_cov): intercept + standardised N(0,1) covariates."""
    rng = np.random.default_rng(seed)
    geno_seed = seed if geno_seed is None else geno_seed
    idx = np.sort(rng.permutation(p)[:k])
    scale = 1.0 if d in ("Normal", "Bernoulli") else 0.3
    beta = rng.normal(0.0, scale, size=k)
    z = np.ones((n, 1 + n_cov))
    if n_cov:
        zc = rng.normal(size=(n, n_cov))
        zc = (zc - zc.mean(axis=0)) / zc.std(axis=0, ddof=1)
        z[:, 1:] = zc
    c = np.concatenate([[1.0], rng.normal(0.0, 0.5, size=n_cov)]) if d == "Normal" else \
        np.concatenate([[0.0 if d == "Bernoulli" else 1.0], rng.normal(0.0, 0.1, size=n_cov)])
    eta = z @ c
    # causal columns in blocks to bound memory
    for s in range(0, k, 16):
        xs = standardized_columns(geno_seed, n, idx[s:s + 16], missing_rate)
        eta += xs @ beta[s:s + 16]
    if d == "Normal":
        y = eta + rng.normal(size=n)
    else:
        eta = np.clip(eta, -20, 20)
        if d == "Bernoulli":
            y = (rng.random(n) < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64)
        elif d == "Poisson":
            y = rng.poisson(np.exp(eta)).astype(np.float64)
        elif d == "NegativeBinomial":
            mu = np.exp(eta)
            y = rng.negative_binomial(nb_r, nb_r / (mu + nb_r)).astype(np.float64)
        else:
            raise ValueError(d)
    return y, z, idx, beta, c


def folds_for(seed: int, n: int, q: int) -> np.ndarray:
    """Explicit CV folds 1..q: folds_i = 1 + (hash(seed, i) mod q)."""
    with np.errstate(over="ignore"):
        h = mix64(np.uint64(seed) ^ (np.arange(n, dtype=np.uint64) * _GOLD))
    return (1 + (h % np.uint64(q))).astype(np.int64)
