"""Oracle (test infrastructure): PLINK .bed genotype operator, numpy restatement.

Restates what `SnpLinAlg{Float64}(s; model=ADDITIVE_MODEL, center=true, scale=true,
impute=true)` means to MendelIHT (constructed at reference `src/wrapper.jl:68-69,318-319`;
fields read at `src/fit.jl:97-101`).  SnpArrays.jl itself is an un-vendored dependency
(`Project.toml:20,35`, compat 0.3.15); its semantics are cross-checked against the
reference's in-repo dense restatement `standardize_genotypes!` (`src/wrapper.jl:406-423`)
which `test/wrapper_test.jl:186-202` asserts is equivalent.

.bed layout: 3 magic bytes 6c 1b 01, then SNP-major columns of ceil(n/4) bytes; sample i
of a column lives in byte i>>2, bits 2*(i&3)..+1; codes 00->0, 01->missing, 10->1, 11->2
(consistent with reference `src/simulate_utilities.jl:85-101`, `src/utilities.jl:871-893`).
"""
from __future__ import annotations

import numpy as np

BED_MAGIC = bytes([0x6C, 0x1B, 0x01])

# code -> dosage (NaN = missing)
_DOSAGE = np.array([0.0, np.nan, 1.0, 2.0])


def read_bed(path: str, n: int) -> np.ndarray:
    """Return the packed matrix as uint8 [p, ceil(n/4)] (one row per SNP column)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw[:3].tobytes() != BED_MAGIC:
        raise ValueError("not a SNP-major PLINK .bed file")
    stride = (n + 3) // 4
    body = raw[3:]
    if body.size % stride:
        raise ValueError("bed size is not a multiple of ceil(n/4)")
    return body.reshape(-1, stride)


def write_bed(path: str, bed: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(BED_MAGIC)
        f.write(np.ascontiguousarray(bed, dtype=np.uint8).tobytes())


def unpack_codes(bed: np.ndarray, n: int) -> np.ndarray:
    """2-bit codes as uint8 [n, p] (sample-major rows, like the Julia SnpArray)."""
    p, stride = bed.shape
    codes = np.empty((p, stride * 4), dtype=np.uint8)
    for k in range(4):
        codes[:, k::4] = (bed >> (2 * k)) & 3
    return np.ascontiguousarray(codes[:, :n].T)


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """Inverse of unpack_codes: codes [n, p] -> bed [p, ceil(n/4)]."""
    n, p = codes.shape
    stride = (n + 3) // 4
    pad = np.zeros((stride * 4, p), dtype=np.uint8)
    pad[:n] = codes
    pad = pad.T.reshape(p, stride, 4)
    return (pad[:, :, 0] | (pad[:, :, 1] << 2) | (pad[:, :, 2] << 4) | (pad[:, :, 3] << 6)).astype(np.uint8)


def dosage_to_codes(g: np.ndarray) -> np.ndarray:
    """{0,1,2,-1(missing)} -> PLINK codes (reference `_make_snparray`, simulate_utilities.jl:85-101)."""
    lut = np.array([0, 2, 3, 1], dtype=np.uint8)  # index 3 used for missing (-1 -> 3)
    return lut[np.where(g < 0, 3, g).astype(np.int64)]


def dosages(bed: np.ndarray, n: int) -> np.ndarray:
    """float64 [n, p] dosages with NaN for missing."""
    return _DOSAGE[unpack_codes(bed, n)]


def column_stats(bed: np.ndarray, n: int):
    """mu_j = (n_het + 2 n_hom2) / n_observed ; sigma_inv_j = 1/sqrt(mu_j (1 - mu_j/2)), 1 if 0.

    Same statistic as the reference's `standardize_genotypes!` (`src/wrapper.jl:409-416`):
    binomial, not sample, standard deviation.  Integer counts, one float64 division.
    Returns (mu, sigma_inv, n_missing).
    """
    codes = unpack_codes(bed, n)
    n1 = (codes == 2).sum(axis=0).astype(np.int64)
    n2 = (codes == 3).sum(axis=0).astype(np.int64)
    nmiss = (codes == 1).sum(axis=0).astype(np.int64)
    nobs = n - nmiss
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = (n1 + 2 * n2).astype(np.float64) / nobs.astype(np.float64)
    s = np.sqrt(mu * (1.0 - mu / 2.0))
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma_inv = np.where(s > 0, 1.0 / s, 1.0)
    return mu, sigma_inv, nmiss


class SnpLinAlgOracle:
    """Dense float64 stand-in for SnpLinAlg{Float64}(center, scale, impute) on a packed matrix.

    * getindex path  x_ij = ((missing ? mu_j : g_ij) - mu_j) * sigma_inv_j
      (what `v.x[i, j]` returns at reference `src/utilities.jl:102,735`).
    * mul! path      X'v_j = sigma_inv_j * (sum_i gimp_ij v_i - mu_j * sum_i v_i)
      (what `mul!(v.df, Transpose(x), v.r)` computes at `src/utilities.jl:133`).
    Only for sizes where an n x p float64 matrix fits in host memory.
    """

    def __init__(self, bed: np.ndarray, n: int, center=True, scale=True, impute=True):
        self.bed = bed
        self.n = n
        self.p = bed.shape[0]
        self.center, self.scale, self.impute = center, scale, impute
        self.mu, sinv, self.nmiss = column_stats(bed, n)
        self.sigma_inv = sinv if scale else np.ones_like(sinv)
        g = dosages(bed, n)
        miss = np.isnan(g)
        self.gimp = np.where(miss, self.mu[None, :] if impute else 0.0, g)
        self._x = None

    @property
    def shape(self):
        return (self.n, self.p)

    def dense(self) -> np.ndarray:
        """Standardised matrix via the getindex formula."""
        if self._x is None:
            x = self.gimp.copy()
            if self.center:
                x -= self.mu[None, :]
            x *= self.sigma_inv[None, :]
            self._x = x
        return self._x

    def getindex(self, i, j):
        return self.dense()[i, j]

    def columns(self, cols) -> np.ndarray:
        """x[:, cols] through the getindex formula (n x len(cols))."""
        return self.dense()[:, np.asarray(cols, dtype=np.int64)]

    def xt_v(self, v: np.ndarray) -> np.ndarray:
        """mul!(out, Transpose(x), v); v is [n] or [n, m] -> [p] or [p, m]."""
        out = self.gimp.T @ v
        if self.center:
            s = v.sum(axis=0)
            out = out - (self.mu[:, None] * s[None, :] if v.ndim == 2 else self.mu * s)
        return out * (self.sigma_inv[:, None] if v.ndim == 2 else self.sigma_inv)

    def x_v(self, v: np.ndarray) -> np.ndarray:
        """mul!(out, x, v): out_i = sum_j gimp_ij (sigma_inv_j v_j) - sum_j mu_j sigma_inv_j v_j."""
        w = self.sigma_inv * v
        out = self.gimp @ w
        if self.center:
            out = out - np.dot(self.mu, w)
        return out

    def support_xb(self, idx: np.ndarray, coef: np.ndarray) -> np.ndarray:
        """sum_{j in idx} x[:, j] * coef_j through the getindex path, ascending j
        (reference `update_xb!` memory-efficient branch, `src/utilities.jl:95-106`)."""
        x = self.dense()
        out = np.zeros(self.n)
        for j, cj in zip(idx, coef):
            out += x[:, j] * cj
        return out
