"""Oracle (test infrastructure): column-streamed genotype operator for matrices larger than host memory.

Same interface as oracle.cpu.PackedSnpLinAlgCPU (shape, mu, sigma_inv, xt_v, support_xb, columns), but the packed
matrix is never held: every pass regenerates the synthetic PLINK columns chunk by chunk with the C twin of the device
generator (oracle/csrc/cpu_ref.c `cpu_synth`) and runs the C + OpenMP kernels on the chunk.  Used once, offline, to
produce the golden answer of BASELINE configs[4] (n = 500k x p = 1M, 125 GB packed): scripts/make_northstar_golden.py.
Restates SnpArrays' SnpLinAlg semantics exactly like oracle/cpu.py (mul! at reference src/utilities.jl:133)."""
from __future__ import annotations

import numpy as np

from . import cpu as ocpu


class SynthStreamSnpLinAlgCPU:
    def __init__(self, seed: int, n: int, p: int, missing_rate: float = 0.0, chunk_cols: int = 20000, verbose=False):
        self.seed, self.n, self.p, self.missing_rate = int(seed), int(n), int(p), float(missing_rate)
        self.chunk = int(chunk_cols)
        self.verbose = verbose
        self.mu = np.empty(self.p)
        self.sigma_inv = np.empty(self.p)
        self.nmiss = np.empty(self.p, dtype=np.int64)
        self.threads = 1
        self.passes = 0
        self._have_stats = False

    @property
    def shape(self):
        return (self.n, self.p)

    def _chunk_op(self, j0, j1):
        bed = ocpu.synth_columns(self.seed, self.n, j0, j1 - j0, self.missing_rate)
        x = ocpu.PackedSnpLinAlgCPU(bed, self.n) if not self._have_stats else _with_stats(bed, self.n, self.mu[j0:j1],
                                                                                          self.sigma_inv[j0:j1],
                                                                                          self.nmiss[j0:j1])
        self.threads = x.threads
        return x

    def xt_v(self, v: np.ndarray) -> np.ndarray:
        """mul!(out, Transpose(x), v), one pass over regenerated column chunks; the first pass also fills mu / sigma_inv."""
        v = np.asarray(v, dtype=np.float64)
        out = np.empty((self.p,) + v.shape[1:], order="F")
        for j0 in range(0, self.p, self.chunk):
            j1 = min(j0 + self.chunk, self.p)
            x = self._chunk_op(j0, j1)
            if not self._have_stats:
                self.mu[j0:j1], self.sigma_inv[j0:j1], self.nmiss[j0:j1] = x.mu, x.sigma_inv, x.nmiss
            out[j0:j1] = x.xt_v(v)
            if self.verbose and (j0 // self.chunk) % 10 == 0:
                print(f"  pass {self.passes}: columns {j1}/{self.p}", flush=True)
        self._have_stats = True
        self.passes += 1
        return out

    def ensure_stats(self):
        if not self._have_stats:
            self.xt_v(np.zeros(self.n))

    def _cols(self, cols):
        cols = np.asarray(cols, dtype=np.int64).reshape(-1)
        self.ensure_stats()
        bed = np.concatenate([ocpu.synth_columns(self.seed, self.n, int(j), 1, self.missing_rate) for j in cols]) \
            if cols.size else np.zeros((0, (self.n + 3) // 4), dtype=np.uint8)
        return _with_stats(bed, self.n, self.mu[cols], self.sigma_inv[cols], self.nmiss[cols])

    def support_xb(self, idx, coef) -> np.ndarray:
        idx = np.asarray(idx, dtype=np.int64)
        x = self._cols(idx)
        return x.support_xb(np.arange(idx.shape[0]), np.asarray(coef, dtype=np.float64))

    def columns(self, cols) -> np.ndarray:
        cols = np.asarray(cols, dtype=np.int64).reshape(-1)
        return self._cols(cols).columns(np.arange(cols.shape[0]))


def _with_stats(bed, n, mu, sinv, nmiss):
    """PackedSnpLinAlgCPU over `bed` with known column statistics (skips the statistics pass)."""
    x = ocpu.PackedSnpLinAlgCPU.__new__(ocpu.PackedSnpLinAlgCPU)
    x.bed = np.ascontiguousarray(bed, dtype=np.uint8)
    x.n, x.p, x.stride = int(n), int(bed.shape[0]), int(bed.shape[1])
    x.mu = np.ascontiguousarray(mu, dtype=np.float64)
    x.sigma_inv = np.ascontiguousarray(sinv, dtype=np.float64)
    x.nmiss = np.ascontiguousarray(nmiss, dtype=np.int64)
    x.threads = int(ocpu.load().cpu_num_threads())
    return x
