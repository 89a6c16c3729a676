"""Oracle (test infrastructure / CPU baseline): packed-genotype operator backed by the C + OpenMP restatement
(oracle/csrc/cpu_ref.c).  Same interface as oracle.snp.SnpLinAlgOracle (shape, mu, sigma_inv, xt_v, support_xb), so
oracle.iht.fit_iht / oracle.cv.cv_iht run unchanged on matrices too large to decode densely.  This is what
bench.py's `cpu_baseline` and `--impl reference` legs time; the product never imports it."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libihtcpu.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise ImportError(f"{LIB} missing: run `make -C oracle`")
        lib = C.CDLL(LIB)
        lib.cpu_num_threads.restype = C.c_int
        lib.cpu_set_num_threads.restype = C.c_int
        _lib = lib
        # torchrun exports OMP_NUM_THREADS=1 to its workers: the baseline uses every core this process may run on
        # unless IHTCPU_THREADS says otherwise (round-1 SCALE reference arm ran on one core because of this)
        set_threads(int(os.environ.get("IHTCPU_THREADS", "0")) or available_cores())
    return _lib


def available_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(t: int) -> int:
    """OpenMP threads used by every kernel of the C restatement; returns the count in effect."""
    return int(load().cpu_set_num_threads(C.c_int(int(t)))) if _lib is not None else int(t)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def synth_columns(seed: int, n: int, j0: int, ncols: int, missing_rate: float = 0.0) -> np.ndarray:
    out = np.empty((ncols, (n + 3) // 4), dtype=np.uint8)
    load().cpu_synth(C.c_uint64(seed), C.c_int64(n), C.c_int64(j0), C.c_int64(ncols), C.c_double(missing_rate),
                     _p(out, C.c_uint8))
    return out


class PackedSnpLinAlgCPU:
    def __init__(self, bed: np.ndarray, n: int):
        self.bed = np.ascontiguousarray(bed, dtype=np.uint8)
        self.n, self.p = int(n), int(bed.shape[0])
        self.stride = int(bed.shape[1])
        self.mu = np.empty(self.p); self.sigma_inv = np.empty(self.p); self.nmiss = np.empty(self.p, dtype=np.int64)
        load().cpu_col_stats(_p(self.bed, C.c_uint8), C.c_int64(self.n), C.c_int64(self.p), C.c_int64(self.stride),
                             _p(self.mu, C.c_double), _p(self.sigma_inv, C.c_double), _p(self.nmiss, C.c_int64))
        self.threads = int(load().cpu_num_threads())

    @property
    def shape(self):
        return (self.n, self.p)

    def columns(self, cols) -> np.ndarray:
        """x[:, cols] through the getindex formula ((g or mu_j when missing) - mu_j) * sigma_inv_j, n x len(cols);
        same arithmetic as oracle.snp.SnpLinAlgOracle.dense() on the selected columns."""
        cols = np.asarray(cols, dtype=np.int64).reshape(-1)
        out = np.empty((self.n, cols.shape[0]))
        for t, j in enumerate(cols):
            b = self.bed[j]
            codes = np.empty(self.stride * 4, dtype=np.uint8)
            for s in range(4):
                codes[s::4] = (b >> (2 * s)) & 3
            g = np.array([0.0, self.mu[j], 1.0, 2.0])[codes[: self.n]]
            out[:, t] = (g - self.mu[j]) * self.sigma_inv[j]
        return out

    def xt_v(self, v: np.ndarray) -> np.ndarray:
        v = np.asarray(v, dtype=np.float64)
        one = v.ndim == 1
        vm = np.asfortranarray(v.reshape(self.n, -1))
        m = vm.shape[1]
        out = np.empty((self.p, m), order="F")
        load().cpu_xt_v(_p(self.bed, C.c_uint8), C.c_int64(self.n), C.c_int64(self.p), C.c_int64(self.stride),
                        _p(self.mu, C.c_double), _p(self.sigma_inv, C.c_double), _p(vm, C.c_double), C.c_int64(m),
                        _p(out, C.c_double))
        return out[:, 0].copy() if one else out

    def support_xb(self, idx, coef) -> np.ndarray:
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        out = np.empty(self.n)
        load().cpu_x_support(_p(self.bed, C.c_uint8), C.c_int64(self.n), C.c_int64(self.stride),
                             _p(self.mu, C.c_double), _p(self.sigma_inv, C.c_double), _p(idx, C.c_int64),
                             C.c_int64(idx.shape[0]), _p(coef, C.c_double), _p(out, C.c_double))
        return out
