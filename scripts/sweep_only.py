"""Runs the X'r sweep alone on a device-generated matrix (for ncu captures and quick kernel timing)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mendeliht_jl_b200 as m

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 500000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
g = m.B200SnpLinAlg.synthetic(n, p, 2024)
mk, mt = C.c_double(0), C.c_double(0)
m._lib.check(m.load().ihtb_sweep_bench(g._h, mode, 2, reps, C.byref(mk), C.byref(mt)))
b = p * ((n + 3) // 4) + 8 * n + 24 * p
print(f"n={n} p={p} mode={mode} kernel_ms={mk.value:.4f} total_ms={mt.value:.4f} kernel_GBs={b / mk.value / 1e6:.1f} "
      f"total_GBs={b / mt.value / 1e6:.1f}")
