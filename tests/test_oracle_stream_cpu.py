"""CPU checks of the test infrastructure added for the BASELINE-size parity runs: the column-streamed oracle operator
(oracle/stream.py) against the in-memory one, the AVX-512 twin of the synthetic generator against the scalar loop and
the numpy twin, the committed golden answer of configs[4], and the slice arithmetic of the two-phase all-reduce."""
import json
import os

import numpy as np

from conftest import GOLDEN
from mendeliht_jl_b200 import synth
from oracle import cpu as ocpu
from oracle import glm, iht
from oracle.stream import SynthStreamSnpLinAlgCPU


def test_streamed_operator_equals_in_memory_operator():
    n, p, k = 3001, 5000, 6
    y, z, *_ = synth.simulate_response(9, n, p, k, "Normal", n_cov=2, geno_seed=9, missing_rate=0.01)
    a = SynthStreamSnpLinAlgCPU(9, n, p, 0.01, chunk_cols=700)
    b = ocpu.PackedSnpLinAlgCPU(ocpu.synth_columns(9, n, 0, p, 0.01), n)
    ra, rb = iht.fit_iht(y, a, z, k=k + 2), iht.fit_iht(y, b, z, k=k + 2)
    assert ra.iter == rb.iter and np.array_equal(ra.beta, rb.beta) and ra.logl == rb.logl
    assert a.passes == ra.iter + 1                      # one pass over the matrix per sweep (init + one per iteration)
    assert np.array_equal(a.mu, b.mu) and np.array_equal(a.sigma_inv, b.sigma_inv)
    cols = np.array([3, 77, 4999])
    assert np.array_equal(a.columns(cols), b.columns(cols))


def test_generator_twins_agree_in_every_simd_mode():
    lib = ocpu.load()
    for n in (7, 1003, 50000):
        want = synth.packed_columns(2027, n, np.arange(5, 42))
        for level in (0, 1, 2):                         # scalar loop, AVX2 dispatch level, AVX-512 (capped at what the CPU has)
            lib.cpu_set_simd_level(level)
            assert np.array_equal(ocpu.synth_columns(2027, n, 5, 37), want), (n, level)
    lib.cpu_set_simd_level(2)


def test_northstar_golden_is_well_formed():
    g = json.load(open(os.path.join(GOLDEN, "northstar_500000_x_1000000.json")))
    assert g["iter"] == len(g["trace_logl"]) == len(g["trace_backtracks"]) == 5
    assert len(g["support"]) == len(g["beta"]) == 100 and len(g["c"]) == 11
    assert g["support"] == sorted(g["support"]) and all(b != 0 for b in g["beta"])
    assert all(x < y for x, y in zip(g["trace_logl"], g["trace_logl"][1:]))          # ascent
    assert abs(g["logl"] - max(g["trace_logl"])) < 1e-9 * abs(g["logl"])
    assert g["true_positives"] >= 95


def test_two_phase_allreduce_slices_tile_the_vector():
    """p2p.cu k_p2p_allreduce2: rank r reduces elements [2*(pairs*r/R), min(2*(pairs*(r+1)/R), count)) with
    pairs = ceil(count/2): every element exactly once, every slice starts on an even index (16-byte lanes)."""
    for R in range(1, 9):
        for count in (1, 2, 3, 7, 64, 1000, 50001, 262145, 500000, 2000001):
            pairs = (count + 1) // 2
            covered = np.zeros(count, dtype=np.int32)
            for r in range(R):
                lo, hi = 2 * (pairs * r // R), min(2 * (pairs * (r + 1) // R), count)
                assert lo % 2 == 0
                if hi > lo:
                    covered[lo:hi] += 1
            assert np.all(covered == 1), (R, count)
