// Counter-based synthetic PLINK generator shared by device (k_synth) and host code.
// Mirrors the *distribution* of the reference's simulate_random_snparray (src/simulate_utilities.jl:32-50):
// maf_j ~ 0.5*U(0,1) (clipped to [0.01, 0.5]), genotype = Bernoulli(maf) + Bernoulli(maf); optional missing rate.
// Every genotype is a pure function of (seed, global column j, sample i) so CPU and GPU produce identical bytes
// (python twin: mendeliht.jl_b200/synth.py; C twin: oracle/csrc/cpu_ref.c).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define IHTB_HD __host__ __device__ __forceinline__
#else
#define IHTB_HD static inline
#endif

IHTB_HD uint64_t synth_mix64(uint64_t x) {  // splitmix64 finalizer
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

IHTB_HD uint64_t synth_col_key(uint64_t seed, uint64_t j) {
    return synth_mix64(seed ^ ((j + 1) * 0x9E3779B97F4A7C15ull));
}

// threshold T such that P(u32 < T) = maf, maf = clip(0.5 * u, 0.01, 0.5), u = 53-bit uniform
IHTB_HD uint64_t synth_maf_threshold(uint64_t key) {
    uint64_t h = synth_mix64(key ^ 0xA5A5A5A5A5A5A5A5ull);
    double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
    double maf = 0.5 * u;
    if (maf < 0.01) maf = 0.01;
    if (maf > 0.5) maf = 0.5;
    return (uint64_t)(maf * 4294967296.0);
}

IHTB_HD uint32_t synth_missing_threshold(double rate) { return (uint32_t)(rate * 4294967296.0); }

// PLINK 2-bit code of sample i in the column with key `key`
IHTB_HD uint32_t synth_code(uint64_t key, uint64_t thr, uint32_t miss_thr, uint64_t i) {
    uint64_t h = synth_mix64(key ^ (i * 0x9E3779B97F4A7C15ull));
    uint32_t a1 = ((h & 0xFFFFFFFFull) < thr) ? 1u : 0u;
    uint32_t a2 = ((h >> 32) < thr) ? 1u : 0u;
    uint32_t g = a1 + a2;                    // 0, 1, 2
    uint32_t code = g ? g + 1u : 0u;         // 0 -> 00, 1 -> 10, 2 -> 11
    if (miss_thr) {
        uint64_t h2 = synth_mix64(h ^ 0x5851F42D4C957F2Dull);
        if ((uint32_t)(h2 & 0xFFFFFFFFull) < miss_thr) code = 1u;   // 01 = missing
    }
    return code;
}
