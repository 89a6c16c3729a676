/* Oracle (test infrastructure / CPU baseline only): C + OpenMP restatement of the SnpArrays.jl SnpLinAlg kernels that
 * MendelIHT calls on its hot path.  SnpArrays.jl is an un-vendored dependency of the reference (Project.toml:20,35),
 * so its published algorithm is restated: column-wise decode of the 2-bit codes with mean imputation, centring and
 * scaling applied after the dot product.
 *   cpu_xt_v      <- mul!(out, Transpose(x), v)      (called at reference src/utilities.jl:133)
 *   cpu_x_support <- the getindex loops of update_xb! / iht_stepsize! (src/utilities.jl:95-111, 728-743)
 *   cpu_col_stats <- SnpLinAlg constructor statistics (same as standardize_genotypes!, src/wrapper.jl:406-423)
 *   cpu_synth     <- twin of the device generator (mendeliht.jl_b200/csrc/synth.cuh)
 * Never linked into the product library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../../mendeliht.jl_b200/csrc/synth.cuh"

int cpu_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline must say how many threads it really uses */
int cpu_set_num_threads(int t) {
#ifdef _OPENMP
    if (t >= 1) omp_set_num_threads(t);
    return omp_get_max_threads();
#else
    (void)t;
    return 1;
#endif
}

static int simd_level(void);

/* AVX-512 form of the generator for the common case (no missing data): 8 samples per step, same integer arithmetic as
 * synth_code (splitmix64 finalizer, two 32-bit threshold compares), so the bytes are identical to the scalar loop. */
#include <immintrin.h>
__attribute__((target("avx512f,avx512dq,avx512bw")))
static void synth_col_avx512(uint64_t key, uint64_t thr, int64_t n, uint8_t* col) {
    const __m512i gold = _mm512_set1_epi64((long long)0x9E3779B97F4A7C15ull);
    const __m512i m1 = _mm512_set1_epi64((long long)0xBF58476D1CE4E5B9ull), m2 = _mm512_set1_epi64((long long)0x94D049BB133111EBull);
    const __m512i vkey = _mm512_set1_epi64((long long)key), vthr = _mm512_set1_epi64((long long)thr);
    const __m512i lo32 = _mm512_set1_epi64(0xFFFFFFFFll);
    const __m512i shl = _mm512_setr_epi64(0, 2, 4, 6, 8, 10, 12, 14);
    const __m512i one = _mm512_set1_epi64(1);
    __m512i idx = _mm512_setr_epi64(0, 1, 2, 3, 4, 5, 6, 7);
    const __m512i eight = _mm512_set1_epi64(8);
    const int64_t nfull = n / 8;
    uint16_t* out16 = (uint16_t*)col;
    for (int64_t b = 0; b < nfull; ++b) {
        __m512i x = _mm512_xor_si512(vkey, _mm512_mullo_epi64(idx, gold));
        x = _mm512_mullo_epi64(_mm512_xor_si512(x, _mm512_srli_epi64(x, 30)), m1);
        x = _mm512_mullo_epi64(_mm512_xor_si512(x, _mm512_srli_epi64(x, 27)), m2);
        x = _mm512_xor_si512(x, _mm512_srli_epi64(x, 31));
        const __mmask8 a1 = _mm512_cmplt_epu64_mask(_mm512_and_si512(x, lo32), vthr);
        const __mmask8 a2 = _mm512_cmplt_epu64_mask(_mm512_srli_epi64(x, 32), vthr);
        /* g = a1 + a2 in {0,1,2}; code = g ? g + 1 : 0  ->  bit1 = a1|a2, bit0 = a1&a2 */
        __m512i code = _mm512_maskz_mov_epi64((__mmask8)(a1 | a2), _mm512_set1_epi64(2));
        code = _mm512_mask_or_epi64(code, (__mmask8)(a1 & a2), code, one);
        out16[b] = (uint16_t)_mm512_reduce_or_epi64(_mm512_sllv_epi64(code, shl));
        idx = _mm512_add_epi64(idx, eight);
    }
    for (int64_t i = 8 * nfull; i < n; ++i) {          /* tail: fewer than 8 samples */
        uint32_t c = synth_code(key, thr, 0u, (uint64_t)i);
        if ((i & 3) == 0) col[i >> 2] = 0;
        col[i >> 2] |= (uint8_t)(c << (2 * (i & 3)));
    }
}

void cpu_synth(uint64_t seed, int64_t n, int64_t j0, int64_t ncols, double missing_rate, uint8_t* out) {
    const int64_t nbytes = (n + 3) / 4;
    const uint32_t miss_thr = synth_missing_threshold(missing_rate);
    const int fast = simd_level() == 2 && miss_thr == 0;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncols; ++c) {
        uint64_t key = synth_col_key(seed, (uint64_t)(j0 + c));
        uint64_t thr = synth_maf_threshold(key);
        uint8_t* col = out + c * nbytes;
        if (fast) { synth_col_avx512(key, thr, n, col); continue; }
        for (int64_t b = 0; b < nbytes; ++b) {
            uint32_t byte = 0;
            for (int s = 0; s < 4; ++s) {
                int64_t i = 4 * b + s;
                if (i < n) byte |= synth_code(key, thr, miss_thr, (uint64_t)i) << (2 * s);
            }
            col[b] = (uint8_t)byte;
        }
    }
}

void cpu_col_stats(const uint8_t* bed, int64_t n, int64_t p, int64_t stride, double* mu, double* sinv,
                   int64_t* nmiss) {
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < p; ++j) {
        const uint8_t* col = bed + j * stride;
        int64_t c1 = 0, c2 = 0, cm = 0;
        for (int64_t i = 0; i < n; ++i) {
            uint32_t code = (col[i >> 2] >> (2 * (i & 3))) & 3u;
            c1 += code == 2; c2 += code == 3; cm += code == 1;
        }
        double m = (double)(c1 + 2 * c2) / (double)(n - cm);
        double s = sqrt(m * (1.0 - m / 2.0));
        mu[j] = m;
        sinv[j] = s > 0.0 ? 1.0 / s : 1.0;
        nmiss[j] = cm;
    }
}

/* ---- SIMD column dot for the CPU baseline ---------------------------------------------------------------------
 * sum_i dosage_ij v_i over the first nb packed bytes of a column (code 01 counts 0) and, separately, the sum of v over
 * the missing samples (the mean-imputation term mu_j * sum_{missing} v_i is added by the caller).  Same decode + FMA
 * per genotype as SnpArrays' SnpLinAlg kernels (LoopVectorization-tiled in the reference), written with explicit
 * AVX2 / AVX-512 intrinsics and dispatched at run time, because gcc does not vectorise the table-lookup loop below.
 * FP64 accumulation throughout; only the order of the additions differs from the scalar loop. */
static inline double miss16(uint32_t w, const double* vv) {      /* w: 16 bits = 8 samples; sum of v over codes 01 */
    uint32_t m = w & ~(w >> 1) & 0x5555u;
    double s = 0.0;
    while (m) {
        int bit = __builtin_ctz(m);
        s += vv[bit >> 1];
        m &= m - 1;
    }
    return s;
}

__attribute__((target("avx2,fma")))
static double col_dot_avx2(const uint8_t* col, const double* v, int64_t nb, double* miss_sum) {
    const __m256i sh = _mm256_setr_epi32(0, 2, 4, 6, 8, 10, 12, 14);
    const __m256i three = _mm256_set1_epi32(3);
    const __m256 tab = _mm256_setr_ps(0.f, 0.f, 1.f, 2.f, 0.f, 0.f, 1.f, 2.f);
    __m256d a0 = _mm256_setzero_pd(), a1 = _mm256_setzero_pd(), a2 = _mm256_setzero_pd(), a3 = _mm256_setzero_pd();
    double ms = 0.0;
    int64_t b = 0;
    for (; b + 4 <= nb; b += 4) {
        uint32_t w0 = (uint32_t)col[b] | ((uint32_t)col[b + 1] << 8);
        uint32_t w1 = (uint32_t)col[b + 2] | ((uint32_t)col[b + 3] << 8);
        const double* vv = v + 4 * b;
        if ((w0 & ~(w0 >> 1) & 0x5555u) | (w1 & ~(w1 >> 1) & 0x5555u)) ms += miss16(w0, vv) + miss16(w1, vv + 8);
        __m256i c0 = _mm256_and_si256(_mm256_srlv_epi32(_mm256_set1_epi32((int)w0), sh), three);
        __m256i c1 = _mm256_and_si256(_mm256_srlv_epi32(_mm256_set1_epi32((int)w1), sh), three);
        __m256 d0 = _mm256_permutevar_ps(tab, c0), d1 = _mm256_permutevar_ps(tab, c1);
        a0 = _mm256_fmadd_pd(_mm256_cvtps_pd(_mm256_castps256_ps128(d0)), _mm256_loadu_pd(vv), a0);
        a1 = _mm256_fmadd_pd(_mm256_cvtps_pd(_mm256_extractf128_ps(d0, 1)), _mm256_loadu_pd(vv + 4), a1);
        a2 = _mm256_fmadd_pd(_mm256_cvtps_pd(_mm256_castps256_ps128(d1)), _mm256_loadu_pd(vv + 8), a2);
        a3 = _mm256_fmadd_pd(_mm256_cvtps_pd(_mm256_extractf128_ps(d1, 1)), _mm256_loadu_pd(vv + 12), a3);
    }
    __m256d t = _mm256_add_pd(_mm256_add_pd(a0, a1), _mm256_add_pd(a2, a3));
    double lanes[4];
    _mm256_storeu_pd(lanes, t);
    double acc = (lanes[0] + lanes[1]) + (lanes[2] + lanes[3]);
    for (; b < nb; ++b) {
        uint32_t byte = col[b];
        for (int s = 0; s < 4; ++s) {
            uint32_t code = (byte >> (2 * s)) & 3u;
            double x = v[4 * b + s];
            if (code == 1u) ms += x; else if (code == 2u) acc += x; else if (code == 3u) acc += 2.0 * x;
        }
    }
    *miss_sum = ms;
    return acc;
}

__attribute__((target("avx512f,avx512bw,avx512dq,avx2,fma")))
static double col_dot_avx512(const uint8_t* col, const double* v, int64_t nb, double* miss_sum) {
    const __m512i sh = _mm512_setr_epi32(0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 30);
    const __m512i three = _mm512_set1_epi32(3);
    const __m512 tab = _mm512_setr_ps(0.f, 0.f, 1.f, 2.f, 0.f, 0.f, 1.f, 2.f, 0.f, 0.f, 1.f, 2.f, 0.f, 0.f, 1.f, 2.f);
    __m512d a0 = _mm512_setzero_pd(), a1 = _mm512_setzero_pd(), a2 = _mm512_setzero_pd(), a3 = _mm512_setzero_pd();
    double ms = 0.0;
    int64_t b = 0;
    for (; b + 8 <= nb; b += 8) {
        uint32_t w0, w1;
        memcpy(&w0, col + b, 4); memcpy(&w1, col + b + 4, 4);
        const double* vv = v + 4 * b;
        if ((w0 & ~(w0 >> 1) & 0x55555555u) | (w1 & ~(w1 >> 1) & 0x55555555u))
            ms += miss16(w0 & 0xFFFFu, vv) + miss16(w0 >> 16, vv + 8) + miss16(w1 & 0xFFFFu, vv + 16) + miss16(w1 >> 16, vv + 24);
        __m512i c0 = _mm512_and_si512(_mm512_srlv_epi32(_mm512_set1_epi32((int)w0), sh), three);
        __m512i c1 = _mm512_and_si512(_mm512_srlv_epi32(_mm512_set1_epi32((int)w1), sh), three);
        __m512 d0 = _mm512_permutexvar_ps(c0, tab), d1 = _mm512_permutexvar_ps(c1, tab);
        a0 = _mm512_fmadd_pd(_mm512_cvtps_pd(_mm512_castps512_ps256(d0)), _mm512_loadu_pd(vv), a0);
        a1 = _mm512_fmadd_pd(_mm512_cvtps_pd(_mm512_extractf32x8_ps(d0, 1)), _mm512_loadu_pd(vv + 8), a1);
        a2 = _mm512_fmadd_pd(_mm512_cvtps_pd(_mm512_castps512_ps256(d1)), _mm512_loadu_pd(vv + 16), a2);
        a3 = _mm512_fmadd_pd(_mm512_cvtps_pd(_mm512_extractf32x8_ps(d1, 1)), _mm512_loadu_pd(vv + 24), a3);
    }
    double acc = _mm512_reduce_add_pd(_mm512_add_pd(_mm512_add_pd(a0, a1), _mm512_add_pd(a2, a3)));
    for (; b < nb; ++b) {
        uint32_t byte = col[b];
        for (int s = 0; s < 4; ++s) {
            uint32_t code = (byte >> (2 * s)) & 3u;
            double x = v[4 * b + s];
            if (code == 1u) ms += x; else if (code == 2u) acc += x; else if (code == 3u) acc += 2.0 * x;
        }
    }
    *miss_sum = ms;
    return acc;
}

/* 0: scalar table loop, 1: AVX2, 2: AVX-512 (env IHTCPU_SIMD=0/1/2 overrides the detection) */
static int g_simd_level = -1;
static int simd_level(void) {
    int level = g_simd_level;
    if (level < 0) {
        int l = 0;
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) l = 1;
        if (l == 1 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
            __builtin_cpu_supports("avx512dq")) l = 2;
        const char* e = getenv("IHTCPU_SIMD");
        if (e && *e >= '0' && *e <= '2' && (*e - '0') <= l) l = *e - '0';
        level = g_simd_level = l;
    }
    return level;
}
int cpu_simd_level(void) { return simd_level(); }
/* tests: force a lower level (never above what the CPU supports); returns the level in effect */
int cpu_set_simd_level(int l) {
    g_simd_level = -1;
    int max = simd_level();
    if (l >= 0 && l < max) g_simd_level = l;
    return g_simd_level;
}

/* out_j = sinv_j * (sum_i gimp_ij v_i - mu_j * sum_i v_i) for m right-hand sides (V is n x m column-major) */
void cpu_xt_v(const uint8_t* bed, int64_t n, int64_t p, int64_t stride, const double* mu, const double* sinv,
              const double* V, int64_t m, double* out) {
    for (int64_t t = 0; t < m; ++t) {
        const double* v = V + t * n;
        double vsum = 0.0;
        for (int64_t i = 0; i < n; ++i) vsum += v[i];
        const int64_t nfull = n >> 2;
        const int simd = simd_level();
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < p; ++j) {
            const uint8_t* col = bed + j * stride;
            const double mj = mu[j];
            if (simd > 0) {
                double ms = 0.0;
                double acc = simd == 2 ? col_dot_avx512(col, v, nfull, &ms) : col_dot_avx2(col, v, nfull, &ms);
                for (int64_t i = 4 * nfull; i < n; ++i) {
                    uint32_t code = (col[i >> 2] >> (2 * (i & 3))) & 3u;
                    if (code == 1u) ms += v[i]; else if (code == 2u) acc += v[i]; else if (code == 3u) acc += 2.0 * v[i];
                }
                out[j + t * p] = sinv[j] * ((acc + mj * ms) - mj * vsum);
                continue;
            }
            /* per-column 4-entry dosage table; code 01 (missing) is imputed with the column mean */
            const double tab[4] = {0.0, mj, 1.0, 2.0};
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int64_t b = 0; b < nfull; ++b) {
                uint32_t byte = col[b];
                const double* vv = v + 4 * b;
                a0 += tab[byte & 3u] * vv[0];
                a1 += tab[(byte >> 2) & 3u] * vv[1];
                a2 += tab[(byte >> 4) & 3u] * vv[2];
                a3 += tab[(byte >> 6) & 3u] * vv[3];
            }
            double acc = (a0 + a1) + (a2 + a3);
            for (int64_t i = 4 * nfull; i < n; ++i) acc += tab[(col[i >> 2] >> (2 * (i & 3))) & 3u] * v[i];
            out[j + t * p] = sinv[j] * (acc - mj * vsum);
        }
    }
}

/* out_i = sum_c x[i, idx_c] * coef_c with x_ij = ((missing ? mu : g) - mu) * sinv, ascending c */
void cpu_x_support(const uint8_t* bed, int64_t n, int64_t stride, const double* mu, const double* sinv,
                   const int64_t* idx, int64_t k, const double* coef, double* out) {
    /* one parallel region over sample blocks; inside a block the columns are added in ascending order, so every
     * out[i] sees the same sequence of additions as a column-by-column loop */
    double* tabs = (double*)malloc((size_t)(k > 0 ? k : 1) * 4 * sizeof(double));
    for (int64_t c = 0; c < k; ++c) {
        const int64_t j = idx[c];
        const double mj = mu[j], sj = sinv[j], cj = coef[c];
        const double g[4] = {0.0, mj, 1.0, 2.0};
        for (int q = 0; q < 4; ++q) tabs[4 * c + q] = ((g[q] - mj) * sj) * cj;
    }
    const int64_t blk = 4096;                       /* multiple of 4: a block starts on a byte boundary */
    const int64_t nblk = (n + blk - 1) / blk;
#pragma omp parallel for schedule(static)
    for (int64_t bi = 0; bi < nblk; ++bi) {
        const int64_t i0 = bi * blk, i1 = (i0 + blk < n) ? i0 + blk : n;
        for (int64_t i = i0; i < i1; ++i) out[i] = 0.0;
        for (int64_t c = 0; c < k; ++c) {
            const uint8_t* col = bed + idx[c] * stride;
            const double* tab = tabs + 4 * c;
            for (int64_t i = i0; i < i1; ++i) out[i] += tab[(col[i >> 2] >> (2 * (i & 3))) & 3u];
        }
    }
    free(tabs);
}
