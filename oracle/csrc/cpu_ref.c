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

void cpu_synth(uint64_t seed, int64_t n, int64_t j0, int64_t ncols, double missing_rate, uint8_t* out) {
    const int64_t nbytes = (n + 3) / 4;
    const uint32_t miss_thr = synth_missing_threshold(missing_rate);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncols; ++c) {
        uint64_t key = synth_col_key(seed, (uint64_t)(j0 + c));
        uint64_t thr = synth_maf_threshold(key);
        uint8_t* col = out + c * nbytes;
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

/* out_j = sinv_j * (sum_i gimp_ij v_i - mu_j * sum_i v_i) for m right-hand sides (V is n x m column-major) */
void cpu_xt_v(const uint8_t* bed, int64_t n, int64_t p, int64_t stride, const double* mu, const double* sinv,
              const double* V, int64_t m, double* out) {
    for (int64_t t = 0; t < m; ++t) {
        const double* v = V + t * n;
        double vsum = 0.0;
        for (int64_t i = 0; i < n; ++i) vsum += v[i];
        const int64_t nfull = n >> 2;
#pragma omp parallel for schedule(static)
        for (int64_t j = 0; j < p; ++j) {
            const uint8_t* col = bed + j * stride;
            const double mj = mu[j];
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
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = 0.0;
    for (int64_t c = 0; c < k; ++c) {
        const int64_t j = idx[c];
        const uint8_t* col = bed + j * stride;
        const double mj = mu[j], sj = sinv[j], cj = coef[c];
        double tab[4];
        const double g[4] = {0.0, mj, 1.0, 2.0};
        for (int q = 0; q < 4; ++q) tab[q] = ((g[q] - mj) * sj) * cj;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) out[i] += tab[(col[i >> 2] >> (2 * (i & 3))) & 3u];
    }
}
