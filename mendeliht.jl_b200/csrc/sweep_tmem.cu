// FAST / PAIR X'v sweep, third generation: the genotype stream reaches the registers through TENSOR MEMORY instead of
// shared-memory loads.
//
// Why.  The table sweeps are bound by the LSU data pipe of the SM, not by HBM (ncu, profiles/): per 128-byte column
// chunk the stage-ring kernel (sweep_lut.cu) spends 4 wavefronts on table lookups -- irreducible: one lookup per
// packed byte, lane = word position -- plus one for the TMA write of the chunk into shared memory and one for reading
// it back into registers.  Round 2 showed that ANY global -> register path through L1TEX costs those two data-pipe
// passes (profiles/r2_rejected).  Tensor memory has its own datapaths: tcgen05.cp copies shared memory -> TMEM through
// the tensor-core operand path (no LSU wavefront) and tcgen05.ld moves TMEM -> registers (no LSU wavefront either), so
// a chunk costs 4 lookups + 1 TMA write = 5 LSU wavefronts instead of 6.  No MMA is issued: TMEM is used as a
// 256 KB register-side staging buffer, which also deepens the pipeline (8 units of 32 KB in flight per SM).
// RESULT (see sweep_tmem_enabled below): correct, but not faster -- the copy's shared-memory reads still compete for
// the same data RAM, so this kernel is opt-in and the stage ring stays the default.
//
// Data path per unit of 256 columns x 512 samples (64 quads of the quad-interleaved layout, common.cuh):
//   producer warp   : cp.async.bulk, 2 x 16 KB stages, HBM -> shared memory (mbarrier complete_tx)
//   copy warp       : per stage 4 x tcgen05.cp.128x256b, SWIZZLE_NONE descriptor with SBO = 128, LBO = 2048: row r of the
//                     copy = (lane group r / 32, word r % 32), its 32 bytes = word `r % 32` of the quads g and g + 4 of
//                     the 4 KB block -- exactly the quad layout (mapping measured with scripts/probe/tmem_probe.cu);
//                     tcgen05.commit hands the shared-memory stage back and, per unit, signals the consumers
//   16 consumer warps: warp (g = warp % 4, j = warp / 4) reads TMEM lanes 32 g .. 32 g + 31 (its lane = word position,
//                     so table lookups stay bank-conflict free), 16 consecutive TMEM columns = 16 genotype columns,
//                     with one tcgen05.ld.32x32b.x16; then the lookups, butterfly and FP32 slab partials of sweep_lut.cu.
// Table build, error bounds, output layout and the PAIR (half2, two right-hand sides) form are those of sweep_lut.cu.
#include "lut_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace ihtb {

constexpr int TM_CW = 16;                          // consumer warps
constexpr int TM_THREADS = (TM_CW + 2) * 32;       // + TMA producer warp + tcgen05.cp warp
constexpr int TM_STAGE_COLS = 128;
constexpr int TM_STAGE_BYTES = TM_STAGE_COLS * 128;
constexpr int TM_UNIT_COLS = 256;                  // 2 stages
constexpr int TM_SLOTS = 8;                        // TMEM unit slots of 64 columns (512 columns in all)
constexpr int TM_MAX_STAGES = 6;
constexpr int TM_SMEM_BYTES = 232448;

__device__ __forceinline__ void mbar_arrive_plain(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_cp_128x256b(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

// shared-memory plan: 128 KB of tables at a 64 KB boundary, 16 KB stages below and above, barriers in the first gap
struct TmPlan {
    uint32_t bars;                 // smem_full[6] smem_empty[6] tmem_full[8] tmem_empty[8] (8 bytes each) + tmem base word
    uint32_t tab, lo0, hi0;
    int n_lo, n_stages;
    __device__ __forceinline__ uint32_t stage(int s) const {
        return (s < n_lo) ? lo0 + (uint32_t)s * TM_STAGE_BYTES : hi0 + (uint32_t)(s - n_lo) * TM_STAGE_BYTES;
    }
    __device__ __forceinline__ uint32_t smem_full(int s) const { return bars + 8u * (uint32_t)s; }
    __device__ __forceinline__ uint32_t smem_empty(int s) const { return bars + 8u * (uint32_t)(TM_MAX_STAGES + s); }
    __device__ __forceinline__ uint32_t tmem_full(int t) const { return bars + 8u * (uint32_t)(2 * TM_MAX_STAGES + t); }
    __device__ __forceinline__ uint32_t tmem_empty(int t) const { return bars + 8u * (uint32_t)(2 * TM_MAX_STAGES + TM_SLOTS + t); }
    __device__ __forceinline__ uint32_t tmem_word() const { return bars + 8u * (uint32_t)(2 * TM_MAX_STAGES + 2 * TM_SLOTS); }
};
__device__ __forceinline__ TmPlan tm_plan(uint32_t base, uint32_t bytes) {
    TmPlan pl;
    const uint32_t end = base + bytes;
    pl.tab = (base + 65535u) & ~65535u;
    uint32_t lo = (base + 1023u) & ~1023u;                    // stages are 1 KB aligned (matrix descriptors: 16-byte units)
    pl.bars = lo; lo += 1024u;
    pl.lo0 = lo;
    pl.n_lo = (pl.tab > lo) ? (int)((pl.tab - lo) / TM_STAGE_BYTES) : 0;
    pl.hi0 = pl.tab + LUT_TABLE_BYTES;
    const int n_hi = (end > pl.hi0) ? (int)((end - pl.hi0) / TM_STAGE_BYTES) : 0;
    pl.n_stages = pl.n_lo + n_hi;
    if (pl.n_stages > TM_MAX_STAGES) pl.n_stages = TM_MAX_STAGES;
    return pl;
}

template <bool H2>
__global__ void __launch_bounds__(TM_THREADS, 1)
k_sweep_tmem(const uint8_t* __restrict__ bed, int64_t cs_s, int64_t p4, int64_t p_out, int64_t n, int64_t n_slabs,
             const double* __restrict__ v, const double* __restrict__ v1, const double* __restrict__ vbar_p,
             const float* __restrict__ scale_p, float* __restrict__ part, uint32_t dyn_bytes) {
    const double vbar = vbar_p[0];
    const double vbar1 = H2 ? vbar_p[1] : 0.0;
    const float sc0 = H2 ? scale_p[0] : 1.0f, sc1 = H2 ? scale_p[1] : 1.0f;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const TmPlan pl = tm_plan(smem_u32(smem_raw), dyn_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = pl.n_stages;

    const int64_t ncb = (p4 + TM_UNIT_COLS - 1) / TM_UNIT_COLS;          // 256-column blocks per slab
    const int64_t units = n_slabs * ncb;
    const int64_t u_beg = units * (int64_t)blockIdx.x / gridDim.x;
    const int64_t u_end = units * (int64_t)(blockIdx.x + 1) / gridDim.x;
    const int64_t slab_beg = u_beg / ncb;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(pl.smem_full(s), 1); mbar_init(pl.smem_empty(s), 1); }
        for (int t = 0; t < TM_SLOTS; ++t) { mbar_init(pl.tmem_full(t), 1); mbar_init(pl.tmem_empty(t), TM_CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TM_CW + 1) {          // the copy warp owns the tensor-memory allocation: all 512 columns (1 CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(pl.tmem_word()) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = lds_u32(pl.tmem_word());

    if (u_beg < u_end) {
        if (warp == TM_CW) {
            // ===== TMA producer: two 16 KB stages per unit =====
            if (lane == 0) {
                int st = 0; uint32_t ph = 0;
                int64_t slab = slab_beg, cb = u_beg - slab_beg * ncb;
                for (int64_t u = u_beg; u < u_end; ++u) {
                    for (int h = 0; h < 2; ++h) {
                        const int64_t j0 = cb * TM_UNIT_COLS + (int64_t)h * TM_STAGE_COLS;
                        int64_t ncols = p4 - j0;
                        ncols = ncols < 0 ? 0 : (ncols > TM_STAGE_COLS ? TM_STAGE_COLS : ncols);
                        mbar_wait(pl.smem_empty(st), ph ^ 1u);
                        if (ncols > 0) {
                            mbar_expect_tx(pl.smem_full(st), (uint32_t)ncols * 128u);
                            bulk_g2s(pl.stage(st), bed + slab * cs_s + j0 * 128, (uint32_t)ncols * 128u, pl.smem_full(st));
                        } else {
                            mbar_arrive_plain(pl.smem_full(st));         // nothing to load: the stage keeps stale bytes
                        }
                        if (++st == S) { st = 0; ph ^= 1u; }
                    }
                    if (++cb == ncb) { cb = 0; ++slab; }
                }
            }
        } else if (warp == TM_CW + 1) {
            // ===== copy warp: shared memory -> tensor memory, 4 x 4 KB per stage =====
            if (lane == 0) {
                int st = 0; uint32_t ph = 0;          // shared-memory stage ring
                int ts = 0; uint32_t tph = 0;         // tensor-memory unit slots
                // SWIZZLE_NONE matrix descriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
                const uint64_t desc_hi = ((uint64_t)(2048u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
                for (int64_t u = u_beg; u < u_end; ++u) {
                    mbar_wait(pl.tmem_empty(ts), tph ^ 1u);
                    for (int h = 0; h < 2; ++h) {
                        mbar_wait(pl.smem_full(st), ph);
                        tc_fence_after();
                        const uint32_t sbase = pl.stage(st);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            tc_cp_128x256b(tbase + (uint32_t)(64 * ts + 32 * h + 8 * c),
                                           desc_hi | (uint64_t)(((sbase + 4096u * c) >> 4) & 0x3FFFu));
                        tc_commit(pl.smem_empty(st));             // the stage is free once these copies have read it
                        if (++st == S) { st = 0; ph ^= 1u; }
                    }
                    tc_commit(pl.tmem_full(ts));                  // all 8 copies of the unit have landed
                    if (++ts == TM_SLOTS) { ts = 0; tph ^= 1u; }
                }
            }
        } else {
            // ===== consumer warps =====
            const int tid = threadIdx.x;              // 0 .. 511
            const int g = warp & 3, j = warp >> 2;
            uint32_t lb[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                lb[t] = pl.tab + (uint32_t)(t >> 1) * 65536u + (uint32_t)(t & 1) * 128u + 4u * (uint32_t)lane;
            constexpr int NACC = H2 ? 32 : 16;
            constexpr int LPC = 32 / NACC;                         // lanes holding the same value after the butterfly
            const int vidx = lane / LPC;                           // value index after the butterfly: (rhs, a) = (vidx / 16, vidx % 16)
            const int a_idx = vidx & 15;
            // register a = 4 m + i of this warp holds word `lane` of genotype column 64 j + 16 m + 4 g + i of the unit
            const int col = 64 * j + 16 * (a_idx >> 2) + 4 * g + (a_idx & 3);
            const bool writer = (lane & (LPC - 1)) == 0;
            const uint32_t tlane = tbase + ((uint32_t)(32 * g) << 16) + (uint32_t)(16 * j);
            int ts = 0; uint32_t tph = 0;
            int64_t slab = slab_beg, cb = u_beg - slab_beg * ncb;
            bool need_build = true;
            for (int64_t u = u_beg; u < u_end; ++u) {
                if (need_build) {
                    consumer_bar<TM_CW * 32>();       // everyone finished looking up the previous slab's tables
                    if (H2) lut_build_h2(pl.tab, v, v1, vbar, vbar1, sc0, sc1, n, slab, tid);
                    else lut_build<TM_CW * 32>(pl.tab, v, vbar, n, slab, tid);
                    consumer_bar<TM_CW * 32>();
                    need_build = false;
                }
                mbar_wait(pl.tmem_full(ts), tph);
                tc_fence_after();
                uint32_t w[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                      "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                    : "r"(tlane + (uint32_t)(64 * ts)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // the unit's words are in registers: hand the tensor-memory slot back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(pl.tmem_empty(ts));
                float acc[NACC];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    if (!H2) {
                        const float t0 = lds_f32(__byte_perm(w[c], lb[0], 0x7604));
                        const float t1 = lds_f32(__byte_perm(w[c], lb[1], 0x7614));
                        const float t2 = lds_f32(__byte_perm(w[c], lb[2], 0x7624));
                        const float t3 = lds_f32(__byte_perm(w[c], lb[3], 0x7634));
                        acc[c] = (t0 + t1) + (t2 + t3);
                    } else {
                        const uint32_t r0 = lds_u32(__byte_perm(w[c], lb[0], 0x7604));
                        const uint32_t r1 = lds_u32(__byte_perm(w[c], lb[1], 0x7614));
                        const uint32_t r2 = lds_u32(__byte_perm(w[c], lb[2], 0x7624));
                        const uint32_t r3 = lds_u32(__byte_perm(w[c], lb[3], 0x7634));
                        const __half2 hs = __hadd2(__hadd2(*reinterpret_cast<const __half2*>(&r0), *reinterpret_cast<const __half2*>(&r1)),
                                                   __hadd2(*reinterpret_cast<const __half2*>(&r2), *reinterpret_cast<const __half2*>(&r3)));
                        const float2 f = __half22float2(hs);
                        acc[c] = f.x;
                        acc[(H2 ? 16 : 0) + c] = f.y;
                    }
                }
                int o = 16;
#pragma unroll
                for (int h = NACC / 2; h >= 1; h >>= 1, o >>= 1) {
                    const bool upper = (lane & o) != 0;
#pragma unroll
                    for (int c = 0; c < h; ++c) {
                        const float send = upper ? acc[c] : acc[c + h];
                        const float keep = upper ? acc[c + h] : acc[c];
                        acc[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
#pragma unroll
                for (int oo = 16 / NACC; oo >= 1; oo >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], oo);
                const int64_t jj = cb * TM_UNIT_COLS + col;
                if (writer && jj < p_out) part[((int64_t)(H2 ? vidx >> 4 : 0) * n_slabs + slab) * p_out + jj] = acc[0];
                if (++ts == TM_SLOTS) { ts = 0; tph ^= 1u; }
                if (++cb == ncb) { cb = 0; ++slab; need_build = true; }
            }
        }
    }
    // every tcgen05 operation of this CTA has completed (the consumers waited for the last unit): free tensor memory
    tc_fence_before();
    __syncthreads();
    if (warp == TM_CW + 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
    }
}

template <bool H2>
static void launch_tmem(const ihtb_geno* g, const double* d_v, const double* d_v1, const double* d_vbar,
                        const float* d_scale, float* d_part, cudaStream_t s) {
    ensure_dynamic_smem(k_sweep_tmem<H2>, TM_SMEM_BYTES);
    const int64_t n_slabs = g->stride / 128;
    const int64_t units = n_slabs * ceil_div(g->p4, TM_UNIT_COLS);
    int grid = g->sm_count;
    if (units < grid) grid = (int)units;
    IHTB_CHECK(g->p < (int64_t(1) << 31) - 256, IHTB_EDIM, "more than 2^31 SNP columns on one GPU");
    IHTB_LAUNCH((k_sweep_tmem<H2>), grid, TM_THREADS, TM_SMEM_BYTES, s, g->bed.p, g->cs_s, g->p4, g->p, g->n, n_slabs,
                d_v, d_v1, d_vbar, d_scale, d_part, (uint32_t)TM_SMEM_BYTES);
}

// Opt-in (IHTB_SWEEP_TMEM=1, read at every launch so that tests can switch it): measured on B200 at n = 50k, p = 500k
// this kernel runs at 1.29 ms (0.74 of the HBM peak) against 1.135 ms (0.84) for the stage ring of sweep_lut.cu.  ncu
// (profiles/r2_sweep_tmem_ncu_full.txt) shows why the saved LDS does not pay: the tcgen05.cp reads of shared memory
// (l1tex__data_pipe_tc_wavefronts_mem_shared_op_utccp, one per chunk) take their data-RAM cycles from the LSU pipe --
// they appear one for one as l1tex__data_bank_conflicts_pipe_lsu_mem_shared -- so a chunk still costs six passes over
// the 128 B/clk shared-memory RAM (TMA write, UTCCP read, four lookups).  The bound of the table sweeps is that RAM.
bool sweep_tmem_enabled(const ihtb_geno* g) {
    const char* e = getenv("IHTB_SWEEP_TMEM");
    return g->quad && e && *e == '1';
}
void sweep_tmem_partials(const ihtb_geno* g, const double* d_v, const double* d_vbar, float* d_part, cudaStream_t s) {
    launch_tmem<false>(g, d_v, nullptr, d_vbar, nullptr, d_part, s);
}
void sweep_tmem_pair_partials(const ihtb_geno* g, const double* d_v0, const double* d_v1, const double* d_vbar,
                              const float* d_scale, float* d_part, cudaStream_t s) {
    launch_tmem<true>(g, d_v0, d_v1, d_vbar, d_scale, d_part, s);
}

}  // namespace ihtb
