// FAST X'v sweep for sm_100a: byte-indexed lookup tables ("four Russians") in shared memory.
//
// Why: at 6.5 TB/s a B200 SM must consume ~100 genotypes per clock; a decode + FMA per genotype is ALU-bound at
// less than half of that.  Here every packed byte (4 genotypes) costs one PRMT (address),
// one conflict-free LDS.32 (table lookup) and one FADD.
//
// Work decomposition: the matrix is cut into slabs of 512 samples (= one 128-byte chunk per column).  A CTA owns a
// contiguous range of (slab, 128-column block) units.  For its current slab it builds, from u = v - mean(v) (FP32),
//     T[t][value][w] = sum_{s<4} dosage((value >> 2s) & 3) * u[512*slab + 16*w + 4*t + s]
// for the 32 words w of a chunk and the 4 bytes t of a word: 4*256*32 floats = 128 KB.  Lane w of a consumer warp
// owns word w of a column chunk, so the 32 lookups of one LDS hit 32 distinct banks whatever the data bytes are.
// The 256 rows (value) of a table are 256 bytes apart and each pair of tables fills one 64 KB-aligned window of
// shared memory, so the lookup address is a single PRMT: byte 1 of a per-lane base register is replaced by the
// data byte.
//
// Data movement: a producer warp streams column chunks into a ring of 16 KB shared-memory stages with
// cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on mbarriers; 8 consumer warps each reduce 16 columns per
// stage with a butterfly so that lanes 0..15 end up holding one column sum each.  Slab partial sums are written as
// FP32 [slab][column]; the epilogue (sweep.cu) adds them in FP64 in slab order.  Everything is deterministic.
#include "common.cuh"
#include <stdlib.h>

namespace ihtb {

constexpr int LUT_STAGE_COLS = 128;
constexpr int LUT_STAGE_BYTES = LUT_STAGE_COLS * 128;
constexpr int LUT_MAX_STAGES = 8;
constexpr int LUT_TABLE_BYTES = 131072;
constexpr int LUT_SMEM_BYTES = 232448;   // 227 KB: everything the SM has

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
template <int NT>
__device__ __forceinline__ void consumer_bar() {   // named barrier 1 over the NT consumer threads
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// shared-memory plan (computed identically by every thread)
struct LutPlan {
    uint32_t bar_full, bar_empty;      // arrays of 8-byte mbarriers
    uint32_t tab;                      // 64 KB-aligned, 128 KB
    uint32_t lo0, hi0;                 // first stage below / above the tables
    int n_lo, n_stages;
    __device__ __forceinline__ uint32_t stage(int s) const {
        return (s < n_lo) ? lo0 + (uint32_t)s * LUT_STAGE_BYTES : hi0 + (uint32_t)(s - n_lo) * LUT_STAGE_BYTES;
    }
};
__device__ __forceinline__ LutPlan lut_plan(uint32_t base, uint32_t bytes) {
    LutPlan pl;
    uint32_t end = base + bytes;
    pl.tab = (base + 65535u) & ~65535u;
    pl.bar_full = 0; pl.bar_empty = 0; pl.n_stages = 0;
    uint32_t lo = (base + 127u) & ~127u, lo_end = pl.tab;
    uint32_t hi = pl.tab + LUT_TABLE_BYTES, hi_end = end;
    // barriers: 256 bytes at the first free spot: full[2][LUT_MAX_STAGES] then empty[LUT_MAX_STAGES]
    if (lo + 256u <= lo_end) { pl.bar_full = lo; lo += 256u; }
    else { pl.bar_full = hi; hi += 256u; }
    pl.bar_empty = pl.bar_full + 8u * 2u * LUT_MAX_STAGES;
    pl.lo0 = lo; pl.hi0 = hi;
    pl.n_lo = (lo_end > lo) ? (int)((lo_end - lo) / LUT_STAGE_BYTES) : 0;
    if (pl.n_lo > LUT_MAX_STAGES) pl.n_lo = LUT_MAX_STAGES;
    int n_hi = (hi_end > hi) ? (int)((hi_end - hi) / LUT_STAGE_BYTES) : 0;
    pl.n_stages = pl.n_lo + n_hi;
    if (pl.n_stages > LUT_MAX_STAGES) pl.n_stages = LUT_MAX_STAGES;
    return pl;
}

// build T for one slab; executed by the NT (256 or 512) consumer threads
template <int NT>
__device__ __forceinline__ void lut_build(uint32_t tab, const double* __restrict__ v, double vbar, int64_t n,
                                          int64_t slab, int tid) {
    // thread -> (part, group): group = t*32 + w (128 groups); part selects a range of the top sample's code v3
    constexpr int PARTS = NT / 128;          // 2 or 4
    constexpr int V3_PER = 4 / PARTS;        // 2 or 1
    const int group = tid & 127, part = tid >> 7;
    const int t = group >> 5, w = group & 31;
    float u[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int64_t i = slab * 512 + 16 * w + 4 * t + s;
        u[s] = (i < n) ? __double2float_rn(__dsub_rn(v[i], vbar)) : 0.0f;
    }
    // dosage table of one sample: codes 00, 01 (missing -> 0), 10, 11
    auto f = [&](int s, int code) -> float { return code == 2 ? u[s] : (code == 3 ? u[s] + u[s] : 0.0f); };
    // address of row `value`: window(t>>1) + value*256 + (t&1)*128 + 4*w
    const uint32_t rowbase = tab + (uint32_t)(t >> 1) * 65536u + (uint32_t)(t & 1) * 128u + 4u * (uint32_t)w;
#pragma unroll
    for (int c3 = 0; c3 < V3_PER; ++c3) {
        const int v3 = part * V3_PER + c3;                 // runtime, but only used arithmetically
        const float a3 = (v3 == 2) ? u[3] : ((v3 == 3) ? u[3] + u[3] : 0.0f);
        const uint32_t base3 = rowbase + (uint32_t)v3 * (64u * 256u);
#pragma unroll
        for (int v2 = 0; v2 < 4; ++v2) {
            const float a2 = a3 + f(2, v2);
#pragma unroll
            for (int v1 = 0; v1 < 4; ++v1) {
                const float a1 = a2 + f(1, v1);
#pragma unroll
                for (int v0 = 0; v0 < 4; ++v0)
                    sts_f32(base3 + (uint32_t)(v2 << 4 | v1 << 2 | v0) * 256u, a1 + f(0, v0));
            }
        }
    }
}

// Units: u = slab * n_cblocks + cblock, CTA b handles [u_begin(b), u_begin(b+1)), slab by slab.
// CW consumer warps in G groups (+ 1 producer warp).  Group g consumes the CTA's units with (local index % G) == g, and
// each of its CW/G warps reduces CPW = 128*G/CW columns of that unit, so per-unit bookkeeping is amortised over 16
// columns per warp while 16 warps keep the shared-memory pipe busy.
// The groups share one ring of S stages (unit i uses stage i % S) and one "empty" mbarrier per stage, but every
// group has its OWN "full" mbarrier per stage: the producer signals full[i % G][i % S].  A warp must observe every
// phase of a barrier it waits on -- a parity wait issued a whole phase early is satisfied by the preceding phase
// (the mbarrier ABA hazard) and the warp would read a slot before it is refilled.  With per-group full barriers a
// group sees exactly the fills meant for it, in order, while the groups still progress independently.
template <int CW, int G>
__global__ void __launch_bounds__((CW + 1) * 32, 1)
k_sweep_lut(const uint8_t* __restrict__ bed, int64_t cs_j, int64_t cs_s, int64_t p, int64_t n, int64_t n_slabs,
            const double* __restrict__ v, const double* __restrict__ vbar_p, float* __restrict__ part,
            uint32_t dyn_bytes) {
    const double vbar = *vbar_p;
    constexpr int WPG = CW / G;                      // warps per group
    constexpr int CPW = LUT_STAGE_COLS / WPG;        // columns per warp per unit
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const LutPlan pl = lut_plan(smem_u32(smem_raw), dyn_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = pl.n_stages;
    // (group, stage) pairs repeat with period lcm(S, G) in the unit index; G is 1 or 2
    const int period = (S % G == 0) ? S : S * G;

    const int64_t n_cblocks = (p + LUT_STAGE_COLS - 1) / LUT_STAGE_COLS;
    const int64_t units = n_slabs * n_cblocks;
    const int64_t u_beg = units * (int64_t)blockIdx.x / gridDim.x;
    const int64_t u_end = units * (int64_t)(blockIdx.x + 1) / gridDim.x;
    if (u_beg >= u_end) return;
    const int64_t slab_beg = u_beg / n_cblocks, slab_end = (u_end - 1) / n_cblocks;   // inclusive
    const int ncb = (int)n_cblocks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            for (int g = 0; g < G; ++g) mbar_init(pl.bar_full + 8u * (g * LUT_MAX_STAGES + s), 1);
            mbar_init(pl.bar_empty + 8u * s, WPG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // ===== producer warp: stream column chunks with bulk async copies =====
        int st = 0; uint32_t ph = 0;                 // stage of the next unit, parity of its "empty" phase
        int grp_of = 0;                              // group that owns the next unit (local unit index % G)
        for (int64_t slab = slab_beg; slab <= slab_end; ++slab) {
            const int cb0 = (slab == slab_beg) ? (int)(u_beg - slab * n_cblocks) : 0;
            const int cb1 = (slab == slab_end) ? (int)(u_end - slab * n_cblocks) : ncb;    // exclusive
            for (int cb = cb0; cb < cb1; ++cb) {
                const int64_t j0 = (int64_t)cb * LUT_STAGE_COLS;
                const int ncols = (int)((p - j0 < LUT_STAGE_COLS) ? (p - j0) : LUT_STAGE_COLS);
                const uint32_t fullbar = pl.bar_full + 8u * (uint32_t)(grp_of * LUT_MAX_STAGES + st);
                mbar_wait(pl.bar_empty + 8u * st, ph ^ 1u);
                const uint8_t* src = bed + j0 * cs_j + slab * cs_s;
                if (cs_j == 128) {               // slab-major tiled layout: one contiguous copy
                    if (lane == 0) {
                        mbar_expect_tx(fullbar, (uint32_t)ncols * 128u);
                        bulk_g2s(pl.stage(st), src, (uint32_t)ncols * 128u, fullbar);
                    }
                } else {                          // column-major layout: one 128-byte copy per column
                    if (lane == 0) mbar_expect_tx(fullbar, (uint32_t)ncols * 128u);
                    __syncwarp();
                    for (int c = lane; c < ncols; c += 32)
                        bulk_g2s(pl.stage(st) + 128u * c, src + (int64_t)c * cs_j, 128u, fullbar);
                }
                if (++st == S) { st = 0; ph ^= 1u; }
                if (++grp_of == G) grp_of = 0;
            }
        }
    } else {
        // ===== consumer warps =====
        const int tid = threadIdx.x;      // 0 .. CW*32-1
        const int grp = warp / WPG, wg = warp % WPG;
        // per-lane base registers for the 4 bytes of a word: byte 1 gets replaced by the data byte
        uint32_t lb[4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
            lb[t] = pl.tab + (uint32_t)(t >> 1) * 65536u + (uint32_t)(t & 1) * 128u + 4u * (uint32_t)lane;
        constexpr int LPC = 32 / CPW;                 // lanes holding the same column after the butterfly
        const int col = wg * CPW + lane / LPC;        // column (within a unit) this lane stores
        const bool writer = (lane & (LPC - 1)) == 0;
        const uint32_t lane_off = (uint32_t)(wg * CPW) * 128u + 4u * (uint32_t)lane;
        // this group consumes local unit indices i = grp, grp+G, ...: stage i % S, full-barrier parity (i / period) & 1
        int st = grp % S, ip = grp % period; uint32_t ph = (uint32_t)((grp / period) & 1);
        const uint32_t my_full = pl.bar_full + 8u * (uint32_t)(grp * LUT_MAX_STAGES);
        int i_next = grp;                             // local index of this group's next unit
        int i_base = 0;                               // local index of the first unit of the current slab
        for (int64_t slab = slab_beg; slab <= slab_end; ++slab) {
            const int cb0 = (slab == slab_beg) ? (int)(u_beg - slab * n_cblocks) : 0;
            const int cb1 = (slab == slab_end) ? (int)(u_end - slab * n_cblocks) : ncb;
            consumer_bar<CW * 32>();      // everyone finished looking up the previous slab's tables
            lut_build<CW * 32>(pl.tab, v, vbar, n, slab, tid);
            consumer_bar<CW * 32>();
            float* __restrict__ outp = part + slab * p + col;
            const int i_end = i_base + (cb1 - cb0);
            for (; i_next < i_end; i_next += G) {
                const int cb = cb0 + (i_next - i_base);
                mbar_wait(my_full + 8u * st, ph);
                const uint32_t colbase = pl.stage(st) + lane_off;
                float acc[CPW];
#pragma unroll
                for (int c = 0; c < CPW; ++c) {
                    const uint32_t w = lds_u32(colbase + 128u * c);
                    const float t0 = lds_f32(__byte_perm(w, lb[0], 0x7604));
                    const float t1 = lds_f32(__byte_perm(w, lb[1], 0x7614));
                    const float t2 = lds_f32(__byte_perm(w, lb[2], 0x7624));
                    const float t3 = lds_f32(__byte_perm(w, lb[3], 0x7634));
                    acc[c] = (t0 + t1) + (t2 + t3);
                }
                // the stage's bytes are now in registers: hand the slot back to the producer
                __syncwarp();
                if (lane == 0) mbar_arrive(pl.bar_empty + 8u * st);
                // butterfly: CPW column sums per lane -> one per lane; the high lane bits select the column
                int o = 16;
#pragma unroll
                for (int h = CPW / 2; h >= 1; h >>= 1, o >>= 1) {
                    const bool upper = (lane & o) != 0;
#pragma unroll
                    for (int c = 0; c < h; ++c) {
                        const float send = upper ? acc[c] : acc[c + h];
                        const float keep = upper ? acc[c + h] : acc[c];
                        acc[c] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
#pragma unroll
                for (int oo = 16 / CPW; oo >= 1; oo >>= 1) acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], oo);
                const int jj = cb * LUT_STAGE_COLS + col;
                if (writer && jj < (int)p) outp[cb * LUT_STAGE_COLS] = acc[0];
                st += G; if (st >= S) st -= S;
                ip += G; if (ip >= period) { ip -= period; ph ^= 1u; }
            }
            i_base = i_end;
        }
    }
}

void sweep_fast_partials(const ihtb_geno* g, const double* d_v, const double* d_vbar, float* d_part, int64_t* n_slabs_out,
                         cudaStream_t s);

int64_t sweep_fast_num_slabs(const ihtb_geno* g) { return g->stride / 128; }

void sweep_fast_kernel_only(const ihtb_geno* g, const double* d_v, const double* d_vbar, float* d_part, cudaStream_t s) {
    int64_t ns = 0;
    sweep_fast_partials(g, d_v, d_vbar, d_part, &ns, s);
}

void sweep_fast_partials(const ihtb_geno* g, const double* d_v, const double* d_vbar, float* d_part, int64_t* n_slabs_out,
                         cudaStream_t s) {
    static int cw = 0;
    if (!cw) {
        IHTB_CUDA(cudaFuncSetAttribute(k_sweep_lut<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_SMEM_BYTES));
        IHTB_CUDA(cudaFuncSetAttribute(k_sweep_lut<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_SMEM_BYTES));
        IHTB_CUDA(cudaFuncSetAttribute(k_sweep_lut<16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, LUT_SMEM_BYTES));
        const char* e = getenv("IHTB_LUT_WARPS");       // tuning knob: 8 = <8,1>, 16 = <16,1>, default <16,2>
        cw = e ? atoi(e) : 162;
        if (cw != 8 && cw != 16) cw = 162;
    }
    const int64_t n_slabs = sweep_fast_num_slabs(g);
    *n_slabs_out = n_slabs;
    const int64_t n_cblocks = ceil_div(g->p, LUT_STAGE_COLS);
    int64_t units = n_slabs * n_cblocks;
    int grid = g->sm_count;
    if (units < grid) grid = (int)units;
    IHTB_CHECK(g->p < (int64_t(1) << 31) - 256, IHTB_EDIM, "more than 2^31 SNP columns on one GPU");
    if (cw == 8) {
        IHTB_LAUNCH((k_sweep_lut<8, 1>), grid, 9 * 32, LUT_SMEM_BYTES, s, g->bed.p, g->cs_j, g->cs_s, g->p, g->n,
                    n_slabs, d_v, d_vbar, d_part, (uint32_t)LUT_SMEM_BYTES);
    } else if (cw == 16) {
        IHTB_LAUNCH((k_sweep_lut<16, 1>), grid, 17 * 32, LUT_SMEM_BYTES, s, g->bed.p, g->cs_j, g->cs_s, g->p, g->n,
                    n_slabs, d_v, d_vbar, d_part, (uint32_t)LUT_SMEM_BYTES);
    } else {
        IHTB_LAUNCH((k_sweep_lut<16, 2>), grid, 17 * 32, LUT_SMEM_BYTES, s, g->bed.p, g->cs_j, g->cs_s, g->p, g->n,
                    n_slabs, d_v, d_vbar, d_part, (uint32_t)LUT_SMEM_BYTES);
    }
}

}  // namespace ihtb
