// Probe: how does tcgen05.cp.128x256b map a shared-memory matrix descriptor (SWIZZLE_NONE) onto TMEM lanes / columns,
// and what does tcgen05.ld.32x32b.x8 hand to each lane?  Shared memory is filled with word indices, one copy is issued
// for each (LBO, SBO) pair given on the command line, and every lane's 8 registers are printed as word indices.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu ; run: ./tmem_probe LBO SBO [LBO SBO ...]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) k_probe(uint32_t lbo, uint32_t sbo, uint32_t* out /*[128][8]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) w[i] = (uint32_t)i;      // word index pattern, 16 KB
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // make the generic-proxy stores visible to the async proxy (tcgen05.cp reads shared memory through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base_s;
    if (threadIdx.x == 0) {
        uint64_t desc = 0;
        desc |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);
        desc |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
        desc |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
        desc |= (uint64_t)1 << 46;                                  // descriptor version (sm_100)
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tbase), "l"(desc));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // everyone waits for the copy
    {
        uint32_t b = smem_u32(&bar);
        asm volatile(
            "{\n\t.reg .pred p;\n"
            "W_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@!p bra W_%=;\n\t}" ::"r"(b) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[8];
    const uint32_t taddr = tbase + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 8; ++i) out[(32 * warp + lane) * 8 + i] = r[i];
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tbase));
}

int main(int argc, char** argv) {
    uint32_t* d_out;
    cudaMalloc(&d_out, 128 * 8 * 4);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    for (int a = 1; a + 1 < argc; a += 2) {
        uint32_t lbo = (uint32_t)atoi(argv[a]), sbo = (uint32_t)atoi(argv[a + 1]);
        cudaMemset(d_out, 0xFF, 128 * 8 * 4);
        k_probe<<<1, 128, 32768>>>(lbo, sbo, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("== LBO=%u SBO=%u : %s\n", lbo, sbo, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        uint32_t h[128 * 8];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        // print the BYTE offset of the word each (lane, register) received, for a few lanes
        const int lanes[] = {0, 1, 2, 7, 8, 9, 31, 32, 33, 40, 64, 96, 127};
        for (int L : lanes) {
            printf("lane %3d:", L);
            for (int i = 0; i < 8; ++i) printf(" %6u", h[L * 8 + i] * 4u);
            printf("\n");
        }
    }
    return 0;
}
