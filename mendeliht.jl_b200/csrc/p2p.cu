// Fused "X[:,idx]*coef -> all-reduce" over NVLink peer memory for SNP-sharded fits.
//
// Every rank owns one cudaMalloc'ed symmetric region, mapped into all peers with CUDA IPC:
//     [ flags: PAR x nranks u64 | slots: PAR x nranks x slot_elems doubles ]
// The producer kernel (k_x_support_push, support.cu) computes this rank's partial n-vector and STORES it straight into
// slot[parity][my_rank] of EVERY rank (st.global on mapped peer pointers, no staging copy, no NCCL call); its last CTA
// then publishes flag[parity][my_rank] = seq on every rank with a system-scope release.  The consumer kernel
// (k_p2p_reduce) waits for the nranks flags in local memory and adds the nranks local slots in rank order, so all
// ranks obtain bit-identical sums.  `seq` alternates over PAR = 4 parities; a slot is only reused after the host of
// every rank has consumed two later results, which orders the reuse.
#include "comm.cuh"

namespace ihtb {

__global__ void k_p2p_push(const double* __restrict__ src, int64_t n, P2PView v, unsigned long long seq) {
    // plain push of a ready local vector (or zeros when src == NULL), same protocol as the fused producer
    for (int64_t i = 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x); i < n; i += 2 * (int64_t)gridDim.x * blockDim.x) {
        if (i + 1 < n) {
            const double2 x = src ? make_double2(src[i], src[i + 1]) : make_double2(0.0, 0.0);
            for (int r = 0; r < v.nranks; ++r) *reinterpret_cast<double2*>(v.push_slot[r] + i) = x;
        } else {
            const double x = src ? src[i] : 0.0;
            for (int r = 0; r < v.nranks; ++r) v.push_slot[r][i] = x;
        }
    }
    p2p_publish(v, seq);
}

__global__ void __launch_bounds__(256)
k_p2p_reduce(double* __restrict__ out, int64_t n, P2PView v, unsigned long long seq, unsigned long long timeout_ns) {
    if (threadIdx.x < v.nranks) {
        const volatile unsigned long long* f = v.local_flag + threadIdx.x;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (true) {
            unsigned long long cur;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f) : "memory");
            if (cur >= seq) break;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > timeout_ns) { atomicExch(v.err, 1); break; }    // a peer is gone: report, do not hang
        }
    }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < v.nranks; ++r) a += __ldcv(v.local_slot[r] + i);
        out[i] = a;
    }
}

static const int kPar = 4;

P2PView p2p_view(ihtb_comm* c) {
    P2PView v{};
    const int par = (int)(c->p2p_seq % kPar);
    v.nranks = c->nranks; v.rank = c->rank;
    for (int r = 0; r < c->nranks; ++r) {
        uint8_t* base = c->p2p_peer[r];
        double* slots = reinterpret_cast<double*>(base + 4096);
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(base);
        // where I write on rank r: slot[par][my_rank], flag[par][my_rank]
        v.push_slot[r] = slots + ((size_t)par * c->nranks + c->rank) * c->p2p_slot_elems;
        v.push_flag[r] = flags + (size_t)par * c->nranks + c->rank;
    }
    double* lslots = reinterpret_cast<double*>(c->p2p_local + 4096);
    for (int r = 0; r < c->nranks; ++r) v.local_slot[r] = lslots + ((size_t)par * c->nranks + r) * c->p2p_slot_elems;
    v.local_flag = reinterpret_cast<unsigned long long*>(c->p2p_local) + (size_t)par * c->nranks;
    v.counter = c->p2p_counter;
    v.err = c->p2p_err;
    return v;
}

static bool p2p_size_ok(size_t n);
// Every rank stores its whole vector to every peer, which beats NCCL's latency for the n of a GWAS cohort (8 GPUs:
// 21 vs 29 us at n = 50k) but not its bandwidth (67 vs 53 us at n = 500k): long vectors stay on NCCL.
bool p2p_ready(const ihtb_comm* c, size_t n) {
    return c && c->p2p_local && n <= c->p2p_slot_elems && p2p_size_ok(n);
}
static bool p2p_size_ok(size_t n) {
    static const size_t max_n = [] {
        const char* e = getenv("IHTB_P2P_MAX_N");
        return e ? (size_t)atoll(e) : (size_t)262144;
    }();
    return n <= max_n;
}

void p2p_setup(ihtb_comm* c, size_t n, cudaStream_t s, bool any_size) {
    if (!c || c->nranks <= 1 || c->nranks > P2P_MAX_RANKS || c->p2p_tried) return;
    c->p2p_tried = true;
    const char* e = getenv("IHTB_P2P");
    if (e && *e == '0') return;
    if (!any_size && !p2p_size_ok(n)) { c->p2p_tried = false; return; }     // a later, shorter fit may still set it up
    const size_t slot = (n + 63) / 64 * 64;
    const size_t bytes = 4096 + (size_t)kPar * c->nranks * slot * sizeof(double);
    uint8_t* local = nullptr;
    bool ok = cudaMalloc((void**)&local, bytes) == cudaSuccess && cudaMemset(local, 0, bytes) == cudaSuccess;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (ok) ok = cudaIpcGetMemHandle(&h, local) == cudaSuccess;
    // exchange [ok, handle] with the NCCL all-gather (9 x int64 per rank)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    int64_t mine[9];
    mine[0] = ok ? 1 : 0;
    memcpy(mine + 1, &h, 64);
    DBuf<int64_t> d_send(9), d_recv((size_t)9 * c->nranks);
    std::vector<int64_t> all((size_t)9 * c->nranks);
    IHTB_CUDA(cudaMemcpyAsync(d_send.p, mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    comm_allgather_i64(c, d_send.p, d_recv.p, 9, s);
    IHTB_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    IHTB_CUDA(cudaStreamSynchronize(s));
    bool all_ok = true;
    for (int r = 0; r < c->nranks; ++r) all_ok = all_ok && all[(size_t)9 * r] == 1;
    std::vector<uint8_t*> peer((size_t)c->nranks, nullptr);
    if (all_ok) {
        for (int r = 0; r < c->nranks && all_ok; ++r) {
            if (r == c->rank) { peer[r] = local; continue; }
            cudaIpcMemHandle_t hr;
            memcpy(&hr, &all[(size_t)9 * r + 1], 64);
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { all_ok = false; cudaGetLastError(); }
            peer[r] = static_cast<uint8_t*>(ptr);
        }
    }
    // agree on the outcome (a failed open on any rank disables the path everywhere) -- also the barrier that makes
    // sure every mapping exists before the first push
    DBuf<double> d_flag(1);
    double okv = all_ok ? 0.0 : 1.0;
    IHTB_CUDA(cudaMemcpyAsync(d_flag.p, &okv, sizeof(double), cudaMemcpyHostToDevice, s));
    comm_allreduce_sum_f64(c, d_flag.p, 1, s);
    IHTB_CUDA(cudaMemcpyAsync(&okv, d_flag.p, sizeof(double), cudaMemcpyDeviceToHost, s));
    IHTB_CUDA(cudaStreamSynchronize(s));
    if (okv != 0.0) {
        for (int r = 0; r < c->nranks; ++r)
            if (r != c->rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
        if (local) cudaFree(local);
        cudaGetLastError();
        return;
    }
    c->p2p_local = local;
    c->p2p_peer = peer;
    c->p2p_slot_elems = slot;
    IHTB_CUDA(cudaMalloc((void**)&c->p2p_counter, sizeof(unsigned)));
    IHTB_CUDA(cudaMalloc((void**)&c->p2p_err, sizeof(int)));
    IHTB_CUDA(cudaMemset(c->p2p_counter, 0, sizeof(unsigned)));
    IHTB_CUDA(cudaMemset(c->p2p_err, 0, sizeof(int)));
    c->p2p_seq = 0;
}

void p2p_teardown(ihtb_comm* c) {
    if (!c || !c->p2p_local) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->nranks; ++r)
        if (r != c->rank && c->p2p_peer[r]) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    cudaFree(c->p2p_local);
    if (c->p2p_counter) cudaFree(c->p2p_counter);
    if (c->p2p_err) cudaFree(c->p2p_err);
    c->p2p_local = nullptr;
}

// push a ready vector (src may be NULL = zeros) and reduce: the non-fused form, used when this rank owns no support column
void p2p_push(ihtb_comm* c, const double* d_src, size_t n, cudaStream_t s) {
    P2PView v = p2p_view(c);
    int grid = (int)std::min<size_t>(148, (n + 511) / 512);
    IHTB_LAUNCH(k_p2p_push, grid, 256, 0, s, d_src, (int64_t)n, v, (unsigned long long)(c->p2p_seq + 1));
}

void p2p_reduce(ihtb_comm* c, double* d_out, size_t n, cudaStream_t s) {
    P2PView v = p2p_view(c);
    int grid = (int)std::min<size_t>(148, (n + 255) / 256);
    static const unsigned long long timeout_ns = [] {
        const char* e = getenv("IHTB_P2P_TIMEOUT_S");
        double sec = e ? atof(e) : 30.0;
        return (unsigned long long)((sec > 0 ? sec : 30.0) * 1e9);
    }();
    IHTB_LAUNCH(k_p2p_reduce, grid, 256, 0, s, d_out, (int64_t)n, v, (unsigned long long)(c->p2p_seq + 1), timeout_ns);
    ++c->p2p_seq;
    ++c->n_collectives;
}

bool p2p_failed(ihtb_comm* c) {
    if (!c || !c->p2p_err) return false;
    int e = 0;
    cudaMemcpy(&e, c->p2p_err, sizeof(int), cudaMemcpyDeviceToHost);
    return e != 0;
}

}  // namespace ihtb

using namespace ihtb;

extern "C" int32_t ihtb_comm_allreduce_bench(ihtb_comm* c, int64_t n, int32_t reps, int32_t use_p2p, double* us_per_op) {
    return guard([&] {
        IHTB_CHECK(c && n > 0 && reps > 0 && us_per_op, IHTB_EINVAL, "bad argument");
        cudaStream_t s = nullptr;
        DBuf<double> src((size_t)n), dst((size_t)n);
        IHTB_CUDA(cudaMemset(src.p, 0, (size_t)n * sizeof(double)));
        p2p_setup(c, (size_t)n, s, /*any_size=*/true);
        IHTB_CHECK(!use_p2p || (c->p2p_local && (size_t)n <= c->p2p_slot_elems), IHTB_EUNSUPPORTED, "peer-memory path is not available");
        cudaEvent_t e0, e1;
        IHTB_CUDA(cudaEventCreate(&e0)); IHTB_CUDA(cudaEventCreate(&e1));
        for (int it = -10; it < reps; ++it) {
            if (it == 0) IHTB_CUDA(cudaEventRecord(e0, s));
            if (use_p2p) { p2p_push(c, src.p, (size_t)n, s); p2p_reduce(c, dst.p, (size_t)n, s); }
            else comm_allreduce_sum_f64(c, src.p, (size_t)n, s);     // in place, as in the fit
        }
        IHTB_CUDA(cudaEventRecord(e1, s));
        IHTB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        IHTB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        IHTB_CHECK(!p2p_failed(c), IHTB_ECUDA, "peer-memory all-reduce timed out waiting for another rank");
        *us_per_op = (double)ms * 1e3 / reps;
    });
}
