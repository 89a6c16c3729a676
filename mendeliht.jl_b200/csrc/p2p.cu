// Collectives of SNP-sharded fits over NVLink peer memory -- no NCCL call inside the IHT loop.
//
// Every rank owns one cudaMalloc'ed symmetric region with the same layout, mapped into all peers (CUDA IPC between
// processes, cudaDeviceEnablePeerAccess between the devices of one process):
//     [ flags : KINDS x PAR x MAXR u64 | push-all slots : PAR x R x pa_cap f64 | partial : PAR x red_cap f64 |
//       result : PAR x red_cap f64 | gather : PAR x R x gat_cap i64 ]
// Operations of one kind carry a sequence number; `seq % PAR` selects the parity (slot set), flags hold the last
// sequence number published and waiters test `flag >= seq`, so nothing is ever reset.
//
//  * push-all all-reduce (short vectors, n <= pa_cap = 262144): the producer kernel (k_x_support<1, PUSH>, support.cu,
//    or k_p2p_push) STORES its partial n-vector straight into slot[parity][my_rank] of EVERY rank and its last CTA
//    publishes the flag; k_p2p_reduce waits for the R local flags and adds the R local slots in rank order.  One
//    flag hop, 2 kernels; every rank moves R x n x 8 bytes (8 GPUs, n = 50k: 21 us vs 29 us for ncclAllReduce).
//  * two-phase all-reduce (long vectors; n = 500k moves 32 MB per rank with push-all and loses to NCCL, round 1):
//    the producer writes its partial into its OWN partial area.  k_p2p_reduce_scatter announces it, waits for the
//    peers' announcements, then reduces slice `rank` of the vector -- loads of the R partials over NVLink, summed in
//    rank order -- stores the reduced slice into the result area of every rank, publishes it, waits for the R slices
//    and copies the result out: one kernel (k_p2p_allreduce2; its grid is co-resident).  Every rank moves
//    2 (R-1)/R x n x 8 bytes instead of R x n x 8.
//  * all-gather (top-k candidate blocks): k_p2p_gpush stores this rank's block into gather[parity][my_rank] of every
//    rank, k_p2p_gwait waits for the R blocks.
// All sums are taken in rank order by exactly one rank per element, so every rank sees bit-identical results.
// Slot reuse needs no host synchronisation: an operation of sequence s+1 on any rank completes only after every peer
// finished reading what s published (each rank's kernels are stream-ordered), and PAR = 4 leaves margin.
// A %globaltimer timeout (30 s) turns a dead peer into an error instead of a hang.
#include "comm.cuh"

namespace ihtb {

// ---- in-process rank group --------------------------------------------------------------------------------------
void LocalGroup::barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) throw Error(IHTB_ECUDA, "another device of the multi-device handle failed");
    const unsigned long long gen = generation;
    if (++waiting == nranks) {
        waiting = 0;
        ++generation;
        cv.notify_all();
    } else {
        cv.wait(lk, [&] { return generation != gen || failed; });
        if (failed && generation == gen) throw Error(IHTB_ECUDA, "another device of the multi-device handle failed");
    }
}
void LocalGroup::fail() {
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
}

// ---- device side ------------------------------------------------------------------------------------------------
struct SymView {                       // everything a two-phase / gather kernel needs
    uint8_t* base[P2P_MAX_RANKS];      // mapped base of every rank's region (base[rank] = local)
    int nranks, rank;
    unsigned* counter;
    int* err;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long* sym_flag(uint8_t* base, int kind, int par, int r) {
    return reinterpret_cast<unsigned long long*>(base) + ((size_t)kind * P2P_PAR + par) * P2P_MAX_RANKS + r;
}

// threads 0..nranks-1 of the CTA poll flag[kind][par][thread] in LOCAL memory until it reaches seq
__device__ __forceinline__ void sym_wait_all(const SymView& v, int kind, int par, unsigned long long seq) {
    if (threadIdx.x < v.nranks) {
        const unsigned long long* f = sym_flag(v.base[v.rank], kind, par, threadIdx.x);
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (true) {
            unsigned long long cur;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(cur) : "l"(f) : "memory");
            if (cur >= seq) break;
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > v.timeout_ns) { atomicExch(v.err, 1); break; }    // a peer is gone: report, do not hang
        }
    }
    __syncthreads();
}

// last CTA of the grid publishes flag[kind][par][my_rank] = seq on every rank (all CTAs fenced their stores first)
__device__ __forceinline__ void sym_publish(const SymView& v, int kind, int par, unsigned long long seq) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(v.counter, 1u);
        if (done == gridDim.x - 1) {
            *v.counter = 0u;
            __threadfence_system();
            for (int r = 0; r < v.nranks; ++r)
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(sym_flag(v.base[r], kind, par, v.rank)), "l"(seq) : "memory");
        }
    }
}

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

// counts (histograms of the distributed radix select, topk.cu): pushed as doubles -- sums of small integers are exact
__global__ void k_p2p_push_i32(const int* __restrict__ src, int64_t n, P2PView v, unsigned long long seq) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = (double)src[i];
        for (int r = 0; r < v.nranks; ++r) v.push_slot[r][i] = x;
    }
    p2p_publish(v, seq);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_p2p_reduce(T* __restrict__ out, int64_t n, P2PView v, unsigned long long seq, unsigned long long timeout_ns) {
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
            if (t1 - t0 > timeout_ns) { atomicExch(v.err, 1); break; }
        }
    }
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < v.nranks; ++r) a += __ldcv(v.local_slot[r] + i);
        out[i] = (T)a;
    }
}

// two-phase all-reduce in ONE kernel (the grid is co-resident, so CTAs may wait on flags that other CTAs of the same
// grid help to publish): announce my partial, wait for everyone's, reduce slice `rank` (R independent peer loads in
// flight per lane), store it into every rank's result area, publish; then wait for the R slices and copy the result out.
constexpr int P2P_AR_THREADS = 512;
__global__ void __launch_bounds__(P2P_AR_THREADS)
k_p2p_allreduce2(SymView v, size_t off_partial, size_t off_result, int64_t count, double* __restrict__ out, int par,
                 unsigned long long seq) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // the producer kernel ran earlier on this stream: its stores are complete; make them visible system-wide
        __threadfence_system();
        for (int r = 0; r < v.nranks; ++r)
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(sym_flag(v.base[r], 0, par, v.rank)), "l"(seq) : "memory");
    }
    sym_wait_all(v, 0, par, seq);
    // slices are multiples of 2 elements so that every access is a 16-byte lane (the areas hold an even element count)
    const int64_t pairs = (count + 1) / 2;
    const int64_t lo = 2 * (pairs * v.rank / v.nranks), hi = min((int64_t)(2 * (pairs * (v.rank + 1) / v.nranks)), count);
    const int64_t stride = 2 * (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x); i < hi; i += 2 * stride) {
        // two elements per lane and trip: 2 R loads in flight before the first add
        const int64_t i2 = i + stride;
        const bool two = i2 < hi;
        double2 x[P2P_MAX_RANKS], y[P2P_MAX_RANKS];
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (r < v.nranks) {
                const double* src = reinterpret_cast<const double*>(v.base[r] + off_partial);
                x[r] = __ldcv(reinterpret_cast<const double2*>(src + i));
                if (two) y[r] = __ldcv(reinterpret_cast<const double2*>(src + i2));
            }
        double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (r < v.nranks) { a.x += x[r].x; a.y += x[r].y; if (two) { b.x += y[r].x; b.y += y[r].y; } }
#pragma unroll
        for (int r = 0; r < P2P_MAX_RANKS; ++r)
            if (r < v.nranks) {
                double* dst = reinterpret_cast<double*>(v.base[r] + off_result);
                *reinterpret_cast<double2*>(dst + i) = a;
                if (two) *reinterpret_cast<double2*>(dst + i2) = b;
            }
    }
    sym_publish(v, 1, par, seq);
    sym_wait_all(v, 1, par, seq);
    const double* res = reinterpret_cast<const double*>(v.base[v.rank] + off_result);
    for (int64_t i = 2 * (blockIdx.x * (int64_t)blockDim.x + threadIdx.x); i < count; i += stride) {
        if (i + 1 < count) {
            const double2 x = __ldcv(reinterpret_cast<const double2*>(res + i));
            out[i] = x.x; out[i + 1] = x.y;
        } else {
            out[i] = __ldcv(res + i);
        }
    }
}

__global__ void __launch_bounds__(256)
k_p2p_gpush(SymView v, size_t off_block /* gather[par][my_rank] */, const int64_t* __restrict__ src, int64_t count, int par,
            unsigned long long seq) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t x = src[i];
        for (int r = 0; r < v.nranks; ++r) reinterpret_cast<int64_t*>(v.base[r] + off_block)[i] = x;
    }
    sym_publish(v, 2, par, seq);
}

__global__ void __launch_bounds__(256)
k_p2p_gwait(SymView v, size_t off_par /* gather[par] */, size_t gat_cap, int64_t count, int64_t* __restrict__ out, int par,
            unsigned long long seq) {
    sym_wait_all(v, 2, par, seq);
    const int64_t total = count * v.nranks;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / count, e = i - r * count;
        out[i] = __ldcv(reinterpret_cast<const int64_t*>(v.base[v.rank] + off_par) + r * gat_cap + e);
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
static unsigned long long timeout_ns() {
    static const unsigned long long t = [] {
        const char* e = getenv("IHTB_P2P_TIMEOUT_S");
        double sec = e ? atof(e) : 30.0;
        return (unsigned long long)((sec > 0 ? sec : 30.0) * 1e9);
    }();
    return t;
}
static size_t pushall_max() {
    static const size_t m = [] {
        const char* e = getenv("IHTB_P2P_MAX_N");      // vectors up to this length take the one-hop push-all path
        return e ? (size_t)atoll(e) : (size_t)262144;
    }();
    return m;
}

static SymView sym_view(ihtb_comm* c) {
    SymView v{};
    for (int r = 0; r < c->nranks; ++r) v.base[r] = c->sym_peer[r];
    v.nranks = c->nranks; v.rank = c->rank;
    v.counter = c->p2p_counter; v.err = c->p2p_err;
    v.timeout_ns = timeout_ns();
    return v;
}

P2PView p2p_view(ihtb_comm* c) {
    P2PView v{};
    const int par = (int)((c->seq[0] + 1) % P2P_PAR);
    v.nranks = c->nranks; v.rank = c->rank;
    for (int r = 0; r < c->nranks; ++r) {
        uint8_t* base = c->sym_peer[r];
        double* slots = reinterpret_cast<double*>(base + c->off_pa);
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(base);
        v.push_slot[r] = slots + ((size_t)par * c->nranks + c->rank) * c->pa_cap;
        v.push_flag[r] = flags + ((size_t)0 * P2P_PAR + par) * P2P_MAX_RANKS + c->rank;
    }
    double* lslots = reinterpret_cast<double*>(c->sym_local + c->off_pa);
    for (int r = 0; r < c->nranks; ++r) v.local_slot[r] = lslots + ((size_t)par * c->nranks + r) * c->pa_cap;
    v.local_flag = reinterpret_cast<unsigned long long*>(c->sym_local) + ((size_t)0 * P2P_PAR + par) * P2P_MAX_RANKS;
    v.counter = c->p2p_counter;
    v.err = c->p2p_err;
    return v;
}

bool p2p_mapped(const ihtb_comm* c) { return c && c->sym_local; }
bool p2p_pushall_ok(const ihtb_comm* c, size_t count) { return p2p_mapped(c) && count <= c->pa_cap; }

static size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

void p2p_teardown(ihtb_comm* c) {
    if (!c || !c->sym_local) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->ipc_mapped)
        for (int r = 0; r < c->nranks; ++r)
            if (r != c->rank && c->sym_peer[r]) cudaIpcCloseMemHandle(c->sym_peer[r]);
    cudaFree(c->sym_local);
    if (c->p2p_counter) cudaFree(c->p2p_counter);
    if (c->p2p_err) cudaFree(c->p2p_err);
    c->sym_local = nullptr; c->p2p_counter = nullptr; c->p2p_err = nullptr;
    c->sym_peer.clear();
    c->pa_cap = c->red_cap = c->gat_cap = 0;
}

void p2p_setup(ihtb_comm* c, size_t red_elems, size_t gat_elems, cudaStream_t s) {
    if (!c || c->nranks <= 1 || c->nranks > P2P_MAX_RANKS) return;
    const char* e = getenv("IHTB_P2P");
    if (!c->local && e && *e == '0') return;                       // multi-process: IHTB_P2P=0 forces NCCL
    red_elems = round_up(std::max<size_t>(red_elems, 64), 64);
    gat_elems = round_up(std::max<size_t>(gat_elems, 64), 64);
    if (c->sym_local && c->red_cap >= red_elems && c->gat_cap >= gat_elems) return;     // same decision on every rank
    if (!c->local && c->p2p_tried && !c->sym_local) return;        // mapping failed before: stay on NCCL
    // re-size: everyone drains the old region first (a peer may still be reading it)
    if (c->sym_local) {
        IHTB_CUDA(cudaStreamSynchronize(s));
        if (c->local) c->local->barrier();
        else { DBuf<double> z(1); IHTB_CUDA(cudaMemsetAsync(z.p, 0, 8, s)); nccl_allreduce_sum_f64(c, z.p, 1, s); IHTB_CUDA(cudaStreamSynchronize(s)); }
        p2p_teardown(c);
        if (c->local) c->local->barrier();
    }
    c->p2p_tried = true;
    const size_t pa = std::min(red_elems, round_up(pushall_max(), 64));
    const size_t off_pa = 4096;
    const size_t off_partial = off_pa + (size_t)P2P_PAR * c->nranks * pa * sizeof(double);
    const size_t off_result = off_partial + (size_t)P2P_PAR * red_elems * sizeof(double);
    const size_t off_gather = off_result + (size_t)P2P_PAR * red_elems * sizeof(double);
    const size_t bytes = off_gather + (size_t)P2P_PAR * c->nranks * gat_elems * sizeof(int64_t);
    static_assert(P2P_KINDS * P2P_PAR * P2P_MAX_RANKS * 8 <= 4096, "flag block");
    uint8_t* local = nullptr;
    bool ok = cudaMalloc((void**)&local, bytes) == cudaSuccess && cudaMemset(local, 0, bytes) == cudaSuccess;
    if (!ok) cudaGetLastError();
    std::vector<uint8_t*> peer((size_t)c->nranks, nullptr);
    bool all_ok = ok;
    if (c->local) {
        // ---- devices of one process: enable peer access and trade raw pointers over the group's board ----
        LocalGroup& g = *c->local;
        for (int r = 0; r < c->nranks && all_ok; ++r) {
            if (r == c->rank) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->device, g.devices[r]) != cudaSuccess || !can) { all_ok = false; break; }
            cudaError_t pe = cudaDeviceEnablePeerAccess(g.devices[r], 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) all_ok = false;
            cudaGetLastError();
        }
        g.board[c->rank] = all_ok ? local : nullptr;
        g.barrier();
        for (int r = 0; r < c->nranks; ++r) { peer[r] = static_cast<uint8_t*>(g.board[r]); all_ok = all_ok && peer[r]; }
        g.barrier();                                               // everyone has read the board
        if (!all_ok) {
            if (local) cudaFree(local);
            throw Error(IHTB_ECUDA, "peer access between the devices of a multi-device handle is unavailable");
        }
        c->ipc_mapped = false;
    } else {
        // ---- one process per GPU: exchange [ok, IPC handle] with the NCCL all-gather (9 x int64 per rank) ----
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof(h));
        if (ok) ok = cudaIpcGetMemHandle(&h, local) == cudaSuccess;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
        int64_t mine[9];
        mine[0] = ok ? 1 : 0;
        memcpy(mine + 1, &h, 64);
        DBuf<int64_t> d_send(9), d_recv((size_t)9 * c->nranks);
        std::vector<int64_t> all((size_t)9 * c->nranks);
        IHTB_CUDA(cudaMemcpyAsync(d_send.p, mine, sizeof(mine), cudaMemcpyHostToDevice, s));
        nccl_allgather_i64(c, d_send.p, d_recv.p, 9, s);
        IHTB_CUDA(cudaMemcpyAsync(all.data(), d_recv.p, all.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        IHTB_CUDA(cudaStreamSynchronize(s));
        all_ok = true;
        for (int r = 0; r < c->nranks; ++r) all_ok = all_ok && all[(size_t)9 * r] == 1;
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
        nccl_allreduce_sum_f64(c, d_flag.p, 1, s);
        IHTB_CUDA(cudaMemcpyAsync(&okv, d_flag.p, sizeof(double), cudaMemcpyDeviceToHost, s));
        IHTB_CUDA(cudaStreamSynchronize(s));
        if (okv != 0.0) {
            for (int r = 0; r < c->nranks; ++r)
                if (r != c->rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
            if (local) cudaFree(local);
            cudaGetLastError();
            return;                                                  // stay on NCCL
        }
        c->ipc_mapped = true;
    }
    c->sym_local = local;
    c->sym_peer = peer;
    c->pa_cap = pa; c->red_cap = red_elems; c->gat_cap = gat_elems;
    c->off_pa = off_pa; c->off_partial = off_partial; c->off_result = off_result; c->off_gather = off_gather;
    c->sym_bytes = bytes;
    IHTB_CUDA(cudaMalloc((void**)&c->p2p_counter, sizeof(unsigned)));
    IHTB_CUDA(cudaMalloc((void**)&c->p2p_err, sizeof(int)));
    IHTB_CUDA(cudaMemset(c->p2p_counter, 0, sizeof(unsigned)));
    IHTB_CUDA(cudaMemset(c->p2p_err, 0, sizeof(int)));
    for (int k = 0; k < P2P_KINDS; ++k) c->seq[k] = 0;
}

// push a ready vector (src may be NULL = zeros) and reduce: the non-fused form, used when this rank owns no support column
void p2p_push(ihtb_comm* c, const double* d_src, size_t n, cudaStream_t s) {
    P2PView v = p2p_view(c);
    int grid = (int)std::min<size_t>(148, (n + 511) / 512);
    IHTB_LAUNCH(k_p2p_push, grid, 256, 0, s, d_src, (int64_t)n, v, (unsigned long long)(c->seq[0] + 1));
}

void p2p_reduce(ihtb_comm* c, double* d_out, size_t n, cudaStream_t s) {
    P2PView v = p2p_view(c);
    int grid = (int)std::min<size_t>(148, (n + 255) / 256);
    IHTB_LAUNCH(k_p2p_reduce<double>, grid, 256, 0, s, d_out, (int64_t)n, v, (unsigned long long)(c->seq[0] + 1), timeout_ns());
    ++c->seq[0];
    ++c->n_collectives;
}

static void p2p_allreduce_i32(ihtb_comm* c, int* d_buf, size_t n, cudaStream_t s) {
    P2PView v = p2p_view(c);
    int grid = (int)std::min<size_t>(148, (n + 255) / 256);
    IHTB_LAUNCH(k_p2p_push_i32, grid, 256, 0, s, d_buf, (int64_t)n, v, (unsigned long long)(c->seq[0] + 1));
    IHTB_LAUNCH(k_p2p_reduce<int>, grid, 256, 0, s, d_buf, (int64_t)n, v, (unsigned long long)(c->seq[0] + 1), timeout_ns());
    ++c->seq[0];
    ++c->n_collectives;
}

// Sequence numbers of kinds 0 and 1 advance together for a two-phase operation; a push-all operation advances kind 0
// only.  Both use flag kind 0 with the same counter, so the parity of a partial area never collides with a push slot.
double* p2p_partial_ptr(ihtb_comm* c) {
    const int par = (int)((c->seq[0] + 1) % P2P_PAR);
    return reinterpret_cast<double*>(c->sym_local + c->off_partial) + (size_t)par * c->red_cap;
}

void p2p_allreduce_2phase(ihtb_comm* c, size_t count, double* d_out, cudaStream_t s) {
    IHTB_CHECK(count <= c->red_cap, IHTB_EINVAL, "two-phase all-reduce larger than the mapped area");
    const unsigned long long seq = c->seq[0] + 1;
    const int par = (int)(seq % P2P_PAR);
    SymView v = sym_view(c);
    // every CTA waits on flags inside the kernel, so the grid must be co-resident: at most one CTA per SM here
    static int sm_count[64] = {};
    if (!sm_count[c->device & 63])
        IHTB_CUDA(cudaDeviceGetAttribute(&sm_count[c->device & 63], cudaDevAttrMultiProcessorCount, c->device));
    const size_t slice_pairs = ((count + 1) / 2 + c->nranks - 1) / c->nranks;
    int grid = (int)std::min<size_t>((size_t)sm_count[c->device & 63],
                                     std::max<size_t>(1, (slice_pairs + 2 * P2P_AR_THREADS - 1) / (2 * P2P_AR_THREADS)));
    IHTB_LAUNCH(k_p2p_allreduce2, grid, P2P_AR_THREADS, 0, s, v,
                c->off_partial + (size_t)par * c->red_cap * sizeof(double),
                c->off_result + (size_t)par * c->red_cap * sizeof(double), (int64_t)count, d_out, par, seq);
    ++c->seq[0];
    c->seq[1] = c->seq[0];
    ++c->n_collectives;
}

void p2p_allgather(ihtb_comm* c, const int64_t* d_send, size_t count, int64_t* d_recv, cudaStream_t s) {
    IHTB_CHECK(count <= c->gat_cap, IHTB_EINVAL, "all-gather block larger than the mapped area");
    const unsigned long long seq = c->seq[2] + 1;
    const int par = (int)(seq % P2P_PAR);
    SymView v = sym_view(c);
    const size_t off_par = c->off_gather + (size_t)par * c->nranks * c->gat_cap * sizeof(int64_t);
    int grid = (int)std::min<size_t>(32, std::max<size_t>(1, (count + 255) / 256));
    IHTB_LAUNCH(k_p2p_gpush, grid, 256, 0, s, v, off_par + (size_t)c->rank * c->gat_cap * sizeof(int64_t), d_send,
                (int64_t)count, par, seq);
    grid = (int)std::min<size_t>(32, std::max<size_t>(1, (count * c->nranks + 255) / 256));
    IHTB_LAUNCH(k_p2p_gwait, grid, 256, 0, s, v, off_par, c->gat_cap, (int64_t)count, d_recv, par, seq);
    ++c->seq[2];
    ++c->n_collectives;
}

bool p2p_failed(ihtb_comm* c) {
    if (!c || !c->p2p_err) return false;
    int e = 0;
    cudaMemcpy(&e, c->p2p_err, sizeof(int), cudaMemcpyDeviceToHost);
    return e != 0;
}

// ---- generic collectives: any size over peer memory (chunked), NCCL when nothing is mapped -------------------------
void comm_allreduce_sum_f64(ihtb_comm* c, double* d_buf, size_t count, cudaStream_t s) {
    if (!c || c->nranks == 1 || count == 0) return;
    if (!p2p_mapped(c)) { nccl_allreduce_sum_f64(c, d_buf, count, s); return; }
    for (size_t o = 0; o < count; o += c->red_cap) {
        const size_t m = std::min(c->red_cap, count - o);
        if (m <= c->pa_cap) {
            p2p_push(c, d_buf + o, m, s);
            p2p_reduce(c, d_buf + o, m, s);
        } else {
            IHTB_CUDA(cudaMemcpyAsync(p2p_partial_ptr(c), d_buf + o, m * sizeof(double), cudaMemcpyDeviceToDevice, s));
            p2p_allreduce_2phase(c, m, d_buf + o, s);
        }
    }
}

void comm_allreduce_sum_i32(ihtb_comm* c, int* d_buf, size_t count, cudaStream_t s) {
    if (!c || c->nranks == 1 || count == 0) return;
    if (!p2p_mapped(c)) { nccl_allreduce_sum_i32(c, d_buf, count, s); return; }
    for (size_t o = 0; o < count; o += c->pa_cap) p2p_allreduce_i32(c, d_buf + o, std::min(c->pa_cap, count - o), s);
}

void comm_allgather_i64(ihtb_comm* c, const int64_t* d_send, int64_t* d_recv, size_t count_per_rank, cudaStream_t s) {
    if (!c || c->nranks == 1) {
        if (d_send != d_recv)
            IHTB_CUDA(cudaMemcpyAsync(d_recv, d_send, count_per_rank * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
        return;
    }
    if (!p2p_mapped(c)) { nccl_allgather_i64(c, d_send, d_recv, count_per_rank, s); return; }
    if (count_per_rank <= c->gat_cap) { p2p_allgather(c, d_send, count_per_rank, d_recv, s); return; }
    // larger than the mapped block: gather chunk by chunk into a scratch and scatter to the [rank][count] layout
    DBuf<int64_t> tmp((size_t)c->nranks * c->gat_cap);
    for (size_t o = 0; o < count_per_rank; o += c->gat_cap) {
        const size_t m = std::min(c->gat_cap, count_per_rank - o);
        p2p_allgather(c, d_send + o, m, tmp.p, s);
        for (int r = 0; r < c->nranks; ++r)
            IHTB_CUDA(cudaMemcpyAsync(d_recv + (size_t)r * count_per_rank + o, tmp.p + (size_t)r * m, m * sizeof(int64_t),
                                      cudaMemcpyDeviceToDevice, s));
    }
    IHTB_CUDA(cudaStreamSynchronize(s));         // tmp is freed on return
}

}  // namespace ihtb

using namespace ihtb;

extern "C" int32_t ihtb_comm_allreduce_bench(ihtb_comm* c, int64_t n, int32_t reps, int32_t use_p2p, double* us_per_op) {
    return guard([&] {
        IHTB_CHECK(c && n > 0 && reps > 0 && us_per_op, IHTB_EINVAL, "bad argument");
        IHTB_CUDA(cudaSetDevice(c->device));
        cudaStream_t s = nullptr;
        DBuf<double> src((size_t)n), dst((size_t)n);
        IHTB_CUDA(cudaMemset(src.p, 0, (size_t)n * sizeof(double)));
        // use_p2p: 0 = ncclAllReduce, 1 = the path sharded fits take for this n (push-all or two-phase),
        //          2 = force push-all, 3 = force two-phase
        if (use_p2p) p2p_setup(c, (size_t)n, 64, s);
        IHTB_CHECK(!use_p2p || (p2p_mapped(c) && (size_t)n <= c->red_cap), IHTB_EUNSUPPORTED, "peer-memory path is not available");
        IHTB_CHECK(use_p2p != 2 || (size_t)n <= c->pa_cap, IHTB_EUNSUPPORTED, "vector too long for the push-all path");
        IHTB_CHECK(use_p2p || c->comm, IHTB_EUNSUPPORTED, "this communicator has no NCCL backend");
        const bool pushall = use_p2p == 2 || (use_p2p == 1 && (size_t)n <= c->pa_cap);
        cudaEvent_t e0, e1;
        IHTB_CUDA(cudaEventCreate(&e0)); IHTB_CUDA(cudaEventCreate(&e1));
        for (int it = -10; it < reps; ++it) {
            if (it == 0) IHTB_CUDA(cudaEventRecord(e0, s));
            if (!use_p2p) nccl_allreduce_sum_f64(c, src.p, (size_t)n, s);     // in place, as in the fit
            else if (pushall) { p2p_push(c, src.p, (size_t)n, s); p2p_reduce(c, dst.p, (size_t)n, s); }
            else {
                IHTB_CUDA(cudaMemcpyAsync(p2p_partial_ptr(c), src.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
                p2p_allreduce_2phase(c, (size_t)n, dst.p, s);
            }
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
