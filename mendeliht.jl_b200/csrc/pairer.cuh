// Lock-step pairing of the X'r sweeps of two fits that run on the same device (cross-validation grid, reference
// src/cross_validation.jl:98-121: independent fits on masks of the SAME matrix).  Each fit runs on its own host thread
// and stream; when both reach their sweep, ONE pass over the matrix serves both residual vectors (PAIR sweep, half2
// lookup tables, sweep_lut.cu) instead of two passes.  A fit whose partner has finished sweeps alone (FAST).
#pragma once
#include "common.cuh"
#include <condition_variable>
#include <mutex>

namespace ihtb {

struct SweepPairer {
    std::mutex mu;
    std::condition_variable cv;
    int active = 2;                        // fits still running on this device
    struct Req {
        const double* v; const double* vbar; double* out; cudaStream_t s;
    } req[2] = {};
    bool has[2] = {false, false};
    unsigned long long round = 0;
    cudaEvent_t ready[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;
    void* scratch = nullptr;               // SweepScratch (sweep.cu)
    double* d_vbar2 = nullptr;
    double* d_l2 = nullptr;                // ||v - vbar||_2 of the two vectors of the last pair sweep
    int64_t n_pair = 0, n_solo = 0;
    int device = 0;

    explicit SweepPairer(int device);
    ~SweepPairer();
    SweepPairer(const SweepPairer&) = delete;
    SweepPairer& operator=(const SweepPairer&) = delete;
    // df = X'(v - vbar) for the calling fit (slot 0 or 1), enqueued on its stream `s`; returns true when the result
    // comes from a PAIR sweep: the error bound is then kPairBound * ||u||_2 * sgn_j with ||u||_2 at d_l2[slot]
    bool sweep(int slot, const ihtb_geno* g, const double* d_v, const double* d_vbar, double* d_out, cudaStream_t s,
               void* solo_scratch);
    void leave();                          // the calling fit runs no more sweeps
};

}  // namespace ihtb
