#pragma once
#include "common.cuh"

struct ihtb_comm;

namespace ihtb {

constexpr int DB_MAX_K = 256;     // largest support the debiasing refit handles

// device / pinned scratch of the debiasing refit, grown on demand and kept with the fit workspace
struct DebiasWs {
    DBuf<double> xk;       // n x (k + 1): decoded support columns, then the working residual / response
    DBuf<double> w;        // working weights
    DBuf<double> gpart, G, dpart, beta;
    HBuf<double> hG, hbeta;
    void ensure(int64_t n, int k);
};

// GLM.jl-default IRLS of y on x[:, cols] (k local column indices on the device, -1 = column of another shard when
// comm != NULL); beta_out[k] on the host
void debias_irls(const ihtb_geno* g, const double* d_y, int dist, int link, double nb_r, const int64_t* d_cols, int k,
                 double* beta_out, DebiasWs& ws, cudaStream_t s, ihtb_comm* comm = nullptr);

}  // namespace ihtb
