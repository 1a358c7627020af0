// pi evaluation over the candidate set: SOBER/_pi.py:20-38 on top of the GP posterior
// SOBER/_gp.py:212-238.
//
//   mean_i = c + sum_j K[i, j] alpha_j                       K = k(x_i, Xobs_j): one K1 launch per chunk (Gram mode)
//   var_i  = kxx_i - sum_j K[i, j] T[i, j] + noise           T = K W (W = (K_obs + noise I)^-1, symmetric)
//   pi_i   = Phi((mean_i - eta) / sqrt(var_i))
//
// One warp per candidate row: both dot products, the clamp of the variance and the normal CDF in one streaming pass over
// the two (m x n_obs) tiles -- HBM-bound, 16 n_obs bytes read and 24 bytes written per candidate.
#include "common.cuh"

namespace sober {

__global__ void __launch_bounds__(256) gp_rows_kernel(const double* __restrict__ K, int64_t ldk,
                                                      const double* __restrict__ T, int64_t ldt,
                                                      const double* __restrict__ alpha, int64_t m, int n_obs,
                                                      double mean_const, const double* __restrict__ kxx,
                                                      double kxx_const, double noise, double min_var, double eta,
                                                      double* __restrict__ mean, double* __restrict__ var,
                                                      double* __restrict__ pi) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= m) return;
    const double* kr = K + row * ldk;
    const double* tr = T ? T + row * ldt : nullptr;
    double a = 0.0, q = 0.0;
    for (int j = lane; j < n_obs; j += 32) {
        const double kv = kr[j];
        a = fma(kv, __ldg(alpha + j), a);
        if (tr) q = fma(kv, tr[j], q);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane == 0) {
        const double mu = mean_const + a;
        double v = (kxx ? kxx[row] : kxx_const) - q + noise;
        v = fmax(v, min_var);
        if (mean) mean[row] = mu;
        if (var) var[row] = v;
        if (pi) pi[row] = normcdf((mu - eta) * rsqrt(v));
    }
}

}  // namespace sober

using namespace sober;

extern "C" int sober_gp_rows(const double* K, int64_t ldk, const double* T, int64_t ldt, const double* alpha, int64_t m,
                             int32_t n_obs, double mean_const, const double* kxx, double kxx_const, double noise,
                             double min_var, double eta, double* mean, double* var, double* pi, void* stream) {
    if (!K || !alpha || m < 0 || n_obs <= 0 || ldk < n_obs || (T && ldt < n_obs)) return SOBER_ERR_ARG;
    if (m == 0) return SOBER_OK;
    gp_rows_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(K, ldk, T, ldt, alpha, m, n_obs, mean_const,
                                                                            kxx, kxx_const, noise, min_var, eta, mean,
                                                                            var, pi);
    SOBER_LAUNCH_CHECK("gp_rows");
    return SOBER_OK;
}
