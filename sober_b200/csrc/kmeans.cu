// Assignment step of Lloyd's algorithm -- SURVEY.md 8(f) row 2, ``KMeans`` of SOBER/_weights.py:100-126, the producer of
// the Nystrom landmarks ``pts_nys`` for continuous domains (SOBER/_sampler.py:316-317).  The reference broadcasts an
// (N, K, D) difference tensor (48 GB at N = 1e6, K = 1000, D = 6, f64); here each thread keeps one point in registers,
// the centroids stream through shared memory (every lane reads the same centroid: broadcast loads) and only the (N,)
// labels are written: 8 N bytes of traffic for 2 N K D flops -- FP64-pipe bound like K1.
//
//   label_i = argmin_k sum_d (x_id - c_kd)^2, first minimum on ties, first NaN wins (torch.argmin's rule: an empty
//   cluster of the previous iteration is a NaN centroid, SOBER/_weights.py:122-124)
#include "common.cuh"

namespace sober {

constexpr int KM_THREADS = 256;
constexpr int KM_SMEM_DOUBLES = 6144;   // 48 KB of centroids per chunk

template <int D>
__global__ void __launch_bounds__(KM_THREADS) kmeans_assign_kernel(const double* __restrict__ X, int64_t ldx, int64_t n,
                                                                  const double* __restrict__ C, int K,
                                                                  int64_t* __restrict__ labels) {
    __shared__ double cs[KM_SMEM_DOUBLES];
    constexpr int KC = KM_SMEM_DOUBLES / D;
    const int64_t stride = (int64_t)gridDim.x * KM_THREADS;
    const int64_t batches = (n + stride - 1) / stride;      // uniform over the grid: the chunk loop has block barriers
    for (int64_t b = 0; b < batches; ++b) {
        const int64_t i = b * stride + (int64_t)blockIdx.x * KM_THREADS + threadIdx.x;
        const bool live = i < n;
        double x[D];
#pragma unroll
        for (int k = 0; k < D; ++k) x[k] = live ? X[i * ldx + k] : 0.0;
        double best = __longlong_as_double(0x7ff0000000000000ll);   // +inf
        int arg = 0;
        for (int k0 = 0; k0 < K; k0 += KC) {
            const int kc = min(KC, K - k0);
            __syncthreads();                                 // previous chunk consumed
            for (int e = threadIdx.x; e < kc * D; e += KM_THREADS) cs[e] = C[(int64_t)k0 * D + e];
            __syncthreads();
            int k = 0;
            for (; k + 4 <= kc; k += 4) {                    // four independent distance chains per trip
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int dd = 0; dd < D; ++dd) {
                        const double diff = x[dd] - cs[(k + u) * D + dd];
                        acc[u] = fma(diff, diff, acc[u]);
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (best == best && !(acc[u] >= best)) {  // strictly smaller, or the first NaN (which then sticks)
                        best = acc[u];
                        arg = k0 + k + u;
                    }
            }
            for (; k < kc; ++k) {
                double acc = 0.0;
#pragma unroll
                for (int dd = 0; dd < D; ++dd) {
                    const double diff = x[dd] - cs[k * D + dd];
                    acc = fma(diff, diff, acc);
                }
                if (best == best && !(acc >= best)) {
                    best = acc;
                    arg = k0 + k;
                }
            }
        }
        if (live) labels[i] = arg;
    }
}

}  // namespace sober

using namespace sober;

extern "C" int sober_kmeans_assign(const double* X, int64_t ldx, int64_t n, int32_t d, const double* C, int32_t K,
                                   int64_t* labels, void* stream) {
    if (n < 0 || d <= 0 || K <= 0 || !C || ldx < d || (n > 0 && (!X || !labels))) return SOBER_ERR_ARG;
    if (d > 16) return SOBER_ERR_UNSUPPORTED;
    if (n == 0) return SOBER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = ceil_div(n, (int64_t)KM_THREADS);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
#define KM_LAUNCH(D_) \
    case D_: kmeans_assign_kernel<D_><<<(unsigned)blocks, KM_THREADS, 0, st>>>(X, ldx, n, C, K, labels); break;
    switch (d) {
        KM_LAUNCH(1) KM_LAUNCH(2) KM_LAUNCH(3) KM_LAUNCH(4) KM_LAUNCH(5) KM_LAUNCH(6) KM_LAUNCH(7) KM_LAUNCH(8)
        KM_LAUNCH(9) KM_LAUNCH(10) KM_LAUNCH(11) KM_LAUNCH(12) KM_LAUNCH(13) KM_LAUNCH(14) KM_LAUNCH(15) KM_LAUNCH(16)
        default: return SOBER_ERR_UNSUPPORTED;
    }
#undef KM_LAUNCH
    SOBER_LAUNCH_CHECK("kmeans_assign");
    return SOBER_OK;
}
