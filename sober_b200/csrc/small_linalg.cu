// Small dense helpers for the N-independent steps (Cholesky-QR of the Nystrom range finder and of the projector
// null space).  cuBLAS/cuSOLVER spend ~0.1 ms per call on these q ~ 200 problems (single-CTA, latency-bound
// kernels); they are called ~40 times per recombination.  (A one-CTA packed-triangle Cholesky was also tried here:
// 0.32 ms per call at q = 200 against 0.11 ms for cuSOLVER's potrf -- three block barriers per column on a 1024-thread
// CTA -- and was dropped; the factorisation stays with torch.linalg.cholesky_ex.)
#include "common.cuh"

namespace sober {

// X * R = Y  for upper-triangular R (q x q, row-major, q <= 256): one WARP per row of Y.
// Lane l holds elements k = 32 r + l of its row in registers y[r]; step j broadcasts x_j = y_j / R_jj by shuffle and
// applies y_k -= x_j R_jk to the elements right of j.  R is staged through shared memory 16 rows at a time (a step
// touches a NEW row of R: read straight from global memory every step pays an L2 round trip -- measured 86 us per
// call at q = 200, no better than cuBLAS; staged: the dependent chain is shuffle -> mul -> fma).
constexpr int TR_TILE = 16;
template <int NR>
__global__ void __launch_bounds__(256) trsm_right_upper_kernel(const double* __restrict__ Y, int64_t ldy,
                                                               const double* __restrict__ R, int64_t ldr, int m, int q,
                                                               double* __restrict__ X, int64_t ldx) {
    __shared__ double rt[TR_TILE][32 * NR];
    __shared__ double rinv[32 * NR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = threadIdx.x; j < 32 * NR; j += blockDim.x) rinv[j] = j < q ? 1.0 / R[(int64_t)j * ldr + j] : 0.0;
    const int row = blockIdx.x * (blockDim.x >> 5) + warp;
    const bool live = row < m;
    double y[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int k = 32 * r + lane;
        y[r] = (live && k < q) ? Y[(int64_t)row * ldy + k] : 0.0;
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll
        for (int half = 0; half < 32 / TR_TILE; ++half) {
            const int j0 = 32 * r + TR_TILE * half;
            __syncthreads();   // previous tile consumed (also orders the rinv fill before its first use)
            if (j0 < q) {
                for (int e = threadIdx.x; e < TR_TILE * 32 * NR; e += blockDim.x) {
                    const int jr = e / (32 * NR), k = e - jr * (32 * NR);
                    const int j = j0 + jr;
                    rt[jr][k] = (j < q && k < q && k > j) ? R[(int64_t)j * ldr + k] : 0.0;
                }
            }
            __syncthreads();
            if (j0 < q) {
#pragma unroll
                for (int jr = 0; jr < TR_TILE; ++jr) {
                    const int j = j0 + jr;
                    const int jj = TR_TILE * half + jr;       // lane that holds element j of register r
                    const double xj = __shfl_sync(0xffffffffu, y[r], jj) * rinv[j];
                    if (lane == jj) y[r] = xj;
#pragma unroll
                    for (int rr = r; rr < NR; ++rr) y[rr] = fma(-xj, rt[jr][32 * rr + lane], y[rr]);   // zeros left of j
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int k = 32 * r + lane;
            if (k < q) X[(int64_t)row * ldx + k] = y[r];
        }
    }
}

}  // namespace sober

using namespace sober;

extern "C" int sober_trsm_right_upper(const double* Y, int64_t ldy, const double* R, int64_t ldr, int32_t m, int32_t q,
                                      double* X, int64_t ldx, void* stream) {
    if (m < 0 || q <= 0 || q > 256 || !Y || !R || !X || ldy < q || ldr < q || ldx < q) return SOBER_ERR_ARG;
    if (m == 0) return SOBER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nr = (q + 31) / 32;
    const dim3 grid((unsigned)ceil_div(m, 8));
    switch (nr) {
        case 1: trsm_right_upper_kernel<1><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        case 2: trsm_right_upper_kernel<2><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        case 3: trsm_right_upper_kernel<3><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        case 4: trsm_right_upper_kernel<4><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        case 5: trsm_right_upper_kernel<5><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        case 6: trsm_right_upper_kernel<6><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        case 7: trsm_right_upper_kernel<7><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
        default: trsm_right_upper_kernel<8><<<grid, 256, 0, st>>>(Y, ldy, R, ldr, m, q, X, ldx); break;
    }
    SOBER_LAUNCH_CHECK("trsm_right_upper");
    return SOBER_OK;
}
