// Projection of the group sums onto the Nystrom basis, fused with everything between K1 and the CAR kernel
// (SOBER/_rchq.py:148-166 and the ones-column of :229):
//
//   design[g, 0]     = 1
//   design[g, 1 + j] = ( sum_l (At[g, l] + [g == S-1] tail[l]) * Uext[j, l] ) / (totw[g] + [g == S-1] tail_tw)
//   totw_out[g]      =   totw[g] + [g == S-1] tail_tw
//
// i.e. X_tmp = (U_svd @ X_for_nys)^T, the second count of the remainder into the last group, the division by the
// group masses and the concatenation with the ones column, in one kernel.  The contraction runs on the FP64 tensor
// pipe: mma.sync.aligned.m8n8k4 f64 (SASS DMMA) -- tcgen05 has no FP64 kind.  CTA = 4 warps = 32 x 32 output tile,
// each warp 16 x 16 (2 x 2 MMA tiles), operands staged through shared memory in K-chunks of 32.
#include "common.cuh"

namespace sober {

constexpr int PJ_T = 32;    // output tile edge
constexpr int PJ_K = 32;    // K chunk

__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) project_design_kernel(const double* __restrict__ At, int64_t lda,
                                                             const double* __restrict__ tail,
                                                             const double* __restrict__ totw,
                                                             const double* __restrict__ tail_tw,
                                                             const double* __restrict__ Uext, int64_t ldu, int S, int Lp,
                                                             int n, double* __restrict__ design, int64_t ldd,
                                                             double* __restrict__ totw_out) {
    __shared__ double As[PJ_T][PJ_K + 1];   // [g][k]
    __shared__ double Bs[PJ_T][PJ_K + 1];   // [j][k]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g0 = blockIdx.y * PJ_T, j0 = blockIdx.x * PJ_T;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 16;   // warp quadrant
    double acc[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int k0 = 0; k0 < Lp; k0 += PJ_K) {
        for (int e = t; e < PJ_T * PJ_K; e += 128) {
            const int r = e / PJ_K, k = e - r * PJ_K;
            const int g = g0 + r, l = k0 + k, j = j0 + r;
            double av = 0.0, bv = 0.0;
            if (g < S && l < Lp) {
                av = At[(int64_t)g * lda + l];
                if (tail && g == S - 1) av += tail[l];
            }
            if (j < n && l < Lp) bv = Uext[(int64_t)j * ldu + l];
            As[r][k] = av;
            Bs[r][k] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < PJ_K; kk += 4) {
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                a[i] = As[wr + 8 * i + (lane >> 2)][kk + (lane & 3)];   // A fragment: row lane/4, col lane%4
                b[i] = Bs[wc + 8 * i + (lane >> 2)][kk + (lane & 3)];   // B fragment (col-major): k lane%4, n lane/4
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) dmma_8x8x4(acc[i][jj][0], acc[i][jj][1], a[i], b[jj]);
        }
        __syncthreads();
    }
    // epilogue: C fragment: row lane/4, columns 2 (lane%4), 2 (lane%4) + 1
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int g = g0 + wr + 8 * i + (lane >> 2);
        if (g >= S) continue;
        double tw = totw[g];
        if (tail_tw && g == S - 1) tw += tail_tw[0];
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int j = j0 + wc + 8 * jj + 2 * (lane & 3) + c;
                if (j < n) design[(int64_t)g * ldd + 1 + j] = acc[i][jj][c] / tw;
            }
        }
        if (blockIdx.x == 0 && wc == 0 && (lane & 3) == 0) {
            design[(int64_t)g * ldd] = 1.0;
            if (totw_out) totw_out[g] = tw;
        }
    }
}

// FP64 tensor-pipe throughput probe: 8 independent accumulator pairs per warp, back-to-back DMMA m8n8k4
__global__ void __launch_bounds__(256) dmma_probe_kernel(long long iters, double* sink) {
    double c[8][2];
#pragma unroll
    for (int q = 0; q < 8; ++q) c[q][0] = c[q][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999;
    for (long long i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) dmma_8x8x4(c[q][0], c[q][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += c[q][0] + c[q][1];
    if (s == 123.456) sink[0] = s;
}

}  // namespace sober

using namespace sober;

// blocks x 256 threads, each warp iters * 8 DMMA m8n8k4 (= 512 flop each): flops = blocks * 8 * iters * 8 * 512
extern "C" int sober_dmma_probe(int32_t blocks, int64_t iters, double* sink, void* stream) {
    if (blocks <= 0 || iters <= 0 || !sink) return SOBER_ERR_ARG;
    dmma_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, sink);
    SOBER_LAUNCH_CHECK("dmma_probe");
    return SOBER_OK;
}


extern "C" int sober_project_design(const double* At, int64_t lda, const double* tail, const double* totw,
                                    const double* tail_tw, const double* Uext, int64_t ldu, int32_t S, int32_t Lp,
                                    int32_t n, double* design, int64_t ldd, double* totw_out, void* stream) {
    if (S <= 0 || Lp <= 0 || n <= 0 || !At || !totw || !Uext || !design || lda < Lp || ldu < Lp || ldd < n + 1)
        return SOBER_ERR_ARG;
    dim3 grid((unsigned)ceil_div(n, PJ_T), (unsigned)ceil_div(S, PJ_T));
    project_design_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(At, lda, tail, totw, tail_tw, Uext, ldu, S, Lp, n, design,
                                                                 ldd, totw_out);
    SOBER_LAUNCH_CHECK("project_design");
    return SOBER_OK;
}
