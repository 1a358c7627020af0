// Small fused helpers around one Caratheodory step (SOBER/_rchq.py:166, 229-230 and the bookkeeping of :198-221): each
// replaces a handful of library elementwise / reduction launches of ~4 us inside the per-step CUDA graph (the ncu launch
// list of a C2 step had ~450 of them, 19 % of the GPU time).
#include "common.cuh"

namespace sober {

// scaled[i, 0] = 1 / sqrt(S);  scaled[i, 1 + j] = (F[i, j] / div[i]) / ||column||   -- barycentres (SOBER/_rchq.py:166),
// the ones column of the design matrix (:229) and the column normalisation of the projector null space in one pass.
// 32 columns per CTA, 32 row lanes per column (the FP64 divisions are the cost: 2 S / 32 per thread).
constexpr int PREP_RL = 32;
__global__ void __launch_bounds__(32 * PREP_RL) car_prepare_kernel(const double* __restrict__ F, int64_t ldf,
                                                                   const double* __restrict__ div, int S, int n,
                                                                   double* __restrict__ out, int64_t ldo) {
    __shared__ double part[PREP_RL][33];
    __shared__ double inv_norm[32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;          // column of the design matrix, 0 = ones
    double ss = 0.0;
    if (c >= 1 && c <= n) {
        for (int i = ty; i < S; i += PREP_RL) {
            double v = F[(int64_t)i * ldf + (c - 1)];
            if (div) v /= div[i];
            ss = fma(v, v, ss);
        }
    }
    part[ty][tx] = ss;
    __syncthreads();
    if (ty == 0) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < PREP_RL; ++q) s += part[q][tx];
        if (c == 0) s = (double)S;
        inv_norm[tx] = 1.0 / fmax(sqrt(s), 1e-300);
    }
    __syncthreads();
    if (c <= n) {
        const double sc = inv_norm[tx];
        for (int i = ty; i < S; i += PREP_RL) {
            double v = 1.0;
            if (c >= 1) {
                v = F[(int64_t)i * ldf + (c - 1)];
                if (div) v /= div[i];
            }
            out[(int64_t)i * ldo + c] = v * sc;
        }
    }
}

// Second count of the remainder (SOBER/_rchq.py:153-164) folded into the group sums in one launch:
//   at[S-1, :] += tail_at (tail_at may be NULL: no remainder on this call);  totw_out = totw_in, + tail_tw[0] at S-1.
__global__ void __launch_bounds__(256) apply_tail_kernel(double* __restrict__ at_last, const double* __restrict__ tail_at,
                                                         int Lp, const double* __restrict__ totw_in,
                                                         const double* __restrict__ tail_tw, int S,
                                                         double* __restrict__ totw_out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (tail_at && i < Lp) at_last[i] += tail_at[i];
    if (i < S) totw_out[i] = totw_in[i] + ((tail_tw && i == S - 1) ? tail_tw[0] : 0.0);
}

// After the elimination: poison the weights when the one-pass Cholesky-QR behind the projector was not accurate enough
// (|Delta|_F >= 1e-5) or something is not finite, then the survivor bookkeeping the host reads with ONE copy:
//   summary[i] = number of kept groups among 0..i (inclusive), summary[S] = 1 if all weights are finite, rank[i] =
//   number of kept groups below i.  One CTA, S <= 4096.
__global__ void __launch_bounds__(1024) car_summary_kernel(double* __restrict__ w, int S, const double* __restrict__ delta,
                                                           int64_t ndelta, double defect_limit, int* __restrict__ summary,
                                                           int* __restrict__ rank) {
    __shared__ double red[32];
    __shared__ int redi[32];
    __shared__ int warp_tot[32];
    __shared__ int bad_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    double ss = 0.0;
    for (int64_t e = t; e < ndelta; e += 1024) ss = fma(delta[e], delta[e], ss);
    int finite = 1;
    for (int i = t; i < S; i += 1024) finite &= isfinite(w[i]) ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
        finite &= __shfl_xor_sync(0xffffffffu, finite, off);
    }
    if (lane == 0) { red[warp] = ss; redi[warp] = finite; }
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        int f = 1;
        for (int q = 0; q < 32; ++q) { s += red[q]; f &= redi[q]; }
        const bool ok = f && (ndelta == 0 || (s < defect_limit * defect_limit));   // NaN in delta -> not ok
        bad_s = ok ? 0 : 1;
    }
    __syncthreads();
    const bool bad = bad_s != 0;
    // inclusive scan of the kept flags, 4 consecutive groups per thread
    int k[4], local = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = 4 * t + q;
        double v = i < S ? w[i] : 0.0;
        if (bad && i < S) { v = __longlong_as_double(0x7ff8000000000000ll); w[i] = v; }
        k[q] = (i < S && v > 0.0) ? 1 : 0;
        local += k[q];
    }
    int incl = local;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = warp_tot[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int up = __shfl_up_sync(0xffffffffu, v, off);
            if (lane >= off) v += up;
        }
        warp_tot[lane] = v;
    }
    __syncthreads();
    int run = incl - local + (warp > 0 ? warp_tot[warp - 1] : 0);   // kept groups before this thread's first group
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = 4 * t + q;
        if (i < S) {
            rank[i] = run;
            run += k[q];
            summary[i] = run;
        }
    }
    if (t == 0) summary[S] = bad ? 0 : 1;
}

}  // namespace sober

using namespace sober;

extern "C" int sober_car_prepare(const double* F, int64_t ldf, const double* div, int32_t S, int32_t n, double* out,
                                 int64_t ldo, void* stream) {
    if (!F || !out || S <= 0 || n < 0 || ldf < n || ldo < n + 1) return SOBER_ERR_ARG;
    car_prepare_kernel<<<(unsigned)ceil_div(n + 1, 32), 32 * PREP_RL, 0, (cudaStream_t)stream>>>(F, ldf, div, S, n, out, ldo);
    SOBER_LAUNCH_CHECK("car_prepare");
    return SOBER_OK;
}

extern "C" int sober_apply_tail(double* at_last_row, const double* tail_at, int32_t Lp, const double* totw_in,
                                const double* tail_tw, int32_t S, double* totw_out, void* stream) {
    if (!at_last_row || !totw_in || !totw_out || S <= 0 || Lp <= 0 || (tail_at == nullptr) != (tail_tw == nullptr))
        return SOBER_ERR_ARG;
    const int n = S > Lp ? S : Lp;
    apply_tail_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(at_last_row, tail_at, Lp, totw_in, tail_tw,
                                                                                  S, totw_out);
    SOBER_LAUNCH_CHECK("apply_tail");
    return SOBER_OK;
}

extern "C" int sober_car_summary(double* w, int32_t S, const double* delta, int64_t ndelta, double defect_limit,
                                 int32_t* summary, int32_t* rank, void* stream) {
    if (!w || !summary || !rank || S <= 0 || S > 4096 || ndelta < 0 || (ndelta > 0 && !delta)) return SOBER_ERR_ARG;
    car_summary_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(w, S, delta, ndelta, defect_limit, summary, rank);
    SOBER_LAUNCH_CHECK("car_summary");
    return SOBER_OK;
}
