// K3b -- Caratheodory elimination on a given null-space basis (SOBER/_rchq.py:237-266).
//
// The reference runs k dependent steps, each ~12 tiny torch ops + host syncs.  Here the whole loop is ONE
// persistent cooperative kernel:
//   * basis column c (= row c of the k x S input) is owned by CTA (c mod G) and lives in that CTA's shared
//     memory for its whole life (cyclic ownership keeps the shrinking work balanced);
//   * at step s the owner publishes column s to global memory and raises flag[s] (release); every CTA waits
//     on that flag (acquire), reads the column through L2 and REDUNDANTLY finds the pivot and updates its
//     private copy of the weights -- the arithmetic is deterministic so all copies stay bit-identical and no
//     second exchange is needed;
//   * each CTA then applies the rank-1 update to the columns it owns, the next pivot column first so that its
//     publication overlaps everybody else's updates (software pipeline; no grid-wide barrier anywhere).
// Arithmetic order per element is exactly the reference's: ratio = mu/v ; mu - (ratio_j * v) ;
// phi - (phi_j * v) / v_j, all unfused, so for the same basis the pivot sequence is bit-identical.
#include "common.cuh"

#include <stdlib.h>

namespace sober {

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Best {
    double ratio;
    int idx;
};
__device__ __forceinline__ Best better(Best a, Best b) {
    // first minimum in index order (torch.argmin over the compressed positive set)
    if (b.idx < 0) return a;
    if (a.idx < 0) return b;
    if (b.ratio < a.ratio || (b.ratio == a.ratio && b.idx < a.idx)) return b;
    return a;
}


template <bool SMEM, bool EXACT, int CAR_THREADS>
__global__ void __launch_bounds__(CAR_THREADS) car_eliminate_kernel(double* __restrict__ basis, int k, int S,
                                                                    double* __restrict__ mu_g,
                                                                    int* __restrict__ pivots, int* __restrict__ steps,
                                                                    int* flags) {
    extern __shared__ double sm[];
    double* mu_s = sm;
    double* v_s = sm + S;
    double* cols = sm + 2 * (int64_t)S;  // [slot][S] when SMEM
    __shared__ Best red[CAR_THREADS / 32];
    __shared__ Best chosen;

    const int G = gridDim.x, b = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < S; i += CAR_THREADS) mu_s[i] = mu_g[i];
    if (SMEM) {
        for (int c = b, slot = 0; c < k; c += G, ++slot)
            for (int i = t; i < S; i += CAR_THREADS) cols[(int64_t)slot * S + i] = basis[(int64_t)c * S + i];
    }
    if (b == 0 && t == 0) st_release(flags + 0, 1);  // column 0 is already in place in global memory
    __syncthreads();

    int done = 0;
    for (int s = 0; s < k; ++s) {
        const int owner = s % G;
        if (b == owner) {
            const double* src = SMEM ? cols + (int64_t)(s / G) * S : basis + (int64_t)s * S;
            for (int i = t; i < S; i += CAR_THREADS) v_s[i] = src[i];
        } else {
            if (t == 0) {
                while (ld_acquire(flags + s) == 0) {
                }
            }
            __syncthreads();
            const double* src = basis + (int64_t)s * S;
            for (int i = t; i < S; i += CAR_THREADS) v_s[i] = __ldcg(src + i);
        }
        __syncthreads();

        // pivot: argmin over {i : v_i > 0} of mu_i / v_i, first minimum
        Best mine{0.0, -1};
        for (int i = t; i < S; i += CAR_THREADS) {
            const double v = v_s[i];
            if (v > 0.0) {
                Best c{__ddiv_rn(mu_s[i], v), i};
                mine = better(mine, c);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            Best o;
            o.ratio = __shfl_down_sync(0xffffffffu, mine.ratio, off);
            o.idx = __shfl_down_sync(0xffffffffu, mine.idx, off);
            mine = better(mine, o);
        }
        if ((t & 31) == 0) red[t >> 5] = mine;
        __syncthreads();
        if (t < 32) {
            Best r = t < CAR_THREADS / 32 ? red[t] : Best{0.0, -1};
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                Best o;
                o.ratio = __shfl_down_sync(0xffffffffu, r.ratio, off);
                o.idx = __shfl_down_sync(0xffffffffu, r.idx, off);
                r = better(r, o);
            }
            if (t == 0) chosen = r;
        }
        __syncthreads();
        const Best piv = chosen;
        if (piv.idx < 0) break;  // no positive entry: the guard of SOBER/_rchq.py:241-242
        const int j = piv.idx;
        const double vj = v_s[j];
        if (!EXACT) __syncthreads();   // everybody holds v_j before v_s is overwritten with t_i below
        if (b == 0 && t == 0 && pivots) pivots[s] = j;
        done = s + 1;

        for (int i = t; i < S; i += CAR_THREADS) {
            const double vi = v_s[i];
            const double m = __dsub_rn(mu_s[i], __dmul_rn(piv.ratio, vi));
            mu_s[i] = (i == j) ? 0.0 : m;
            // fast variant: one division per ROW (t_i = v_i / v_j) instead of one per matrix element
            if (!EXACT) v_s[i] = __ddiv_rn(vi, vj);
        }
        if (!EXACT) __syncthreads();

        // rank-1 update of the columns this CTA owns, the next pivot column first
        const int nxt = s + 1;
        if (nxt < k && (nxt % G) == b) {
            double* col = SMEM ? cols + (int64_t)(nxt / G) * S : basis + (int64_t)nxt * S;
            const double pj = col[j];
            __syncthreads();  // everyone has read col[j] before it is overwritten
            for (int i = t; i < S; i += CAR_THREADS) {
                const double upd = EXACT ? __dsub_rn(col[i], __ddiv_rn(__dmul_rn(pj, v_s[i]), vj))
                                         : fma(-pj, v_s[i], col[i]);
                const double val = (i == j) ? 0.0 : upd;
                col[i] = val;
                if (SMEM) basis[(int64_t)nxt * S + i] = val;
            }
            __syncthreads();
            if (t == 0) {
                __threadfence();
                st_release(flags + nxt, 1);
            }
        }
        for (int c = b + ((nxt + 1 - b + G - 1) / G) * G; c < k; c += G) {  // first owned column >= s + 2
            double* col = SMEM ? cols + (int64_t)(c / G) * S : basis + (int64_t)c * S;
            const double pj = col[j];
            __syncthreads();
            for (int i = t; i < S; i += CAR_THREADS) {
                const double upd = EXACT ? __dsub_rn(col[i], __ddiv_rn(__dmul_rn(pj, v_s[i]), vj))
                                         : fma(-pj, v_s[i], col[i]);
                col[i] = (i == j) ? 0.0 : upd;
            }
        }
        __syncthreads();  // v_s / mu_s are rewritten at the top of the next step
    }

    if (b == 0) {
        for (int i = t; i < S; i += CAR_THREADS) mu_g[i] = mu_s[i];
        if (t == 0) {
            if (steps) *steps = done;
            if (pivots)
                for (int s = done; s < k; ++s) pivots[s] = -1;
        }
    }
}

}  // namespace sober

using namespace sober;

extern "C" int64_t sober_car_workspace(int32_t k) { return k > 0 ? (int64_t)k * 4 : 4; }

extern "C" int sober_car_eliminate(double* basis, int32_t k, int32_t S, double* mu, int32_t exact, int32_t* pivots_out,
                                   int32_t* steps_out, void* sync_ws, int64_t sync_ws_bytes, void* stream) {
    if (k < 0 || S <= 0 || !mu) return SOBER_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (k == 0) {
        if (steps_out) SOBER_CUDA_CHECK(cudaMemsetAsync(steps_out, 0, 4, st));
        return SOBER_OK;
    }
    if (!basis || !sync_ws) return SOBER_ERR_ARG;
    if (sync_ws_bytes < sober_car_workspace(k)) return SOBER_ERR_WORKSPACE;
    SOBER_CUDA_CHECK(cudaMemsetAsync(sync_ws, 0, (size_t)k * 4, st));

    int dev = 0, max_smem = 0, coop = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    SOBER_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SOBER_CUDA_CHECK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return SOBER_ERR_UNSUPPORTED;
    const int sms = sm_count();

    int G = k < sms ? k : sms;
    const int64_t slots = ceil_div(k, G);
    const int64_t smem_full = (2 + slots) * (int64_t)S * 8;
    const int64_t smem_lite = 2 * (int64_t)S * 8;
    // SOBER_B200_CAR_FORCE_GLOBAL: test hook that exercises the global-memory variant at small sizes
    const bool use_smem = smem_full + 1024 <= max_smem && getenv("SOBER_B200_CAR_FORCE_GLOBAL") == nullptr;
    if (!use_smem && smem_lite + 1024 > max_smem) return SOBER_ERR_UNSUPPORTED;
    const size_t smem = (size_t)(use_smem ? smem_full : smem_lite);

    // 1024 threads when the pivot search (S divisions per CTA per step) would otherwise dominate the step
    const int threads = S > 768 ? 1024 : 256;
    const void* fn;
    if (threads == 1024) {
        fn = use_smem ? (exact ? (const void*)car_eliminate_kernel<true, true, 1024> : (const void*)car_eliminate_kernel<true, false, 1024>)
                      : (exact ? (const void*)car_eliminate_kernel<false, true, 1024> : (const void*)car_eliminate_kernel<false, false, 1024>);
    } else {
        fn = use_smem ? (exact ? (const void*)car_eliminate_kernel<true, true, 256> : (const void*)car_eliminate_kernel<true, false, 256>)
                      : (exact ? (const void*)car_eliminate_kernel<false, true, 256> : (const void*)car_eliminate_kernel<false, false, 256>);
    }
    SOBER_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SOBER_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) return SOBER_ERR_UNSUPPORTED;
    if (G > per_sm * sms) G = per_sm * sms;  // cannot happen with G <= sms, kept as a guard

    int* flags = (int*)sync_ws;
    void* args[] = {&basis, &k, &S, &mu, &pivots_out, &steps_out, &flags};
    SOBER_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(threads), args, smem, st));
    return SOBER_OK;
}
