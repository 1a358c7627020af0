// K3 (cluster, column-distributed) -- the k dependent elimination steps of SOBER/_rchq.py:237-266 on a given
// null-space basis, on one 8-CTA thread-block cluster with the matrix resident in REGISTERS.
//
// Layout: column c of Phi (= row c of the k x S input) lives in CTA (c mod 8); inside the CTA it is spread over a
// group of 8 adjacent lanes (lane rs holds rows i = rs + 8*li, li < RMAX).  A CTA therefore owns ALL rows of its
// columns, so -- unlike the row-distributed kernel in car_cluster.cu -- a step needs no reduction across CTAs:
//   * the owner of pivot column s finds the pivot (argmin of mu_i / v_i over v_i > 0) with its whole block,
//   * broadcasts [v (S values), alpha, v_j, j] to the 7 peers with one bulk DSMEM copy each
//     (cp.async.bulk shared::cta -> shared::cluster, completing a transaction count on the receiver's mbarrier),
//   * every CTA updates its replicated copy of the weights and its own columns: the pivot-row entry Phi[j, c] is
//     already local (one register of the 8-lane group, fetched with a shuffle).
// The owner of column s + 1 simply runs ahead: its pivot search overlaps the peers' updates (software lookahead).
// Eight broadcast buffers / mbarriers (one per owner rank) make run-ahead of up to 7 steps safe.
// EXACT: the reference's unfused (phi_j * v_i) / v_j; otherwise one division per column and an FMA per element.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace sober {

constexpr int CE_THREADS = 256;
constexpr int CE_P = 8;

struct CarColsParams {
    const double* basis;   // k x S rows
    double* mu;            // S in/out
    int* info;             // [0] elimination steps taken
    long long* prof;       // optional 8 cycle counters (CTA 0, thread 0)
    int S, k, Spad, SV;    // Spad = S rounded up to even, SV = Spad + 4 (v | alpha, v_j, j, pad)
};

struct BestRI {
    double ratio;
    int idx;
    int pad;
};
__device__ __forceinline__ bool ri_better(double ra, int ia, double rb, int ib) {
    if (ib < 0) return false;
    if (ia < 0) return true;
    return rb < ra || (rb == ra && ib < ia);
}
__device__ __forceinline__ uint32_t ce_map_rank(uint32_t saddr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(saddr), "r"(rank));
    return out;
}
__device__ __forceinline__ void ce_bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}

template <int RMAX>
__device__ __forceinline__ double ce_col_get(const double (&col)[RMAX], int li) {
    double v = 0.0;
    switch (li) {
        case 0: if (0 < RMAX) v = col[0]; break;
        case 1: if (1 < RMAX) v = col[1]; break;
        case 2: if (2 < RMAX) v = col[2]; break;
        case 3: if (3 < RMAX) v = col[3]; break;
        case 4: if (4 < RMAX) v = col[4]; break;
        case 5: if (5 < RMAX) v = col[5]; break;
        case 6: if (6 < RMAX) v = col[6]; break;
        case 7: if (7 < RMAX) v = col[7]; break;
        case 8: if (8 < RMAX) v = col[8]; break;
        case 9: if (9 < RMAX) v = col[9]; break;
        case 10: if (10 < RMAX) v = col[10]; break;
        case 11: if (11 < RMAX) v = col[11]; break;
        case 12: if (12 < RMAX) v = col[12]; break;
        case 13: if (13 < RMAX) v = col[13]; break;
        case 14: if (14 < RMAX) v = col[14]; break;
        case 15: if (15 < RMAX) v = col[15]; break;
        case 16: if (16 < RMAX) v = col[16]; break;
        case 17: if (17 < RMAX) v = col[17]; break;
        case 18: if (18 < RMAX) v = col[18]; break;
        case 19: if (19 < RMAX) v = col[19]; break;
        case 20: if (20 < RMAX) v = col[20]; break;
        case 21: if (21 < RMAX) v = col[21]; break;
        case 22: if (22 < RMAX) v = col[22]; break;
        case 23: if (23 < RMAX) v = col[23]; break;
        case 24: if (24 < RMAX) v = col[24]; break;
        case 25: if (25 < RMAX) v = col[25]; break;
        case 26: if (26 < RMAX) v = col[26]; break;
        case 27: if (27 < RMAX) v = col[27]; break;
        case 28: if (28 < RMAX) v = col[28]; break;
        case 29: if (29 < RMAX) v = col[29]; break;
        case 30: if (30 < RMAX) v = col[30]; break;
        case 31: if (31 < RMAX) v = col[31]; break;
        case 32: if (32 < RMAX) v = col[32]; break;
        case 33: if (33 < RMAX) v = col[33]; break;
        case 34: if (34 < RMAX) v = col[34]; break;
        case 35: if (35 < RMAX) v = col[35]; break;
        case 36: if (36 < RMAX) v = col[36]; break;
        case 37: if (37 < RMAX) v = col[37]; break;
        case 38: if (38 < RMAX) v = col[38]; break;
        case 39: if (39 < RMAX) v = col[39]; break;
        case 40: if (40 < RMAX) v = col[40]; break;
        case 41: if (41 < RMAX) v = col[41]; break;
        case 42: if (42 < RMAX) v = col[42]; break;
        case 43: if (43 < RMAX) v = col[43]; break;
        case 44: if (44 < RMAX) v = col[44]; break;
        case 45: if (45 < RMAX) v = col[45]; break;
        case 46: if (46 < RMAX) v = col[46]; break;
        case 47: if (47 < RMAX) v = col[47]; break;
        case 48: if (48 < RMAX) v = col[48]; break;
        case 49: if (49 < RMAX) v = col[49]; break;
        case 50: if (50 < RMAX) v = col[50]; break;
        case 51: if (51 < RMAX) v = col[51]; break;
        case 52: if (52 < RMAX) v = col[52]; break;
        case 53: if (53 < RMAX) v = col[53]; break;
        case 54: if (54 < RMAX) v = col[54]; break;
        case 55: if (55 < RMAX) v = col[55]; break;
        default: break;
    }
    return v;
}
template <int RMAX>
__device__ __forceinline__ void ce_col_zero(double (&col)[RMAX], int li) {
    switch (li) {
        case 0: if (0 < RMAX) col[0] = 0.0; break;
        case 1: if (1 < RMAX) col[1] = 0.0; break;
        case 2: if (2 < RMAX) col[2] = 0.0; break;
        case 3: if (3 < RMAX) col[3] = 0.0; break;
        case 4: if (4 < RMAX) col[4] = 0.0; break;
        case 5: if (5 < RMAX) col[5] = 0.0; break;
        case 6: if (6 < RMAX) col[6] = 0.0; break;
        case 7: if (7 < RMAX) col[7] = 0.0; break;
        case 8: if (8 < RMAX) col[8] = 0.0; break;
        case 9: if (9 < RMAX) col[9] = 0.0; break;
        case 10: if (10 < RMAX) col[10] = 0.0; break;
        case 11: if (11 < RMAX) col[11] = 0.0; break;
        case 12: if (12 < RMAX) col[12] = 0.0; break;
        case 13: if (13 < RMAX) col[13] = 0.0; break;
        case 14: if (14 < RMAX) col[14] = 0.0; break;
        case 15: if (15 < RMAX) col[15] = 0.0; break;
        case 16: if (16 < RMAX) col[16] = 0.0; break;
        case 17: if (17 < RMAX) col[17] = 0.0; break;
        case 18: if (18 < RMAX) col[18] = 0.0; break;
        case 19: if (19 < RMAX) col[19] = 0.0; break;
        case 20: if (20 < RMAX) col[20] = 0.0; break;
        case 21: if (21 < RMAX) col[21] = 0.0; break;
        case 22: if (22 < RMAX) col[22] = 0.0; break;
        case 23: if (23 < RMAX) col[23] = 0.0; break;
        case 24: if (24 < RMAX) col[24] = 0.0; break;
        case 25: if (25 < RMAX) col[25] = 0.0; break;
        case 26: if (26 < RMAX) col[26] = 0.0; break;
        case 27: if (27 < RMAX) col[27] = 0.0; break;
        case 28: if (28 < RMAX) col[28] = 0.0; break;
        case 29: if (29 < RMAX) col[29] = 0.0; break;
        case 30: if (30 < RMAX) col[30] = 0.0; break;
        case 31: if (31 < RMAX) col[31] = 0.0; break;
        case 32: if (32 < RMAX) col[32] = 0.0; break;
        case 33: if (33 < RMAX) col[33] = 0.0; break;
        case 34: if (34 < RMAX) col[34] = 0.0; break;
        case 35: if (35 < RMAX) col[35] = 0.0; break;
        case 36: if (36 < RMAX) col[36] = 0.0; break;
        case 37: if (37 < RMAX) col[37] = 0.0; break;
        case 38: if (38 < RMAX) col[38] = 0.0; break;
        case 39: if (39 < RMAX) col[39] = 0.0; break;
        case 40: if (40 < RMAX) col[40] = 0.0; break;
        case 41: if (41 < RMAX) col[41] = 0.0; break;
        case 42: if (42 < RMAX) col[42] = 0.0; break;
        case 43: if (43 < RMAX) col[43] = 0.0; break;
        case 44: if (44 < RMAX) col[44] = 0.0; break;
        case 45: if (45 < RMAX) col[45] = 0.0; break;
        case 46: if (46 < RMAX) col[46] = 0.0; break;
        case 47: if (47 < RMAX) col[47] = 0.0; break;
        case 48: if (48 < RMAX) col[48] = 0.0; break;
        case 49: if (49 < RMAX) col[49] = 0.0; break;
        case 50: if (50 < RMAX) col[50] = 0.0; break;
        case 51: if (51 < RMAX) col[51] = 0.0; break;
        case 52: if (52 < RMAX) col[52] = 0.0; break;
        case 53: if (53 < RMAX) col[53] = 0.0; break;
        case 54: if (54 < RMAX) col[54] = 0.0; break;
        case 55: if (55 < RMAX) col[55] = 0.0; break;
        default: break;
    }
}

// order-preserving map double -> uint64 (total order of the IEEE values, negative numbers first)
__device__ __forceinline__ unsigned long long ordered_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
// warp argmin of (key, idx) with first-index tie-break: three REDUX.MIN instead of five shuffle rounds
__device__ __forceinline__ void warp_argmin(unsigned long long& key, int& idx) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const bool c1 = hi == mhi;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
    const bool c2 = c1 && lo == mlo;
    const unsigned mi = __reduce_min_sync(0xffffffffu, c2 ? (unsigned)idx : 0x7fffffffu);
    key = ((unsigned long long)mhi << 32) | mlo;
    idx = (int)mi;
}
constexpr unsigned long long KEY_NONE = 0xffffffffffffffffull;

template <int RMAX, bool EXACT>
__global__ void __launch_bounds__(CE_THREADS, 1) car_cols_kernel(const CarColsParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int P = CE_P;
    constexpr int VLEN = 8 * RMAX;                // padded column length (zeros beyond S)
    const int r = (int)cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int cs = t >> 3, rs = t & 7;            // column slot / row slice of this thread
    const int S = p.S, k = p.k;
    constexpr int SV = VLEN + 4;                  // v | alpha, v_j, j, pad

    extern __shared__ __align__(16) double sm[];
    double* vbuf = sm;                            // P x SV   broadcast buffers, one per owner rank
    double* mu_s = vbuf + (size_t)P * SV;         // VLEN     replicated weights
    __shared__ __align__(8) uint64_t bars[CE_P];
    __shared__ unsigned long long red_key[CE_THREADS / 32];
    __shared__ int red_idx[CE_THREADS / 32];

    if (t == 0) {
        for (int q = 0; q < P; ++q) mbar_init(&bars[q], 1);
        mbar_fence_init();
    }
    const int c = cs * P + r;                     // my column
    double col[RMAX];
#pragma unroll
    for (int li = 0; li < RMAX; ++li) {
        const int i = rs + 8 * li;
        col[li] = (c < k && i < S) ? p.basis[(size_t)c * S + i] : 0.0;
    }
    for (int i = t; i < VLEN; i += CE_THREADS) mu_s[i] = i < S ? p.mu[i] : 0.0;
    for (int i = t; i < P * SV; i += CE_THREADS) vbuf[i] = 0.0;
    __syncthreads();
    if (c == 0) {                                 // pivot column 0 -> my broadcast buffer
#pragma unroll
        for (int li = 0; li < RMAX; ++li) vbuf[rs + 8 * li] = col[li];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();

    uint32_t phase_bits = 0;                      // bit q = parity to wait for on bars[q]
    int done = 0;
    long long pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool prof = p.prof != nullptr && r == 0 && t == 0;
    long long pt = prof ? clock64() : 0;
#define CE_TICK(slot_)                                  \
    if (prof) {                                         \
        const long long now_ = clock64();               \
        pa[slot_] += now_ - pt;                         \
        pt = now_;                                      \
    }
    // Pivot search for column s by its owner CTA (all threads), then broadcast [v | alpha, v_j, j] to the 7 peers.
    // Column s is already in the owner's buffer.  <= 2 rows per thread; the two divisions are independent.
    auto search_and_broadcast = [&](int s) {
        const int o = s % P;
        double* vb = vbuf + (size_t)o * SV;
        unsigned long long key = KEY_NONE;
        int bi = 0x7fffffff;
        {
            const int i0 = t, i1 = t + CE_THREADS;
            const double v0 = vb[i0], v1 = i1 < VLEN ? vb[i1] : 0.0;
            const double m0 = mu_s[i0], m1 = i1 < VLEN ? mu_s[i1] : 0.0;
            double q0, q1;
            if (EXACT) {
                q0 = __ddiv_rn(m0, v0 > 0.0 ? v0 : 1.0);
                q1 = __ddiv_rn(m1, v1 > 0.0 ? v1 : 1.0);
            } else {
                q0 = div_pos(m0, v0 > 0.0 ? v0 : 1.0);
                q1 = div_pos(m1, v1 > 0.0 ? v1 : 1.0);
            }
            const unsigned long long k0 = v0 > 0.0 ? ordered_key(q0) : KEY_NONE;
            const unsigned long long k1 = v1 > 0.0 ? ordered_key(q1) : KEY_NONE;
            if (k0 != KEY_NONE) { key = k0; bi = i0; }
            if (k1 < key) { key = k1; bi = i1; }
        }
        warp_argmin(key, bi);
        if (lane == 0) { red_key[warp] = key; red_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            key = lane < CE_THREADS / 32 ? red_key[lane] : KEY_NONE;
            bi = lane < CE_THREADS / 32 ? red_idx[lane] : 0x7fffffff;
            warp_argmin(key, bi);
            if (lane == 0) {
                const bool any = key != KEY_NONE;
                vb[VLEN + 0] = any ? key_to_double(key) : 0.0;
                vb[VLEN + 1] = any ? vb[bi] : 1.0;
                vb[VLEN + 2] = any ? (double)bi : -1.0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            __syncwarp();
            if (lane < P - 1) {
                const int peer = lane < r ? lane : lane + 1;
                ce_bulk_s2c(ce_map_rank(smem_addr(vb), peer), smem_addr(vb), (uint32_t)SV * 8u,
                            ce_map_rank(smem_addr(&bars[o]), peer));
            }
        }
        __syncthreads();   // the owner's own threads see the scalars
    };

    if (r == 0) search_and_broadcast(0);
    for (int s = 0; s < k; ++s) {
        const int o = s % P;                      // owner rank of pivot column s, also the buffer index
        double* vb = vbuf + (size_t)o * SV;
        if (r != o) {
            if (t == 0) mbar_expect_tx(&bars[o], (uint32_t)SV * 8u);
            mbar_wait(&bars[o], (phase_bits >> o) & 1u);
            phase_bits ^= (1u << o);
        }
        CE_TICK(3)
        const double alpha = vb[VLEN + 0];
        const double vj = vb[VLEN + 1];
        const int j = (int)vb[VLEN + 2];
        if (j < 0) break;   // no positive entry: the guard of SOBER/_rchq.py:241-242 (uniform over the cluster)
        done = s + 1;
        for (int i = t; i < S; i += CE_THREADS)
            mu_s[i] = (i == j) ? 0.0 : __dsub_rn(mu_s[i], __dmul_rn(alpha, vb[i]));
        // pivot-row entry of my column: register li = j / 8 of the lane with rs == j % 8 in my 8-lane group
        const double mine = ce_col_get<RMAX>(col, j >> 3);
        const double pj = __shfl_sync(0xffffffffu, mine, (lane & ~7) | (j & 7));
        const double nf = EXACT ? 0.0 : -div_pos(fabs(pj), vj) * (pj < 0.0 ? -1.0 : 1.0);
        auto update_my_column = [&]() {
            if (EXACT) {
#pragma unroll
                for (int li = 0; li < RMAX; ++li)
                    col[li] = __dsub_rn(col[li], __ddiv_rn(__dmul_rn(pj, vb[rs + 8 * li]), vj));
            } else {
#pragma unroll
                for (int li = 0; li < RMAX; ++li) col[li] = fma(nf, vb[rs + 8 * li], col[li]);
            }
            if (rs == (j & 7)) ce_col_zero<RMAX>(col, j >> 3);
        };
        // software pipeline: the next pivot column first, its owner searches and broadcasts, THEN everybody applies
        // this step's update to the rest of their columns (off the critical path: it overlaps the broadcast)
        if (c == s + 1 && c < k) {
            update_my_column();
            double* nb = vbuf + (size_t)((s + 1) % P) * SV;
#pragma unroll
            for (int li = 0; li < RMAX; ++li) nb[rs + 8 * li] = col[li];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();   // mu_s and (in its owner) the next pivot column are complete
        CE_TICK(4)
        if (s + 1 < k && r == (s + 1) % P) search_and_broadcast(s + 1);
        CE_TICK(0)
        if (c > s + 1 && c < k) update_my_column();
        CE_TICK(5)
    }
    if (prof)
        for (int i = 0; i < 8; ++i) p.prof[i] = pa[i];
    __syncthreads();
    if (r == 0) {
        for (int i = t; i < S; i += CE_THREADS) p.mu[i] = mu_s[i];
        if (t == 0 && p.info) p.info[0] = done;
    }
    cluster.sync();
}

static int cols_rmax(int S, int k) {
    if (k > CE_P * (CE_THREADS / 8) || k <= 0) return 0;
    const int rows = (S + 7) / 8;
    return rows <= 32 ? 32 : (rows <= 56 ? 56 : 0);
}

}  // namespace sober

using namespace sober;

extern "C" int sober_car_cluster_cols_fits(int32_t S, int32_t k) { return (S > 0 && cols_rmax(S, k)) ? CE_P : 0; }

extern "C" int sober_car_cluster_cols_profiled(double* basis, int32_t k, int32_t S, double* mu, int32_t exact,
                                               int32_t* info, int64_t* prof, void* stream);
extern "C" int sober_car_cluster_cols(double* basis, int32_t k, int32_t S, double* mu, int32_t exact, int32_t* info,
                                      void* stream) {
    return sober_car_cluster_cols_profiled(basis, k, S, mu, exact, info, nullptr, stream);
}
extern "C" int sober_car_cluster_cols_profiled(double* basis, int32_t k, int32_t S, double* mu, int32_t exact,
                                               int32_t* info, int64_t* prof, void* stream) {
    if (S <= 0 || k <= 0 || !basis || !mu) return SOBER_ERR_ARG;
    const int rmax = cols_rmax(S, k);
    if (rmax == 0) return SOBER_ERR_UNSUPPORTED;
    CarColsParams p;
    p.basis = basis; p.mu = mu; p.info = info; p.prof = (long long*)prof; p.S = S; p.k = k;
    p.Spad = 8 * rmax;
    p.SV = p.Spad + 4;
    const size_t smem = ((size_t)CE_P * p.SV + p.Spad) * 8 + 64;
    void (*kern)(const CarColsParams) = nullptr;
    if (rmax == 32) kern = exact ? car_cols_kernel<32, true> : car_cols_kernel<32, false>;
    else kern = exact ? car_cols_kernel<56, true> : car_cols_kernel<56, false>;
    {   // the > 48 KB opt-in is per device and sticky: once per (device, kernel variant), never inside a graph capture
        static int configured[64][4] = {};
        int dev = 0;
        SOBER_CUDA_CHECK(cudaGetDevice(&dev));
        const int variant = (rmax == 32 ? 0 : 2) + (exact ? 1 : 0);
        if (dev < 0 || dev >= 64 || configured[dev][variant] < (int)smem) {
            SOBER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) configured[dev][variant] = (int)smem;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CE_P);
    cfg.blockDim = dim3(CE_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CE_P;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SOBER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    return SOBER_OK;
}
