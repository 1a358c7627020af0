// q x q Cholesky factorisation G = R^T R (R upper triangular, q <= 224) on a 2-CTA thread-block cluster.
//
// Used by the Cholesky-QR passes of the Nystrom range finder and by the projector null space of every Caratheodory
// call (sober_b200/_nystrom.py, _car.py): 15-20 factorisations of a 200 x 200 Gram matrix per recombination, which
// cuSOLVER's potrf does in ~0.13 ms each (1300 cycles per column).  A column-by-column factorisation is a dependency
// chain -- update column j+1 with column j, square root, scale, publish -- and this kernel is organised so that the chain
// is ~300 cycles per column:
//   * the lower triangle lives in REGISTERS.  Warp w of a CTA owns the columns c = w (mod 16), lane a the rows
//     i = a (mod 32); the 32-column blocks alternate between the two CTAs of the cluster, which halves the register
//     tile (<= 32 doubles per thread) and the FP64 work per SM.
//   * a finished column is written to a shared-memory ring (and to row j of R in global memory), a per-column
//     shared-memory flag releases it to the warps of its own CTA, and ONE bulk DSMEM copy
//     (cp.async.bulk.shared::cluster.shared::cta, complete_tx on the peer's mbarrier for that column) hands it to the
//     other CTA.  No block- or cluster-wide barrier inside the loop.
//   * look-ahead: the warp that owns column j+1 applies column j to it FIRST, factors and publishes it, and only then
//     updates its other columns; every other warp lags behind the chain by at most 48 columns (its next own column is
//     at most 48 after its previous one), which the 64-slot ring covers.
// The chain crosses from one CTA to the other once per 32 columns.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;
using namespace sober;

namespace {

constexpr int CP_NCH = 7;                  // 32-row chunks
constexpr int CP_NMAX = 32 * CP_NCH;       // 224
constexpr int CP_THREADS = 512;            // 16 warps, two column residues (w, w + 16) each
constexpr int CP_RING = 64;
#ifndef CP_BACKOFF
#define CP_BACKOFF 128
#endif
constexpr int CP_KMAX = (CP_NCH + 1) / 2;  // column blocks per CTA

#ifdef CP_PROF
__device__ long long cp_prof[8];
__shared__ long long cp_tpub;
#define CP_T(var_) const long long var_ = clock64();
#else
#define CP_T(var_)
#endif

struct CholParams {
    const double* G;
    long long ldg;
    double* R;
    long long ldr;
    int* info;
    int n;
};

__device__ __forceinline__ void cp_flag_set(int* flag) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], 1;" ::"r"(smem_addr(flag)) : "memory");
}
// ``urgent``: the warp that factors one of the next two columns spins on the flag; every other warp backs off between
// polls (16 warps spinning on LDS take more than half of the shared-memory issue slots away from the chain).
__device__ __forceinline__ void cp_flag_wait(const int* flag, bool urgent) {
    int v;
    for (;;) {
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_addr(flag)) : "memory");
        if (v != 0) break;
        if (!urgent) {
            const long long t0 = clock64();
            while (clock64() - t0 < CP_BACKOFF) {}
        }
    }
}
__device__ __forceinline__ uint32_t cp_map_rank(uint32_t saddr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(saddr), "r"(rank));
    return out;
}
__device__ __forceinline__ void cp_bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}

// Register tile of one thread: m[e][k][r] = A[a + 32 r][w + 16 e + 32 (2 k + Q)], used only for r >= 2 k + Q.
template <int Q>
struct Tile {
    double m[2][CP_KMAX][CP_NCH];
};

// Square root + scaling of pivot column c (already fully updated) held in slot (E, K); publishes it.
template <int Q, int E, int K>
__device__ __forceinline__ void factor_column(Tile<Q>& t, const CholParams& p, double* ring, uint64_t* bars, int* flags, int c,
                                              int lane) {
    constexpr int B = 2 * K + Q;
    const int rho = c & 31;
    CP_T(t0)
    const double x = t.m[E][K][B];
    double piv = __shfl_sync(0xffffffffu, x, rho);
    if (!(piv > 0.0) || piv > 1.7e308) {
        if (lane == 0) atomicCAS(p.info, 0, c + 1);
        piv = __longlong_as_double(0x7ff8000000000000ll);
    }
    const double y = rsqrt(piv);
    CP_T(t1)
    double* dst = ring + (size_t)(c % CP_RING) * CP_NMAX;
    double v[CP_NCH];
#pragma unroll
    for (int r = 0; r < CP_NCH; ++r) {
        v[r] = 0.0;
        if (r == B) v[r] = lane < rho ? 0.0 : (lane == rho ? piv * y : x * y);
        if (r > B) v[r] = t.m[E][K][r] * y;
        if (r >= B) dst[32 * r + lane] = v[r];
    }
    // publish to this CTA first (the next pivot column is normally here), then to the peer and to global memory
    __syncwarp();
    if (lane == 0) cp_flag_set(&flags[c]);
#ifdef CP_PROF
    const long long t2 = clock64();
    if (lane == 0) {
        cp_tpub = t2;
        atomicAdd((unsigned long long*)&cp_prof[2], (unsigned long long)(t1 - t0));
        atomicAdd((unsigned long long*)&cp_prof[3], (unsigned long long)(t2 - t1));
    }
#endif
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        const uint32_t src = smem_addr(dst + 32 * B);
        cp_bulk_s2c(cp_map_rank(src, Q ^ 1), src, (uint32_t)(CP_NCH - B) * 256u, cp_map_rank(smem_addr(&bars[c]), Q ^ 1));
    }
    double* out = p.R + (size_t)c * p.ldr;
#pragma unroll
    for (int r = 0; r < CP_NCH; ++r)
        if (32 * r + lane < p.n) out[32 * r + lane] = v[r];
}

template <int Q, int E, int K>
__device__ __forceinline__ void update_slot(Tile<Q>& t, const double (&l)[CP_NCH], double lc) {
    constexpr int B = 2 * K + Q;
#pragma unroll
    for (int r = B; r < CP_NCH; ++r) t.m[E][K][r] = fma(-l[r], lc, t.m[E][K][r]);
}

// All steps j of the 32-column block BJ (static, so that every register-tile index is a compile-time constant).
template <int Q, int BJ>
__device__ __forceinline__ void block_steps(Tile<Q>& t, const CholParams& p, double* ring, uint64_t* bars, int* flags,
                                            int lane, int w) {
    const int n = p.n;
    if (32 * BJ >= n) return;
    const int jend = min(32, n - 32 * BJ);
    for (int jj = 0; jj < jend; ++jj) {
        const int j = 32 * BJ + jj;
        if constexpr ((BJ & 1) == Q)
            cp_flag_wait(&flags[j], ((w - jj - 1) & 15) < 2);   // a column of this CTA: plain shared-memory flag
        else
            mbar_wait(&bars[j], 0);           // a column of the peer: completed by its bulk copy
        const double* col = ring + (size_t)(j % CP_RING) * CP_NMAX;
        double l[CP_NCH];
#pragma unroll
        for (int r = 0; r < CP_NCH; ++r) l[r] = r >= BJ ? col[32 * r + lane] : 0.0;

        // look-ahead: the owner of column j + 1 finishes and publishes it before anything else
        int skip_e = -1, skip_b = -1;
        if (j + 1 < n) {
            if (jj < 31) {
                if constexpr ((BJ & 1) == Q) {
                    if (w == ((jj + 1) & 15)) {
                        constexpr int K = BJ / 2;
#ifdef CP_PROF
                        const long long tw = clock64();
                        if (lane == 0 && jj > 0) {
                            atomicAdd((unsigned long long*)&cp_prof[0], (unsigned long long)(tw - cp_tpub));
                            atomicAdd((unsigned long long*)&cp_prof[4], 1ull);
                        }
#endif
                        const double lc = col[j + 1];
#ifdef CP_PROF
                        {
                            double sink = l[BJ] * lc;
                            if (sink == 1.2345e300) cp_prof[7] = 1;
                            const long long tl = clock64();
                            if (lane == 0) atomicAdd((unsigned long long*)&cp_prof[1], (unsigned long long)(tl - tw));
                        }
#endif
                        if (jj + 1 < 16) {
                            update_slot<Q, 0, K>(t, l, lc);
                            factor_column<Q, 0, K>(t, p, ring, bars, flags, j + 1, lane);
                            skip_e = 0;
                        } else {
                            update_slot<Q, 1, K>(t, l, lc);
                            factor_column<Q, 1, K>(t, p, ring, bars, flags, j + 1, lane);
                            skip_e = 1;
                        }
                        skip_b = BJ;
                    }
                }
            } else {
                if constexpr (((BJ + 1) & 1) == Q && BJ + 1 < CP_NCH) {
                    if (w == 0) {
                        constexpr int K = (BJ + 1) / 2;
                        const double lc = col[j + 1];
                        update_slot<Q, 0, K>(t, l, lc);
                        factor_column<Q, 0, K>(t, p, ring, bars, flags, j + 1, lane);
                        skip_e = 0;
                        skip_b = BJ + 1;
                    }
                }
            }
        }
        // every other column of this warp to the right of j
#define CP_UPDATE(E_, K_)                                                          \
    if constexpr (2 * (K_) + Q < CP_NCH && 2 * (K_) + Q >= BJ) {                   \
        const int c = w + 16 * (E_) + 32 * (2 * (K_) + Q);                         \
        if (c > j && c < n && !((E_) == skip_e && 2 * (K_) + Q == skip_b))         \
            update_slot<Q, E_, K_>(t, l, col[c]);                                  \
    }
        CP_UPDATE(0, 0) CP_UPDATE(1, 0) CP_UPDATE(0, 1) CP_UPDATE(1, 1)
        CP_UPDATE(0, 2) CP_UPDATE(1, 2) CP_UPDATE(0, 3) CP_UPDATE(1, 3)
#undef CP_UPDATE
    }
}

template <int Q>
__device__ __forceinline__ void chol_body(const CholParams& p, double* ring, uint64_t* bars, int* flags) {
    const int n = p.n;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Tile<Q> t;
#pragma unroll
    for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int k = 0; k < CP_KMAX; ++k) {
            const int B = 2 * k + Q;
            const int c = w + 16 * e + 32 * B;
#pragma unroll
            for (int r = 0; r < CP_NCH; ++r) {
                const int i = lane + 32 * r;
                t.m[e][k][r] = (B < CP_NCH && r >= B && c < n && i < n && i >= c) ? p.G[(size_t)c * p.ldg + i] : 0.0;
            }
        }

    if (Q == 0 && w == 0) factor_column<Q, 0, 0>(t, p, ring, bars, flags, 0, lane);
    block_steps<Q, 0>(t, p, ring, bars, flags, lane, w);
    block_steps<Q, 1>(t, p, ring, bars, flags, lane, w);
    block_steps<Q, 2>(t, p, ring, bars, flags, lane, w);
    block_steps<Q, 3>(t, p, ring, bars, flags, lane, w);
    block_steps<Q, 4>(t, p, ring, bars, flags, lane, w);
    block_steps<Q, 5>(t, p, ring, bars, flags, lane, w);
    block_steps<Q, 6>(t, p, ring, bars, flags, lane, w);
    static_assert(CP_NCH == 7, "one block_steps call per 32-column block");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CP_THREADS, 1) chol_pair_kernel(const CholParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double sm[];
    double* ring = sm;                                              // CP_RING x CP_NMAX
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)CP_RING * CP_NMAX);   // one per column
    int* flags = reinterpret_cast<int*>(bars + CP_NMAX);           // one per column (columns of this CTA)
    const int q = (int)cluster.block_rank();
    const int t = threadIdx.x;
    if (t < CP_NMAX) {
        const int B = t >> 5;
        const bool mine = (B & 1) == q;
        flags[t] = 0;
        mbar_init(&bars[t], 1);
        mbar_fence_init();
        if (!mine && t < p.n) mbar_expect_tx(&bars[t], (uint32_t)(CP_NCH - B) * 256u);
    }
    if (q == 0 && t == 0) *p.info = 0;
    __syncthreads();
    cluster.sync();
    if (q == 0)
        chol_body<0>(p, ring, bars, flags);
    else
        chol_body<1>(p, ring, bars, flags);
    cluster.sync();   // no CTA leaves while the peer can still read from / copy into its shared memory
}

}  // namespace

#ifdef CP_PROF
extern "C" int sober_cholesky_prof(long long* out, int reset) {
    long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (reset) return cudaMemcpyToSymbol(cp_prof, z, sizeof(z)) == cudaSuccess ? 0 : 2;
    return cudaMemcpyFromSymbol(out, cp_prof, sizeof(z)) == cudaSuccess ? 0 : 2;
}
#endif

extern "C" int sober_cholesky_upper_fits(int32_t n) { return n > 0 && n <= CP_NMAX; }

extern "C" int sober_cholesky_upper(const double* G, int64_t ldg, int32_t n, double* R, int64_t ldr, int32_t* info,
                                    void* stream) {
    if (!G || !R || !info || n <= 0 || n > CP_NMAX || ldg < n || ldr < n) return SOBER_ERR_ARG;
    const size_t smem = (size_t)CP_RING * CP_NMAX * sizeof(double) + (size_t)CP_NMAX * (sizeof(uint64_t) + sizeof(int));
    static bool configured[64] = {};   // the > 48 KB opt-in is per device and sticky
    int dev = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        SOBER_CUDA_CHECK(cudaFuncSetAttribute(chol_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    CholParams p{G, (long long)ldg, R, (long long)ldr, info, n};
    chol_pair_kernel<<<2, CP_THREADS, smem, (cudaStream_t)stream>>>(p);
    SOBER_LAUNCH_CHECK("cholesky_upper");
    return SOBER_OK;
}
