// Small dense helpers for the N-independent steps (Cholesky-QR of the Nystrom range finder and of the projector
// null space).  cuBLAS/cuSOLVER spend ~0.1 ms per call on these q ~ 200 problems (single-CTA, latency-bound
// kernels); they are called ~40 times per recombination.  (A one-CTA packed-triangle Cholesky was also tried here:
// 0.32 ms per call at q = 200 against 0.11 ms for cuSOLVER's potrf -- three block barriers per column on a 1024-thread
// CTA -- and was dropped; the factorisation stays with torch.linalg.cholesky_ex.)
#include "common.cuh"

namespace sober {

// X * R = Y  for upper-triangular R (q x q, row-major, q <= 256): one WARP per TR_ROWS rows of Y.
// Lane l holds elements k = 32 r + l of its rows in registers y[.][r]; step j broadcasts x_j = y_j / R_jj by shuffle
// and applies y_k -= x_j R_jk to the elements right of j: the dependent chain is shuffle -> mul -> fma, two
// independent chains per warp, and the row of R for step j+1 is fetched into registers during step j.
// R streams through shared memory 16 rows at a time: 16 rows of a row-major matrix are ONE contiguous, 16-byte
// aligned range, so a tile is a single TMA bulk copy (cp.async.bulk + mbarrier) issued by one thread, four tiles in
// flight.  History of this kernel at q = 200: R read from global memory every step 86 us (an L2 round trip per
// step, no better than cuBLAS trsm); tiles loaded synchronously 44 us; fully unrolled steps -> instruction-fetch
// bound; 8-byte cp.async tiles 29 us, a third of it spent issuing the copies; this version: see DESIGN.md.
constexpr int TR_TILE = 16;
constexpr int TR_ROWS = 2;
constexpr int TR_STAGES = 4;
constexpr int TR_WARPS = 4;     // one per SM sub-partition: m / 8 CTAs, nothing to contend with the dependent chain

template <int NR>
__global__ void __launch_bounds__(32 * TR_WARPS) trsm_right_upper_kernel(const double* __restrict__ Y, int64_t ldy,
                                                                         const double* __restrict__ R, int ldr, int m,
                                                                         int q, double* __restrict__ X, int64_t ldx,
                                                                         int stage_elems) {
    constexpr int W = 32 * NR;
    extern __shared__ __align__(16) double tr_smem[];
    double* rt = tr_smem;                                       // TR_STAGES x stage_elems (+ slack for over-reads)
    double* rinv = rt + TR_STAGES * stage_elems + 2 * W;        // W
    __shared__ __align__(8) uint64_t bars[TR_STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntiles = (q + TR_TILE - 1) / TR_TILE;
    // tile t = rows 16 t .. 16 t + 15 of R with their full stride.  The byte count is rounded DOWN to 16: what can be
    // cut off is the last element of the last row -- the diagonal R[q-1][q-1] (ldr == q) or padding (ldr > q), neither
    // of which is read from the tile.
    auto issue = [&](int t) {
        if (t < ntiles) {
            const int j0 = t * TR_TILE;
            const uint32_t bytes = (uint32_t)((min(TR_TILE, q - j0) * ldr) & ~1) * 8u;
            uint64_t* bar = &bars[t % TR_STAGES];
            mbar_expect_tx(bar, bytes);
            bulk_g2s(rt + (t % TR_STAGES) * stage_elems, R + (size_t)j0 * ldr, bytes, bar);
        }
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < TR_STAGES; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    for (int j = threadIdx.x; j < W; j += blockDim.x) rinv[j] = j < q ? 1.0 / R[(size_t)j * ldr + j] : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int t = 0; t < TR_STAGES; ++t) issue(t);
    }
    const int row0 = (blockIdx.x * TR_WARPS + warp) * TR_ROWS;
    double y[TR_ROWS][NR];
#pragma unroll
    for (int p = 0; p < TR_ROWS; ++p)
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int k = 32 * r + lane;
            y[p][r] = (row0 + p < m && k < q) ? Y[(int64_t)(row0 + p) * ldy + k] : 0.0;
        }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
#pragma unroll 1
        for (int half = 0; half < 32 / TR_TILE; ++half) {   // rolled: straight-line code for all q steps does not fit
            const int t = r * (32 / TR_TILE) + half;        // the instruction cache (measured: fetch-bound)
            const int j0 = t * TR_TILE;
            if (j0 < q) {                                   // uniform over the CTA
                mbar_wait(&bars[t % TR_STAGES], (uint32_t)(t / TR_STAGES) & 1u);
                const double* tile = rt + (t % TR_STAGES) * stage_elems;
                // Elements k >= q of a register (lanes beyond the matrix) pick up whatever follows in shared memory;
                // they are never shuffled from nor stored.
                double rv[NR], rn[NR];
#pragma unroll
                for (int rr = r; rr < NR; ++rr) rv[rr] = tile[32 * rr + lane];
                double ri = rinv[j0];
                const int steps = min(TR_TILE, q - j0);     // rows beyond q were not loaded
#pragma unroll 2
                for (int jr = 0; jr < steps; ++jr) {
                    const int jj = TR_TILE * half + jr;     // lane that holds element j0 + jr of register r
                    double xj[TR_ROWS];
#pragma unroll
                    for (int p = 0; p < TR_ROWS; ++p) xj[p] = __shfl_sync(0xffffffffu, y[p][r], jj) * ri;
                    const int jn = jr + 1 < steps ? jr + 1 : jr;
#pragma unroll
                    for (int rr = r; rr < NR; ++rr) rn[rr] = tile[jn * ldr + 32 * rr + lane];
                    ri = rinv[j0 + jn];
                    const double r0 = lane > jj ? rv[r] : 0.0;                  // strictly right of the diagonal
#pragma unroll
                    for (int p = 0; p < TR_ROWS; ++p)
                        y[p][r] = lane == jj ? xj[p] : fma(-xj[p], r0, y[p][r]);
#pragma unroll
                    for (int rr = r + 1; rr < NR; ++rr)
#pragma unroll
                        for (int p = 0; p < TR_ROWS; ++p) y[p][rr] = fma(-xj[p], rv[rr], y[p][rr]);
#pragma unroll
                    for (int rr = r; rr < NR; ++rr) rv[rr] = rn[rr];
                }
                __syncthreads();                            // tile t consumed by all warps: its buffer is free
                if (threadIdx.x == 0) issue(t + TR_STAGES);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < TR_ROWS; ++p)
        if (row0 + p < m) {
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                const int k = 32 * r + lane;
                if (k < q) X[(int64_t)(row0 + p) * ldx + k] = y[p][r];
            }
        }
}

}  // namespace sober

using namespace sober;

extern "C" int sober_trsm_right_upper(const double* Y, int64_t ldy, const double* R, int64_t ldr, int32_t m, int32_t q,
                                      double* X, int64_t ldx, void* stream) {
    if (m < 0 || q <= 0 || q > 256 || !Y || !R || !X || ldy < q || ldr < q || ldx < q) return SOBER_ERR_ARG;
    if (m == 0) return SOBER_OK;
    // the tiles are TMA bulk copies: 16-byte aligned base, and four 16-row tiles must fit the shared memory
    if ((reinterpret_cast<uintptr_t>(R) & 15u) != 0 || ldr > 384) return SOBER_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int nr = (q + 31) / 32;
    const dim3 grid((unsigned)ceil_div(m, TR_WARPS * TR_ROWS));
    const int stage_elems = (int)((TR_TILE * ldr + 1) & ~(int64_t)1);
    const size_t smem = ((size_t)TR_STAGES * stage_elems + 3 * 32 * nr) * sizeof(double);
    // the opt-in for > 48 KB of dynamic shared memory is per device and sticky; its size depends on ldr
    static int configured[64][9] = {};
    int dev = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    const bool need_attr = dev < 0 || dev >= 64 || configured[dev][nr] < (int)smem;
#define TR_LAUNCH(NR_)                                                                                                \
    case NR_:                                                                                                         \
        if (need_attr)                                                                                                \
            SOBER_CUDA_CHECK(cudaFuncSetAttribute(trsm_right_upper_kernel<NR_>,                                       \
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        trsm_right_upper_kernel<NR_><<<grid, 32 * TR_WARPS, smem, st>>>(Y, ldy, R, (int)ldr, m, q, X, ldx, stage_elems); \
        break;
    switch (nr) {
        TR_LAUNCH(1) TR_LAUNCH(2) TR_LAUNCH(3) TR_LAUNCH(4) TR_LAUNCH(5) TR_LAUNCH(6) TR_LAUNCH(7) TR_LAUNCH(8)
        default: return SOBER_ERR_ARG;
    }
#undef TR_LAUNCH
    if (dev >= 0 && dev < 64 && configured[dev][nr] < (int)smem) configured[dev][nr] = (int)smem;
    SOBER_LAUNCH_CHECK("trsm_right_upper");
    return SOBER_OK;
}
