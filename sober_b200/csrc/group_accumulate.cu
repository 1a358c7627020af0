// K1 -- fused cross-kernel + weighted strided group sums.
//
// Replaces SOBER/_rchq.py:116-136,152 (and :35 / :78 when used as a plain Gram): the reference materialises
// K = kernel(pt_nys, samp[idx]) as an (E, L, S) tensor, multiplies by mu and sums over E.  Here every
// candidate is streamed once per pass, its L kernel values are formed in registers and folded straight into
// the (S x L) accumulator; nothing of size N x L ever exists.
//
// Two kernels share one contract:
//   * group_records : d <= 8, record layout.  A CTA owns TG adjacent groups and up to 8*32*TL landmarks (one lane
//                     = TL landmarks held in registers; TL = 2, TG = 4 measured best of 8 variants, tools/k1_variants.sh).  The candidate records of its rows are staged into
//                     shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier, 3 stages) issued by one
//                     thread; every warp then reads each record as a broadcast LDS and evaluates TL x TG kernel
//                     values.  FP64-pipe bound: ~28 FP64 instructions per Matern-5/2 evaluation (common.cuh).
//   * group_tiled   : any d, indexed layout.  64 landmarks x 64 groups per CTA, the <x, z> contraction runs over
//                     d in chunks of 16 staged through shared memory, 4x4 register tile per thread.
// Rows are split across gridDim.z; partial sums go to a workspace and are reduced in a fixed order
// (deterministic: no atomics), which also applies the output scale.
#include "common.cuh"

namespace sober {

// csrc/group_bits_mma.cu: Tanimoto on bit-packed rows through tcgen05 (kind::i8)
struct BitsMmaParams {
    const uint64_t* X;
    const uint64_t* Z;
    const double* xn;
    int64_t xn_stride;
    const double* zn;
    const int32_t* idx;
    const double* mu;
    int64_t n_local, pos0, ES;
    int S, L, W;
    double* out;
    double* totw_out;
    int64_t row_begin, row_end, rows_per_split;
    double scale;
};
bool bits_mma_supported(int W);
int launch_bits_mma(const BitsMmaParams& p, dim3 grid, cudaStream_t st);    // v1: landmark tile resident in shared memory
int launch_bits_mma2(const BitsMmaParams& p, dim3 grid, cudaStream_t st);   // v2: landmark tile resident in TMEM (A operand)

struct GroupParams {
    const double* X;
    int64_t ldx;
    const double* xn;
    int64_t xn_stride;
    const int32_t* idx;
    const double* mu;
    const double* rec;
    int64_t n_local, pos0, ES;
    int S, L, d;
    int unit_weights;
    const double* Zt;
    const double* zn;
    const double* lut;
    const double* kx;  // POST: (N x n_obs) candidate-side rows k(x_i, X_obs), addressed by row id like X
    int64_t ldkx;
    const double* aw;  // POST: (L x n_obs) landmark-side rows k(z_l, X_obs) W
    int n_obs;
    double inner_scale;
    double* out;       // [nsplit][S][L]  (or At itself when nsplit == 1)
    double* totw_out;  // [nsplit][S]
    int64_t row_begin, row_end, rows_per_split;
    double scale;      // applied at store time when nsplit == 1, else by the reduce kernel
};

// -------------------------------------------------------------------------------------------------
// record kernel (small d)
// -------------------------------------------------------------------------------------------------
constexpr int REC_WARPS = 8;
constexpr int REC_THREADS = REC_WARPS * 32;
#ifndef SOBER_REC_ROWS
#define SOBER_REC_ROWS 16
#endif
constexpr int REC_ROWS = SOBER_REC_ROWS;    // rows per pipeline stage
constexpr int REC_STAGES = 3;

#ifndef SOBER_REC_MINB
#define SOBER_REC_MINB 2
#endif
// (A variant without the per-chunk block barrier -- per-stage "empty" mbarriers, 4 stages -- was measured in round 2:
// bitwise-equal output, 2-6 % SLOWER at every shape (profiles/r02_k1_variants.txt), and removed.)
template <int D, int FAM, int TL, int TG, bool UNIT>
__global__ void __launch_bounds__(REC_THREADS, SOBER_REC_MINB) group_records_kernel(const GroupParams p) {
    constexpr int LDR = (D + 3) / 2 * 2;  // d + 2 rounded up to even
    __shared__ __align__(16) double buf[REC_STAGES][REC_ROWS][TG][LDR];
    __shared__ __align__(8) uint64_t bars[REC_STAGES];
    __shared__ double tab[EXP_TAB_SIZE];

    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int g0 = blockIdx.x * TG;
    const int l0 = (blockIdx.y * REC_WARPS + warp) * (32 * TL);
    const bool active = l0 < p.L;

    load_exp_table(tab, t, REC_THREADS);
    uint32_t tab_s = smem_addr(tab);
    asm volatile("" : "+r"(tab_s));   // opaque: otherwise the address is re-derived (S2UR + ULEA + ...) at every use
    if (t == 0) {
        for (int s = 0; s < REC_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t r0 = p.row_begin + (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t r1 = min(p.row_end, r0 + p.rows_per_split);
    const int64_t hi = p.pos0 + p.n_local;
    const int nchunks = (int)((r1 - r0 + REC_ROWS - 1) / REC_ROWS);
    const int jmax = min(TG, p.S - g0);   // groups of this CTA that exist

    // valid group interval [ja, jb) of row e: positions e*S + g0 + j that this device owns
    auto interval = [&](int64_t e, int& ja, int& jb) {
        const int64_t base = e * p.S + g0;
        const int64_t a = max((int64_t)0, p.pos0 - base);
        const int64_t b = min((int64_t)jmax, hi - base);
        ja = (int)min(a, (int64_t)TG);
        jb = (int)max(b, (int64_t)ja);
    };
    auto issue = [&](int c) {   // one thread: TMA bulk copies of chunk c into its stage
        const int stage = c % REC_STAGES;
        const int64_t e0 = r0 + (int64_t)c * REC_ROWS;
        const int nrows = (int)min((int64_t)REC_ROWS, r1 - e0);
        uint32_t bytes = 0;
        for (int r = 0; r < nrows; ++r) {
            int ja, jb;
            interval(e0 + r, ja, jb);
            bytes += (uint32_t)(jb - ja) * LDR * 8;
        }
        mbar_expect_tx(&bars[stage], bytes);
        for (int r = 0; r < nrows; ++r) {
            int ja, jb;
            interval(e0 + r, ja, jb);
            if (jb > ja) {
                const int64_t loc = (e0 + r) * p.S + g0 + ja - p.pos0;
                bulk_g2s(&buf[stage][r][ja][0], p.rec + loc * LDR, (uint32_t)(jb - ja) * LDR * 8, &bars[stage]);
            }
        }
    };
    if (t == 0)
        for (int c = 0; c < REC_STAGES && c < nchunks; ++c) issue(c);

    double zt[TL][D], zn[TL];
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int l = l0 + lane + 32 * i;
        const bool ok = l < p.L;
#pragma unroll
        for (int k = 0; k < D; ++k) zt[i][k] = ok ? __ldg(p.Zt + (int64_t)l * D + k) : 0.0;
        zn[i] = ok ? __ldg(p.zn + l) : 0.0;
    }
    double acc[TL][TG], tw[TG];
#pragma unroll
    for (int j = 0; j < TG; ++j) {
        tw[j] = 0.0;
#pragma unroll
        for (int i = 0; i < TL; ++i) acc[i][j] = 0.0;
    }
    // group weight totals: warp 0 of the CTAs with blockIdx.y == 0, one staged row per lane (round 1 had thread 0 walk
    // all rows of a chunk alone while its warp waited -- and the whole CTA waited for that warp at the chunk barrier)
    const bool tw_warp = (blockIdx.y == 0) && (warp == 0);
    static_assert(REC_ROWS <= 32, "one staged row per lane");

    // all TG candidates of staged row r: TL x TG independent chains
    auto full_row = [&](int stage, int r) {
        double x[TG][D], xn[TG], w[TG];
#pragma unroll
        for (int j = 0; j < TG; ++j) {
            const double* rp = &buf[stage][r][j][0];
#pragma unroll
            for (int k = 0; k < D; ++k) x[j][k] = rp[k];
            xn[j] = rp[D];
            w[j] = UNIT ? 1.0 : rp[D + 1];
        }
        double val[TL][TG];
#pragma unroll
        for (int i = 0; i < TL; ++i)
#pragma unroll
            for (int j = 0; j < TG; ++j) {
                double dot = (FAM == SOBER_TANIMOTO) ? 0.0 : zn[i];
#pragma unroll
                for (int k = 0; k < D; ++k) dot = fma(x[j][k], zt[i][k], dot);
                val[i][j] = (FAM == SOBER_TANIMOTO) ? tanimoto_value(dot, xn[j], zn[i])
                                                    : stationary_value<FAM>(xn[j] + dot, tab_s);
            }
#pragma unroll
        for (int i = 0; i < TL; ++i)
#pragma unroll
            for (int j = 0; j < TG; ++j) acc[i][j] = fma(val[i][j], w[j], acc[i][j]);
    };

    for (int c = 0; c < nchunks; ++c) {
        const int stage = c % REC_STAGES;
        mbar_wait(&bars[stage], (uint32_t)((c / REC_STAGES) & 1));
        const int64_t e0 = r0 + (int64_t)c * REC_ROWS;
        const int nrows = (int)min((int64_t)REC_ROWS, r1 - e0);
        // interior chunk: every row has all TG candidates on this device (the usual case) -- no per-row bookkeeping
        const int64_t base0 = e0 * p.S + g0;
        const bool full = jmax == TG && base0 >= p.pos0 && base0 + (int64_t)(nrows - 1) * p.S + TG <= hi;
        if (tw_warp && lane < nrows) {
            int ja = 0, jb = TG;
            if (!full) interval(e0 + lane, ja, jb);
#pragma unroll
            for (int j = 0; j < TG; ++j)
                if (j >= ja && j < jb && (e0 + lane) * p.S + g0 + j < p.ES)
                    tw[j] += UNIT ? 1.0 : buf[stage][lane][j][D + 1];
        }
        if (active) {
            if (full) {
                for (int r = 0; r < nrows; ++r) full_row(stage, r);
            } else {
                for (int r = 0; r < nrows; ++r) {
                    int ja, jb;
                    interval(e0 + r, ja, jb);
                    if (ja == 0 && jb == TG) {
                        full_row(stage, r);
                    } else {
#pragma unroll
                        for (int j = 0; j < TG; ++j) {
                            if (j < ja || j >= jb) continue;
                            const double* rp = &buf[stage][r][j][0];
                            const double xn = rp[D];
                            const double w = UNIT ? 1.0 : rp[D + 1];
#pragma unroll
                            for (int i = 0; i < TL; ++i) {
                                double dot = (FAM == SOBER_TANIMOTO) ? 0.0 : zn[i];
#pragma unroll
                                for (int k = 0; k < D; ++k) dot = fma(rp[k], zt[i][k], dot);
                                const double v = (FAM == SOBER_TANIMOTO) ? tanimoto_value(dot, xn, zn[i])
                                                                         : stationary_value<FAM>(xn + dot, tab_s);
                                acc[i][j] = fma(v, w, acc[i][j]);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with this stage before it is refilled
        if (t == 0 && c + REC_STAGES < nchunks) issue(c + REC_STAGES);
    }
    if (tw_warp) {
#pragma unroll
        for (int j = 0; j < TG; ++j)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) tw[j] += __shfl_xor_sync(0xffffffffu, tw[j], off);
    }
    const bool count_tw = tw_warp && lane == 0;

    double* out = p.out + (int64_t)blockIdx.z * p.S * p.L;
#pragma unroll
    for (int j = 0; j < TG; ++j) {
        const int g = g0 + j;
        if (g >= p.S) continue;
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            const int l = l0 + lane + 32 * i;
            if (l < p.L) out[(int64_t)g * p.L + l] = acc[i][j] * p.scale;
        }
        if (count_tw) p.totw_out[(int64_t)blockIdx.z * p.S + g] = tw[j];
    }
}

// -------------------------------------------------------------------------------------------------
// generic tiled kernel (indexed layout)
// -------------------------------------------------------------------------------------------------
constexpr int TM = 64;   // landmarks per CTA
constexpr int TN = 64;   // groups per CTA
constexpr int KC = 16;   // contraction chunk
constexpr int LDS_ROW = 65;

// POST (SOBER/BASQ/_scale_mmlt.py:256-275): the kernel value enters a NON-LINEAR function of the GP
// posterior covariance,  v(l, i) = expm1( s k(z_l, x_i) - <aw_l, kx_i> ),  so the stacked-landmark trick of the
// predictive-covariance mode does not apply: the (L x n_obs).(n_obs x N) contraction runs here, through the same
// shared-memory tiles as the distance contraction, and never leaves the registers.
template <int FAM, bool POST = false>
__global__ void __launch_bounds__(256) group_tiled_kernel(const GroupParams p) {
    __shared__ double Zs[KC][LDS_ROW];
    __shared__ double Xs[KC][LDS_ROW];
    __shared__ double s_w[TN], s_xn[TN];
    __shared__ int64_t s_row[TN];
    __shared__ double tab[EXP_TAB_SIZE];

    const int t = threadIdx.x;
    const int tx = t & 15, ty = t >> 4;
    const int g0 = blockIdx.x * TN;
    const int l0 = blockIdx.y * TM;
    load_exp_table(tab, t, 256);
    const uint32_t tab_s = smem_addr(tab);

    double zn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int l = l0 + ty + 16 * i;
        zn[i] = l < p.L ? __ldg(p.zn + l) : 0.0;
    }
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    double tw = 0.0;

    const int64_t r0 = p.row_begin + (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t r1 = min(p.row_end, r0 + p.rows_per_split);
    const int64_t hi = p.pos0 + p.n_local;
    const int kk_ld = t & 15;   // contraction index this thread loads
    const int c_ld = t >> 4;    // first of 4 tile columns this thread loads (c_ld + 16 * pass)

    for (int64_t e = r0; e < r1; ++e) {
        if (t < TN) {
            const int g = g0 + t;
            const int64_t pos = e * p.S + g;
            const bool ok = (g < p.S) && (pos >= p.pos0) && (pos < hi);
            int64_t row = -1;
            double w = 0.0, xn = 0.0;
            if (ok) {
                const int64_t loc = pos - p.pos0;
                row = p.idx ? (int64_t)__ldg(p.idx + loc) : loc;
                w = p.mu ? __ldg(p.mu + loc) : 1.0;
                xn = __ldg(p.xn + row * p.xn_stride);
                if (pos < p.ES) tw += w;
            }
            s_row[t] = row;
            s_w[t] = w;
            s_xn[t] = xn;
        }
        __syncthreads();

        double dot[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dot[i][j] = 0.0;

        for (int k0 = 0; k0 < p.d; k0 += KC) {
            const int k = k0 + kk_ld;
            const bool kok = k < p.d;
#pragma unroll
            for (int pass = 0; pass < 4; ++pass) {
                const int c = c_ld + 16 * pass;
                const int l = l0 + c;
                Zs[kk_ld][c] = (kok && l < p.L) ? __ldg(p.Zt + (int64_t)l * p.d + k) : 0.0;
                const int64_t row = s_row[c];
                Xs[kk_ld][c] = (kok && row >= 0) ? __ldg(p.X + row * p.ldx + k) : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) {
                double z[4], x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) z[i] = Zs[kk][ty + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) x[j] = Xs[kk][tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dot[i][j] = fma(z[i], x[j], dot[i][j]);
            }
            __syncthreads();
        }
        double corr[4][4];
        if (POST) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) corr[i][j] = 0.0;
            for (int k0 = 0; k0 < p.n_obs; k0 += KC) {
                const int k = k0 + kk_ld;
                const bool kok = k < p.n_obs;
#pragma unroll
                for (int pass = 0; pass < 4; ++pass) {
                    const int c = c_ld + 16 * pass;
                    const int l = l0 + c;
                    Zs[kk_ld][c] = (kok && l < p.L) ? __ldg(p.aw + (int64_t)l * p.n_obs + k) : 0.0;
                    const int64_t row = s_row[c];
                    Xs[kk_ld][c] = (kok && row >= 0) ? __ldg(p.kx + row * p.ldkx + k) : 0.0;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < KC; ++kk) {
                    double z[4], x[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) z[i] = Zs[kk][ty + 16 * i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) x[j] = Xs[kk][tx + 16 * j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) corr[i][j] = fma(z[i], x[j], corr[i][j]);
                }
                __syncthreads();
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double w = s_w[tx + 16 * j];
            const double xn = s_xn[tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double kv = kernel_value<FAM>(dot[i][j], xn, zn[i], tab_s);
                if (POST) kv = expm1(fma(kv, p.inner_scale, -corr[i][j]));
                acc[i][j] = fma(kv, w, acc[i][j]);
            }
        }
        __syncthreads();
    }

    double* out = p.out + (int64_t)blockIdx.z * p.S * p.L;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int g = g0 + tx + 16 * j;
        if (g >= p.S) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int l = l0 + ty + 16 * i;
            if (l < p.L) out[(int64_t)g * p.L + l] = acc[i][j] * p.scale;
        }
    }
    if (blockIdx.y == 0 && t < TN && g0 + t < p.S) p.totw_out[(int64_t)blockIdx.z * p.S + g0 + t] = tw;
}

// -------------------------------------------------------------------------------------------------
// Tanimoto on bit-packed fingerprints: <x, z> = popcount(x & z)   (indexed layout)
// -------------------------------------------------------------------------------------------------
// A lane holds TL landmarks' bit rows in registers (W 64-bit words each); a CTA (8 warps) covers 256 * TL landmarks
// and TG groups.  Candidate bit rows of a chunk of rows are staged into shared memory with 16-byte loads and read
// back as broadcast LDS.  Per pair: W x (AND + POPC) on the integer pipe, then the FP64 Tanimoto ratio (one
// division = 7 FP64 instructions) and the weighted accumulate.  d = 1024: 16 words -> ~70 integer + 12 FP64
// instructions per pair instead of 1024 DFMAs.
constexpr int BITS_ROWS = 8;   // rows per staged chunk

template <int W, int TL, int TG, bool LUT>
__global__ void __launch_bounds__(256) group_bits_kernel(const GroupParams p) {
    // LUT: stationary kernel on binary rows -- k = lut[popcount(x ^ z)]; else Tanimoto -- popcount(x & z) + ratio
    extern __shared__ double lut_s[];                    // d + 1 entries (LUT only)
    __shared__ __align__(16) uint64_t xb[BITS_ROWS][TG][W];
    __shared__ double s_w[BITS_ROWS][TG], s_xn[BITS_ROWS][TG];
    __shared__ int64_t s_row[BITS_ROWS][TG];

    const uint64_t* __restrict__ Xw = reinterpret_cast<const uint64_t*>(p.X);
    const uint64_t* __restrict__ Zw = reinterpret_cast<const uint64_t*>(p.Zt);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g0 = blockIdx.x * TG;
    const int l0 = (blockIdx.y * 8 + warp) * (32 * TL);
    const bool active = l0 < p.L;
    if (LUT) {
        for (int h = t; h <= p.d; h += 256) lut_s[h] = p.lut[h];
        // first use is after the block barriers of the staging loop below
    }

    uint64_t zb[TL][W];
    double zn[TL];
#pragma unroll
    for (int i = 0; i < TL; ++i) {
        const int l = l0 + lane + 32 * i;
        const bool ok = l < p.L;
#pragma unroll
        for (int w = 0; w < W; ++w) zb[i][w] = ok ? __ldg(Zw + (int64_t)l * W + w) : 0ull;
        zn[i] = ok ? __ldg(p.zn + l) : 0.0;
    }
    double acc[TL][TG], tw[TG];
#pragma unroll
    for (int j = 0; j < TG; ++j) {
        tw[j] = 0.0;
#pragma unroll
        for (int i = 0; i < TL; ++i) acc[i][j] = 0.0;
    }
    const int64_t r0 = p.row_begin + (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t r1 = min(p.row_end, r0 + p.rows_per_split);
    const int64_t hi = p.pos0 + p.n_local;
    const bool count_tw = (blockIdx.y == 0) && (t == 0);
    constexpr int CHUNK = BITS_ROWS * TG;
    constexpr int VEC_PER_ROW = W / 2 > 0 ? W / 2 : 1;   // 16-byte vectors per bit row (W even) -- W == 1 handled below

    for (int64_t e0 = r0; e0 < r1; e0 += BITS_ROWS) {
        __syncthreads();   // previous chunk fully consumed
        if (t < CHUNK) {
            const int r = t / TG, j = t % TG;
            const int g = g0 + j;
            const int64_t e = e0 + r;
            const int64_t pos = e * p.S + g;
            const bool ok = (e < r1) && (g < p.S) && (pos >= p.pos0) && (pos < hi);
            int64_t row = -1;
            double w = 0.0, xn = 0.0;
            if (ok) {
                const int64_t loc = pos - p.pos0;
                row = p.idx ? (int64_t)__ldg(p.idx + loc) : loc;
                w = p.mu ? __ldg(p.mu + loc) : 1.0;
                xn = __ldg(p.xn + row * p.xn_stride);
            }
            s_row[r][j] = row;
            s_w[r][j] = w;
            s_xn[r][j] = xn;
        }
        __syncthreads();
        if (W >= 2) {
            for (int v = t; v < CHUNK * VEC_PER_ROW; v += 256) {
                const int c = v / VEC_PER_ROW, q = v % VEC_PER_ROW;
                const int64_t row = s_row[c / TG][c % TG];
                uint4 val = make_uint4(0, 0, 0, 0);
                if (row >= 0) val = __ldg(reinterpret_cast<const uint4*>(Xw + row * p.ldx) + q);
                reinterpret_cast<uint4*>(&xb[c / TG][c % TG][0])[q] = val;
            }
        } else {
            for (int c = t; c < CHUNK; c += 256) {
                const int64_t row = s_row[c / TG][c % TG];
                xb[c / TG][c % TG][0] = row >= 0 ? __ldg(Xw + row * p.ldx) : 0ull;
            }
        }
        __syncthreads();
        if (count_tw) {
            for (int r = 0; r < BITS_ROWS; ++r)
#pragma unroll
                for (int j = 0; j < TG; ++j)
                    if (s_row[r][j] >= 0 && (e0 + r) * p.S + g0 + j < p.ES) tw[j] += s_w[r][j];
        }
        if (active) {
            const int nrows = (int)min((int64_t)BITS_ROWS, r1 - e0);
            for (int r = 0; r < nrows; ++r) {
#pragma unroll
                for (int j = 0; j < TG; ++j) {
                    if (s_row[r][j] < 0) continue;     // warp-uniform
                    int cnt[TL];
#pragma unroll
                    for (int i = 0; i < TL; ++i) cnt[i] = 0;
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const uint64_t xw = xb[r][j][w];
#pragma unroll
                        for (int i = 0; i < TL; ++i) cnt[i] += __popcll(LUT ? (xw ^ zb[i][w]) : (xw & zb[i][w]));
                    }
                    const double xn = s_xn[r][j], wgt = s_w[r][j];
#pragma unroll
                    for (int i = 0; i < TL; ++i) {
                        const double kv = LUT ? lut_s[cnt[i]] : tanimoto_bits_value((double)cnt[i], xn, zn[i] + 1e-6);
                        acc[i][j] = fma(kv, wgt, acc[i][j]);
                    }
                }
            }
        }
    }

    double* out = p.out + (int64_t)blockIdx.z * p.S * p.L;
#pragma unroll
    for (int j = 0; j < TG; ++j) {
        const int g = g0 + j;
        if (g >= p.S) continue;
#pragma unroll
        for (int i = 0; i < TL; ++i) {
            const int l = l0 + lane + 32 * i;
            if (l < p.L) out[(int64_t)g * p.L + l] = acc[i][j] * p.scale;
        }
        if (count_tw) p.totw_out[(int64_t)blockIdx.z * p.S + g] = tw[j];
    }
}

// fixed-order reduction of the row-splits; applies the output scale
__global__ void reduce_splits_kernel(const double* __restrict__ part, const double* __restrict__ tw_part,
                                     int nsplit, int64_t SL, int S, double scale, double* __restrict__ At,
                                     double* __restrict__ totw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < SL) {
        double s = 0.0;
        for (int z = 0; z < nsplit; ++z) s += part[(int64_t)z * SL + i];
        At[i] = s * scale;
    }
    if (i < S) {
        double s = 0.0;
        for (int z = 0; z < nsplit; ++z) s += tw_part[(int64_t)z * S + i];
        totw[i] = s;
    }
}

// -------------------------------------------------------------------------------------------------
// reduction of a Gram tile produced by an opaque callable
// -------------------------------------------------------------------------------------------------
__global__ void group_gram_kernel(const double* __restrict__ G, int64_t ldg, int L, int64_t m,
                                  const double* __restrict__ mu, int64_t pos_begin, int64_t ES, int S,
                                  double* __restrict__ At, double* __restrict__ totw) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y * blockDim.y + threadIdx.y;
    if (g >= S || l >= L) return;
    // first j with (pos_begin + j) mod S == g
    int64_t j = (g - (pos_begin % S) + S) % S;
    double s = 0.0, tw = 0.0;
    for (; j < m; j += S) {
        const double w = mu ? mu[j] : 1.0;
        s = fma(G[(int64_t)l * ldg + j], w, s);
        if (pos_begin + j < ES) tw += w;
    }
    At[(int64_t)g * L + l] += s;
    if (l == 0 && totw) totw[g] += tw;
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
#ifndef SOBER_REC_TL
#define SOBER_REC_TL 2
#endif
#ifndef SOBER_REC_TG
#define SOBER_REC_TG 4
#endif
constexpr int REC_TL = SOBER_REC_TL;   // landmarks per lane
constexpr int REC_TG = SOBER_REC_TG;   // groups per CTA

constexpr int BITS_TG = 4;
static int bits_tl(int W) { return W <= 8 ? 4 : (W <= 16 ? 2 : 1); }

struct Plan {
    bool bits_mma;     // bit-packed Tanimoto on the tensor cores (csrc/group_bits_mma.cu)
    bool bits_mma_v1;  // variant 5: the first tcgen05 kernel (landmark tile in shared memory), kept for comparison
    bool records;
    bool bits;
    dim3 grid, block;
    int nsplit;
    int64_t rows_per_split, row_begin, row_end;
};

static bool plan_group(const sober_group_args* a, Plan* pl, int sms = 0) {
    if (sms <= 0) sms = sm_count();
    if (!a || a->S <= 0 || a->L <= 0 || a->d <= 0 || a->n_local < 0 || a->pos0 < 0) return false;
    const int64_t hi = a->pos0 + a->n_local;
    pl->row_begin = a->pos0 / a->S;
    pl->row_end = a->n_local > 0 ? ceil_div(hi, a->S) : pl->row_begin;
    const int64_t rows = pl->row_end - pl->row_begin;
    const bool post = a->kx != nullptr || a->aw != nullptr;
    if (post && (!a->kx || !a->aw || a->n_obs <= 0 || a->ldkx < a->n_obs)) return false;
    pl->records = a->rec != nullptr && a->variant != 1 && !post;
    pl->bits = a->family == SOBER_TANIMOTO_BITS || a->family == SOBER_HAMMING_LUT;
    if (a->family == SOBER_HAMMING_LUT && !a->lut) return false;
    if (pl->bits && post) return false;
    if (pl->bits) {
        const int64_t W = a->ldx;
        if (pl->records || a->d <= 0 || W < (a->d + 63) / 64 || !(W == 1 || W == 2 || W == 4 || W == 8 || W == 16 || W == 32))
            return false;
    }
    if (pl->records && (a->d > 8 || a->ldr != (a->d + 3) / 2 * 2)) return false;
    int64_t gx, gy, target;
    pl->bits_mma = a->family == SOBER_TANIMOTO_BITS && a->variant != 4 && bits_mma_supported((int)a->ldx) &&
                   a->n_local * (int64_t)a->L >= (1 << 20);
    pl->bits_mma_v1 = a->variant == 5;
    if (pl->bits_mma) {
        gx = ceil_div(a->S, pl->bits_mma_v1 ? 128 : 64);
        gy = ceil_div(a->L, pl->bits_mma_v1 ? 64 : 128);
        pl->block = dim3(416);
        target = (int64_t)sms * 3;    // one 167 KB CTA per SM, a few waves
    } else if (pl->bits) {
        gx = ceil_div(a->S, BITS_TG);
        gy = ceil_div(a->L, 8 * 32 * bits_tl((int)a->ldx));
        pl->block = dim3(256);
        target = (int64_t)sms * 12;
    } else if (pl->records) {
        gx = ceil_div(a->S, REC_TG);
        gy = ceil_div(a->L, REC_WARPS * 32 * REC_TL);
        pl->block = dim3(REC_THREADS);
        target = (int64_t)sms * 12;   // many more CTAs than SMs: the tail wave costs < 1/12
    } else {
        gx = ceil_div(a->S, TN);
        gy = ceil_div(a->L, TM);
        pl->block = dim3(256);
        target = (int64_t)sms * 8;
    }
    int64_t ns = ceil_div(target, gx * gy);
    if (pl->bits_mma) {
        // one CTA per SM: pick the split count whose last wave is fullest (3.46 waves would run as 4: -14 %)
        const int64_t smsl = sms;
        int64_t best = ns;
        double best_eff = 0.0;
        for (int64_t c = ns; c <= ns + 6 && c <= rows; ++c) {
            const int64_t ctas = gx * gy * c;
            const double eff = (double)ctas / (double)(ceil_div(ctas, smsl) * smsl);
            if (eff > best_eff + 1e-9) { best_eff = eff; best = c; }
        }
        ns = best;
    }
    if (pl->records && !pl->bits_mma && rows > 0) {
        // Wave-aware split count for the register kernel (2 CTAs of 256 threads x 126 registers per SM): the CTAs of one
        // launch do equal work, so the launch takes ceil(CTAs / resident) "waves" of rows_per_split rows each -- 1800
        // CTAs on 296 slots ran as 7 waves with the last one 8 % full.  Pick the split count that minimises
        // waves x (rows_per_split + 1/2) -- half a row for a CTA's prologue: exponential table, landmark registers --
        // plus a quarter of a row per split for the second-stage reduction it feeds.  Small launches (a few dozen rows)
        // then run as ONE wave of several rows per CTA instead of five waves of single-row CTAs.
        static const int wave_aware = [] { const char* e = getenv("SOBER_B200_K1_WAVES"); return e ? atoi(e) : 1; }();
        const int64_t resident = (int64_t)sms * 2;
        if (wave_aware) {
            int64_t best = ns;
            double best_cost = 1e300;
            for (int64_t c = 1; c <= ns + 8 && c <= rows; ++c) {
                const int64_t rps = ceil_div(rows, c), nsp = ceil_div(rows, rps);
                const double cost = (double)ceil_div(gx * gy * nsp, resident) * ((double)rps + 0.5) + 0.25 * (double)nsp;
                if (cost < best_cost - 1e-9) { best_cost = cost; best = c; }
            }
            ns = best;
        }
        if (const char* e = getenv("SOBER_B200_K1_NS")) { if (atoi(e) > 0) ns = atoi(e); }   // tuning aid
    }
    if (ns > rows) ns = rows;
    if (ns < 1) ns = 1;
    if (ns > 65535) ns = 65535;
    pl->rows_per_split = rows > 0 ? ceil_div(rows, ns) : 1;
    pl->nsplit = rows > 0 ? (int)ceil_div(rows, pl->rows_per_split) : 1;
    pl->grid = dim3((unsigned)gx, (unsigned)gy, (unsigned)pl->nsplit);
    return gy <= 65535;
}

template <int D, int FAM>
static void launch_records(const Plan& pl, const GroupParams& p, cudaStream_t st) {
    if (p.unit_weights)
        group_records_kernel<D, FAM, REC_TL, REC_TG, true><<<pl.grid, pl.block, 0, st>>>(p);
    else
        group_records_kernel<D, FAM, REC_TL, REC_TG, false><<<pl.grid, pl.block, 0, st>>>(p);
}

template <int FAM>
static bool launch_records_d(const Plan& pl, const GroupParams& p, cudaStream_t st) {
    switch (p.d) {
        case 1: launch_records<1, FAM>(pl, p, st); return true;
        case 2: launch_records<2, FAM>(pl, p, st); return true;
        case 3: launch_records<3, FAM>(pl, p, st); return true;
        case 4: launch_records<4, FAM>(pl, p, st); return true;
        case 5: launch_records<5, FAM>(pl, p, st); return true;
        case 6: launch_records<6, FAM>(pl, p, st); return true;
        case 7: launch_records<7, FAM>(pl, p, st); return true;
        case 8: launch_records<8, FAM>(pl, p, st); return true;
        default: return false;
    }
}

template <bool LUT>
static bool launch_bits(const Plan& pl, const GroupParams& p, cudaStream_t st) {
    const size_t sm = LUT ? (size_t)(p.d + 1) * 8 : 0;
    switch ((int)p.ldx) {
        case 1: group_bits_kernel<1, 4, BITS_TG, LUT><<<pl.grid, pl.block, sm, st>>>(p); return true;
        case 2: group_bits_kernel<2, 4, BITS_TG, LUT><<<pl.grid, pl.block, sm, st>>>(p); return true;
        case 4: group_bits_kernel<4, 4, BITS_TG, LUT><<<pl.grid, pl.block, sm, st>>>(p); return true;
        case 8: group_bits_kernel<8, 4, BITS_TG, LUT><<<pl.grid, pl.block, sm, st>>>(p); return true;
        case 16: group_bits_kernel<16, 2, BITS_TG, LUT><<<pl.grid, pl.block, sm, st>>>(p); return true;
        case 32: group_bits_kernel<32, 1, BITS_TG, LUT><<<pl.grid, pl.block, sm, st>>>(p); return true;
        default: return false;
    }
}

template <int FAM>
static bool launch_family(const Plan& pl, const GroupParams& p, cudaStream_t st) {
    if (pl.records) return launch_records_d<FAM>(pl, p, st);
    if (p.kx)
        group_tiled_kernel<FAM, true><<<pl.grid, pl.block, 0, st>>>(p);
    else
        group_tiled_kernel<FAM><<<pl.grid, pl.block, 0, st>>>(p);
    return true;
}

}  // namespace sober

using namespace sober;

static int64_t plan_workspace(const sober_group_args* a, const Plan& pl) {
    if (pl.nsplit <= 1) return 0;  // single split: the kernel writes At / totw directly
    return (int64_t)pl.nsplit * ((int64_t)a->S * a->L + a->S) * 8;
}

// The split count depends on the SMs the launching stream can use (whole device, or the SM partition of
// sober_partition_stream): the query does not know the stream and returns the larger of the two needs.
extern "C" int64_t sober_group_accumulate_workspace(const sober_group_args* a) {
    Plan pl;
    if (!plan_group(a, &pl)) return -1;
    int64_t need = plan_workspace(a, pl);
    const int part = partition_sm_count();
    if (part > 0 && plan_group(a, &pl, part)) need = std::max(need, plan_workspace(a, pl));
    return need;
}

extern "C" int sober_group_accumulate(const sober_group_args* a, void* workspace, int64_t workspace_bytes,
                                      void* stream) {
    Plan pl;
    if (!plan_group(a, &pl, stream_sm_count(stream))) return SOBER_ERR_ARG;
    if (!a->Zt || !a->zn || !a->At || !a->totw) return SOBER_ERR_ARG;
    if (!pl.records && (!a->X || !a->xn)) return SOBER_ERR_ARG;
    if (a->family < SOBER_RBF || a->family > SOBER_HAMMING_LUT) return SOBER_ERR_UNSUPPORTED;
    const int64_t need = plan_workspace(a, pl);
    if (need > workspace_bytes || (need > 0 && !workspace)) return SOBER_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t SL = (int64_t)a->S * a->L;

    if (pl.row_end <= pl.row_begin) {  // nothing owned: zero contribution
        SOBER_CUDA_CHECK(cudaMemsetAsync(a->At, 0, SL * 8, st));
        SOBER_CUDA_CHECK(cudaMemsetAsync(a->totw, 0, (int64_t)a->S * 8, st));
        return SOBER_OK;
    }

    GroupParams p;
    p.X = a->X; p.ldx = a->ldx; p.xn = a->xn; p.xn_stride = a->xn_stride;
    p.idx = a->idx; p.mu = a->mu; p.rec = a->rec;
    p.n_local = a->n_local; p.pos0 = a->pos0; p.ES = a->ES;
    p.S = a->S; p.L = a->L; p.d = a->d;
    p.unit_weights = a->unit_weights;
    p.Zt = a->Zt; p.zn = a->zn; p.lut = a->lut;
    p.kx = a->kx; p.ldkx = a->ldkx; p.aw = a->aw; p.n_obs = a->n_obs;
    p.inner_scale = a->outputscale;   // POST: the output scale belongs INSIDE the transform
    p.row_begin = pl.row_begin; p.row_end = pl.row_end; p.rows_per_split = pl.rows_per_split;
    double* ws = (double*)workspace;
    const double outer_scale = a->kx ? 1.0 : a->outputscale;
    if (pl.nsplit == 1) {
        p.out = a->At;
        p.totw_out = a->totw;
        p.scale = outer_scale;
    } else {
        p.out = ws;
        p.totw_out = ws + (int64_t)pl.nsplit * SL;
        p.scale = 1.0;
    }
    if (pl.bits_mma) {
        BitsMmaParams q;
        q.X = reinterpret_cast<const uint64_t*>(a->X); q.Z = reinterpret_cast<const uint64_t*>(a->Zt);
        q.xn = a->xn; q.xn_stride = a->xn_stride; q.zn = a->zn; q.idx = a->idx; q.mu = a->mu;
        q.n_local = a->n_local; q.pos0 = a->pos0; q.ES = a->ES; q.S = a->S; q.L = a->L; q.W = (int)a->ldx;
        q.out = p.out; q.totw_out = p.totw_out;
        q.row_begin = pl.row_begin; q.row_end = pl.row_end; q.rows_per_split = pl.rows_per_split; q.scale = p.scale;
        const int rc = pl.bits_mma_v1 ? launch_bits_mma(q, pl.grid, st) : launch_bits_mma2(q, pl.grid, st);
        if (rc != SOBER_OK) return rc;
        if (pl.nsplit > 1) {
            const int64_t n = SL > a->S ? SL : a->S;
            reduce_splits_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ws, ws + (int64_t)pl.nsplit * SL, pl.nsplit,
                                                                            SL, a->S, outer_scale, a->At, a->totw);
            SOBER_LAUNCH_CHECK("reduce_splits");
        }
        return SOBER_OK;
    }
    bool ok = false;
    switch (a->family) {
        case SOBER_RBF: ok = launch_family<SOBER_RBF>(pl, p, st); break;
        case SOBER_MATERN12: ok = launch_family<SOBER_MATERN12>(pl, p, st); break;
        case SOBER_MATERN32: ok = launch_family<SOBER_MATERN32>(pl, p, st); break;
        case SOBER_MATERN52: ok = launch_family<SOBER_MATERN52>(pl, p, st); break;
        case SOBER_TANIMOTO: ok = launch_family<SOBER_TANIMOTO>(pl, p, st); break;
        case SOBER_TANIMOTO_BITS: ok = launch_bits<false>(pl, p, st); break;
        case SOBER_HAMMING_LUT: ok = launch_bits<true>(pl, p, st); break;
    }
    if (!ok) return SOBER_ERR_UNSUPPORTED;
    SOBER_LAUNCH_CHECK("group_accumulate");
    if (pl.nsplit > 1) {
        const int64_t n = SL > a->S ? SL : a->S;
        reduce_splits_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ws, ws + (int64_t)pl.nsplit * SL, pl.nsplit,
                                                                        SL, a->S, outer_scale, a->At, a->totw);
        SOBER_LAUNCH_CHECK("reduce_splits");
    }
    return SOBER_OK;
}

extern "C" int sober_group_accumulate_gram(const double* G, int64_t ldg, int32_t L, int64_t m, const double* mu,
                                           int64_t pos_begin, int64_t ES, int32_t S, double* At, double* totw,
                                           void* stream) {
    if (!G || !At || L <= 0 || S <= 0 || m < 0 || ldg < m) return SOBER_ERR_ARG;
    if (m == 0) return SOBER_OK;
    dim3 block(32, 8);
    dim3 grid((unsigned)ceil_div(S, 32), (unsigned)ceil_div(L, 8));
    if (grid.y > 65535) return SOBER_ERR_UNSUPPORTED;
    group_gram_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(G, ldg, L, m, mu, pos_begin, ES, S, At, totw);
    SOBER_LAUNCH_CHECK("group_gram");
    return SOBER_OK;
}
