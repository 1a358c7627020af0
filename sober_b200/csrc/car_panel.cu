// K3 (panelled) -- the k dependent elimination steps of SOBER/_rchq.py:237-266 as a BLOCKED factorisation.
//
// The elimination is Gaussian elimination of the null-space basis Phi (S x k) with the pivot ROW of step s chosen by
// the ratio test (argmin of mu_i / Phi[i, s] over Phi[i, s] > 0).  Only column s has to be up to date when its pivot
// is chosen, so the steps are grouped into panels of nb <= 64 columns:
//
//   car_panel_kernel   one 8-CTA cluster, rows of the panel distributed over the CTAs (CTA r owns rows
//                      [r*rpc, (r+1)*rpc)).  The S x nb panel lives in shared memory (column-major); the 8 columns being
//                      eliminated (a "block") live in REGISTERS of the row threads, because shared memory delivers 16
//                      doubles per clock against 64 FMAs per clock (the first version, all in shared memory, spent
//                      2/3 of a step waiting for the LDS/STS of the rank-1 update).  Per step every CTA finds its best
//                      candidate for the NEXT column right after updating it (lookahead) and posts
//                      [alpha, 1/v, row | the candidate's 8 block entries] to all 8 CTAs with st.async (DSMEM stores that
//                      complete a transaction count on the receiver's mbarrier: no fence, no cluster barrier); each CTA
//                      then picks the same winner and updates its registers.  The pivot row's entries RIGHT of the block
//                      travel off the critical path (owner -> everybody); at the end of the block they are brought up
//                      to date (R = L^-1 G, 8 x 8 unit lower-triangular L) and the rest of the panel gets ONE rank-8
//                      update (1 LDS + 8 FMA + 1 STS per element).
//   car_panel_solve    R = L^-1 Phi[J, rest]: the pivot rows of the trailing columns brought up to date by forward
//                      substitution with the unit lower-triangular L[s, c] = u_c[j_s] (one warp per column).
//   car_panel_update   Phi[:, rest] -= U R  (S x nb by nb x rest FP64 GEMM over the whole GPU, U = scaled pivot
//                      columns u_s = v_s / v_s[j_s]) and exact zeros in the pivot rows.
//
// When the whole S x k basis fits the cluster's shared memory (S = 400, k = 200: 82 KB per CTA) it is ONE panel and
// the two trailing kernels never run.  Arithmetic is the fused variant (one reciprocal per pivot, FMA per element); the
// reference's unfused order (bit-identical pivots on an injected basis) stays with the EXACT kernels of
// car_cluster_cols.cu / car_eliminate.cu.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace sober {

constexpr int CP_THREADS = 256;
constexpr int CP_P = 8;         // CTAs per cluster
constexpr int CP_HDR = 4;       // header of a posted candidate: alpha, 1 / v, global row index (-1: none), pad
constexpr int CP_NB = 64;       // panel width of the blocked path (the solve kernel holds 2 entries per lane)
constexpr int CP_SMEM_MAX = 200 * 1024;

struct CarPanelParams {
    double* basis;   // k x S rows; rows [t0, t0 + nb) are the panel; overwritten with u_s when write_u
    double* mu;      // S in/out
    int* state;      // [0] stop flag (no positive entry: the guard of SOBER/_rchq.py:241-242), [1] steps taken
    int* piv;        // k pivots (global row index per step)
    double* lmat;    // CP_NB x CP_NB: lmat[s][c] = u_c[j_s], c < s (blocked path only)
    long long* prof; // optional 8 cycle counters of CTA 0 / thread 0
    int S, k, t0, nb;
    int rpc, rpcp, tr;
    int write_u;
};

// order-preserving map double -> uint64 (total order of the IEEE values, negative numbers first)
__device__ __forceinline__ unsigned long long cp_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double cp_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
constexpr unsigned long long CP_NONE = 0xffffffffffffffffull;

// warp argmin of (key, idx), lowest idx on ties: three REDUX.MIN
__device__ __forceinline__ void cp_warp_argmin(unsigned long long& key, int& idx) {
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const bool c1 = hi == mhi;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, c1 ? lo : 0xffffffffu);
    const bool c2 = c1 && lo == mlo;
    const unsigned mi = __reduce_min_sync(0xffffffffu, c2 ? (unsigned)idx : 0x7fffffffu);
    key = ((unsigned long long)mhi << 32) | mlo;
    idx = (int)mi;
}

__device__ __forceinline__ uint32_t cp_mapa(uint32_t saddr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(saddr), "r"(rank));
    return out;
}
// 8-byte DSMEM store that completes 8 bytes of the receiver's mbarrier transaction count
__device__ __forceinline__ void cp_st_async(uint32_t dst_cluster, double v, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst_cluster),
                 "l"(__double_as_longlong(v)), "r"(bar_cluster)
                 : "memory");
}

// a / b and 1 / b for b > 0 (MUFU seed + two Newton steps + one correction of the quotient: <= 1 ulp)
__device__ __forceinline__ double cp_div(double a, double b, double& y) {
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}

// st.async of 16 bytes
__device__ __forceinline__ void cp_st_async2(uint32_t dst_cluster, double a, double b, uint32_t bar_cluster) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(
                     dst_cluster),
                 "l"(__double_as_longlong(a)), "l"(__double_as_longlong(b)), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ unsigned long long cp_min64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }

// The ratio key carries the row in its low bits: keys are compared as 64-bit integers, ratios that agree in their
// leading 53 - CP_IDX_BITS mantissa bits (2e-12 relative) count as ties and go to the lowest row, as torch.argmin does
// for exact ties.  (Fused arithmetic already differs from the reference's at the 1e-16 level.)
constexpr int CP_IDX_BITS = 13;   // local row index < 8192
__device__ __forceinline__ unsigned long long cp_pack(double ratio, int row) {
    return (cp_key(ratio) & ~((1ull << CP_IDX_BITS) - 1)) | (unsigned long long)row;
}

constexpr int CP_B = 8;           // register block: columns of the panel a row thread holds in registers
constexpr int CP_HB = CP_HDR + CP_B;

// Roles inside a CTA:  warps 0 .. NPW-1 = PIVOT warps (together they hold all rows of the CTA's block in registers, at
// most two rows per lane: ratio test, argmin and register update stay inside the warp; with NPW > 1 the warps' best
// rows meet in shared memory behind a named barrier of the pivot warps only);  warp 7 = HELPER (follows every
// exchange: records the multipliers L and the pivots, and -- in the CTA that owns the pivot row -- sends that row's
// entries right of the block to everybody);  the other warps sleep at the block-end barrier and join for the rank-8
// update of the rest of the panel.
constexpr int CP_HELPER = CP_THREADS / 32 - 1;

template <int RPL, int NPW>
__global__ void __launch_bounds__(CP_THREADS, 1) car_panel_kernel(const CarPanelParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int P = CP_P, B = CP_B, HB = CP_HB;
    const int r = (int)cluster.block_rank();
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int S = p.S, nb = p.nb, rpcp = p.rpcp, tr = p.tr;
    const int tc = CP_THREADS / tr;
    const int ri = t & (tr - 1), ci = t / tr;     // mapping of the block-end update: row ri, columns ci, ci + tc, ...
    const int row0 = r * p.rpc;
    const int nrows = max(0, min(p.rpc, S - row0));
    const bool pivot = warp < NPW, helper = warp == CP_HELPER;

    extern __shared__ __align__(16) double sm[];
    double* panel = sm;                            // [nb][rpcp]   column-major panel (u columns once a block is done)
    double* recv = panel + (size_t)nb * rpcp;      // [2][P][HB]   candidates of the current / next step
    double* gbuf = recv + 2 * P * HB;              // [nb][B]      raw pivot rows of the block, then R (transposed)
    double* lblk = gbuf + (size_t)nb * B;          // [B][B]       in-block multipliers L[q][q']
    double* stage = lblk + B * B;                  // [2][4][HB]   per exchange parity and pivot warp: its best row
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ __align__(8) uint64_t gbar;
    __shared__ unsigned long long red_key[2][4];
    __shared__ int jblk_s[B];
    __shared__ volatile int helper_read;           // exchanges whose buffer the helper has finished reading
    __shared__ volatile int stop_s;

    if (*(volatile int*)p.state != 0) return;      // an earlier panel stopped the elimination (uniform over the grid)

    constexpr uint32_t XBYTES = P * HB * 8;        // bytes of one exchange
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&gbar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bars[0], XBYTES);
        if (nb > 1) mbar_expect_tx(&bars[1], XBYTES);
        helper_read = 0;
        stop_s = 0;
    }
    for (int e = t; e < nb * rpcp; e += CP_THREADS) {
        const int c = e / rpcp, i = e - c * rpcp;
        panel[e] = i < nrows ? p.basis[(size_t)(p.t0 + c) * S + row0 + i] : 0.0;
    }
    // rows of this lane: (warp * RPL + m) * 32 + lane
    double mu[RPL];
    bool ok[RPL];
    int rowi[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
        rowi[m] = (warp * RPL + m) * 32 + lane;
        ok[m] = pivot && rowi[m] < nrows;
        mu[m] = ok[m] ? p.mu[row0 + rowi[m]] : 0.0;
    }
    __syncthreads();
    cluster.sync();                                // every CTA's barriers are initialised before anybody posts

    const bool prof = p.prof != nullptr && r == 0 && t == 0;
    long long pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long pt = prof ? clock64() : 0;
#define CP_TICK(slot_)                     \
    if (prof) {                            \
        const long long now_ = clock64(); \
        pa[slot_] += now_ - pt;           \
        pt = now_;                        \
    }

    double blk[RPL][B];                            // pivot warps: this lane's rows of the current block
    // Pivot warps: ratio test on block column qc (registers), argmin over the CTA's rows, post of the CTA's candidate
    // for pivot column e to all CTAs.
    auto test_and_post = [&](int e, int qc) {
        const int par = e & 1;
        unsigned long long key = CP_NONE;
        double balpha = 0.0, binv = 0.0;
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
            double v = 0.0;
#pragma unroll
            for (int q = 0; q < B; ++q)
                if (q == qc) v = blk[m][q];        // qc is a compile-time constant at every call site after inlining
            if (ok[m] && v > 0.0) {
                // ranking key from a 2^-44-accurate quotient (MUFU seed + one Newton step; the key keeps 40 bits); the
                // exact quotient and reciprocal are computed once, for the lane's best row
                double y;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
                y = fma(y, fma(-v, y, 1.0), y);
                const unsigned long long kq = cp_pack(mu[m] * y, rowi[m]);
                if (kq < key) { key = kq; balpha = mu[m]; binv = v; }
            }
        }
        if (key != CP_NONE) {
            double y;
            balpha = cp_div(balpha, binv, y);
            binv = y;
        }
        const unsigned long long mine = key;
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
        key = ((unsigned long long)mhi << 32) | mlo;
        double* st = stage + (size_t)(par * 4 + warp) * HB;
        if (key != CP_NONE) {
            const int bi = (int)(key & ((1ull << CP_IDX_BITS) - 1));
            const int mstar = (bi >> 5) - warp * RPL;          // uniform: which of its rows the warp's best lane stages
            if (mine == key) {
                st[0] = balpha; st[1] = binv; st[2] = (double)(row0 + bi); st[3] = 0.0;
            }
#pragma unroll
            for (int m = 0; m < RPL; ++m)
                if (m == mstar && mine == key) {
#pragma unroll
                    for (int q = 0; q < B; ++q) st[CP_HDR + q] = blk[m][q];
                }
        }
        int wsel = warp;
        if (NPW > 1) {
            if (lane == 0) red_key[par][warp] = key;
            asm volatile("bar.sync 1, %0;" ::"r"(32 * NPW) : "memory");   // the pivot warps only
#pragma unroll
            for (int w = NPW - 1; w >= 0; --w) {
                const unsigned long long kw = red_key[par][w];
                if (kw <= key) { key = kw; wsel = w; }
            }
        } else {
            __syncwarp();
        }
        const bool have = key != CP_NONE;
        const double* win = stage + (size_t)(par * 4 + wsel) * HB;
        while (helper_read < e) {
        }
        // pivot warp w posts to the CTAs p with p mod NPW == w: 6 16-byte chunks each
        constexpr int CH = (P / NPW) * (HB / 2);
        for (int ch = lane; ch < CH; ch += 32) {
            const int pl = ch / (HB / 2), part = ch - pl * (HB / 2);
            const int peer = pl * NPW + warp;
            double a = have ? win[2 * part] : 0.0, b = have ? win[2 * part + 1] : 0.0;
            if (part == 1 && !have) a = -1.0;      // header[2] = row index, -1: no candidate
            const uint32_t dst = cp_mapa(smem_addr(recv + (size_t)(par * P + r) * HB), (uint32_t)peer);
            const uint32_t bar = cp_mapa(smem_addr(&bars[par]), (uint32_t)peer);
            cp_st_async2(dst + (uint32_t)part * 16u, a, b, bar);
        }
    };
    // winner of the 8 candidates of exchange s: lowest ratio, then lowest CTA rank (= lowest row); -1: none
    auto winner = [&](const double* rb, double& alpha) {
        double al[P];
        int wi[P];
#pragma unroll
        for (int c = 0; c < P; ++c) {
            const double2 h = *reinterpret_cast<const double2*>(rb + (size_t)c * HB);
            al[c] = h.x;
            wi[c] = rb[(size_t)c * HB + 2] >= 0.0 ? c : -1;
        }
#pragma unroll
        for (int w = 1; w < P; w *= 2)
#pragma unroll
            for (int c = 0; c + w < P; c += 2 * w) {
                const bool take = wi[c + w] >= 0 && (wi[c] < 0 || al[c + w] < al[c]);
                al[c] = take ? al[c + w] : al[c];
                wi[c] = take ? wi[c + w] : wi[c];
            }
        alpha = al[0];
        return wi[0];
    };

    int done = 0;
    bool stopped = false;
    for (int b0 = 0; b0 < nb; b0 += B) {
        const int bw = min(B, nb - b0);
        const int glen = nb - (b0 + bw);           // columns of the panel to the right of this block
        if (pivot) {
#pragma unroll
            for (int m = 0; m < RPL; ++m)
#pragma unroll
                for (int q = 0; q < B; ++q)
                    blk[m][q] = (ok[m] && q < bw) ? panel[(size_t)(b0 + q) * rpcp + rowi[m]] : 0.0;
            if (t == 0 && glen > 0) mbar_expect_tx(&gbar, (uint32_t)(bw * glen * 8));
            test_and_post(b0, 0);
            CP_TICK(4)
        }
        if (pivot || helper) {
#pragma unroll
            for (int q = 0; q < B; ++q) {
                if (q < bw && !stopped) {
                    const int s = b0 + q, par = s & 1;
                    mbar_wait(&bars[par], (uint32_t)((s >> 1) & 1));
                    CP_TICK(0)
                    const double* rb = recv + (size_t)par * P * HB;
                    double alpha;
                    const int wr = winner(rb, alpha);
                    if (wr < 0) {
                        stopped = true;
                    } else {
                        const double* wb = rb + (size_t)wr * HB;
                        const double inv = wb[1];
                        const int jg = (int)wb[2];
                        done = s + 1;
                        if (pivot) {
                            double rowv[B];
#pragma unroll
                            for (int qq = 0; qq < B; qq += 2) {
                                const double2 v2 = *reinterpret_cast<const double2*>(wb + CP_HDR + qq);
                                rowv[qq] = v2.x;
                                rowv[qq + 1] = v2.y;
                            }
                            if (t == 0 && s + 2 < nb) mbar_expect_tx(&bars[par], XBYTES);
#pragma unroll
                            for (int m = 0; m < RPL; ++m) {
                                const double v = blk[m][q];
                                const bool pv = ok[m] && (row0 + rowi[m] == jg);
                                const double u = pv ? 1.0 : v * inv;
                                blk[m][q] = u;
                                mu[m] = pv ? 0.0 : fma(-alpha, v, mu[m]);
#pragma unroll
                                for (int qq = q + 1; qq < B; ++qq)
                                    blk[m][qq] = pv ? 0.0 : fma(-u, rowv[qq], blk[m][qq]);
                            }
                            CP_TICK(1)
                            if (q + 1 < bw) test_and_post(s + 1, q + 1);
                            CP_TICK(2)
                        } else {
                            // helper: multipliers of this pivot row inside the block, the pivot, then release the buffer
                            if (lane < B) lblk[q * B + lane] = lane < q ? wb[CP_HDR + lane] : 0.0;
                            if (lane == 0) {
                                jblk_s[q] = jg;
                                if (r == 0) p.piv[p.t0 + s] = jg;
                            }
                            __syncwarp();
                            if (lane == 0) helper_read = s + 1;
                            // the owner of the pivot row sends its (not yet updated) entries of the columns right of
                            // the block to every CTA; they are brought up to date at the end of the block
                            if (glen > 0 && jg >= row0 && jg < row0 + nrows) {
                                for (int idx = lane; idx < glen * P; idx += 32) {
                                    const int c = idx / P, peer = idx - c * P;
                                    const uint32_t gdst = cp_mapa(smem_addr(gbuf + (size_t)c * B + q), (uint32_t)peer);
                                    const uint32_t gb = cp_mapa(smem_addr(&gbar), (uint32_t)peer);
                                    cp_st_async(gdst, panel[(size_t)(b0 + bw + c) * rpcp + (jg - row0)], gb);
                                }
                            }
                        }
                    }
                }
            }
            if (helper && stopped && lane == 0) helper_read = nb + 1;
            if (pivot) {
                if (stopped && t == 0) stop_s = 1;
                // u columns of the block -> panel (U of the trailing update, L rows, multipliers of the rank-8 update)
                if (!stopped) {
#pragma unroll
                    for (int m = 0; m < RPL; ++m)
                        if (ok[m]) {
#pragma unroll
                            for (int q = 0; q < B; ++q)
                                if (q < bw) panel[(size_t)(b0 + q) * rpcp + rowi[m]] = blk[m][q];
                        }
                }
                CP_TICK(3)
            }
        }
        __syncthreads();                           // (A) u columns, L, pivots of the block and the stop flag are visible
        if (stop_s) break;
        if (glen > 0) mbar_wait(&gbar, (uint32_t)((b0 / B) & 1));
        CP_TICK(5)
        if (p.write_u && helper) {
            // rows of L for the trailing-column solve, written by the CTA that owns the pivot row
            for (int q = 0; q < bw; ++q) {
                const int jg = jblk_s[q];
                if (jg >= row0 && jg < row0 + nrows)
                    for (int cc = lane; cc < b0 + q; cc += 32)
                        p.lmat[(size_t)(b0 + q) * CP_NB + cc] = panel[(size_t)cc * rpcp + (jg - row0)];
            }
        }
        if (glen > 0) {
            for (int c = t; c < glen; c += CP_THREADS) {
                double rr[B];
                double* gp = gbuf + (size_t)c * B;
#pragma unroll
                for (int q = 0; q < B; ++q) rr[q] = q < bw ? gp[q] : 0.0;
#pragma unroll
                for (int q = 1; q < B; ++q)
#pragma unroll
                    for (int qq = 0; qq < q; ++qq) rr[q] = fma(-lblk[q * B + qq], rr[qq], rr[q]);
#pragma unroll
                for (int q = 0; q < B; ++q) gp[q] = rr[q];
            }
            __syncthreads();                       // (B) R complete
            // all threads: row ri, columns ci, ci + tc, ...
            double uu[B];
            const bool okk = ri < nrows;
            bool dd = false;
#pragma unroll
            for (int q = 0; q < B; ++q) {
                uu[q] = (okk && q < bw) ? panel[(size_t)(b0 + q) * rpcp + ri] : 0.0;
                if (q < bw && okk && row0 + ri == jblk_s[q]) dd = true;
            }
            if (okk) {
                // CU_IL columns in flight per thread: their loads are issued together and the CU_IL chains of 8 dependent
                // FMAs interleave (one column at a time left the FP64 pipe waiting on a single chain)
                constexpr int CU_IL = 4;
                for (int c0 = ci; c0 < glen; c0 += CU_IL * tc) {
                    double rr[CU_IL][B], acc[CU_IL];
#pragma unroll
                    for (int u = 0; u < CU_IL; ++u) {
                        const bool live = c0 + u * tc < glen;
                        const int c = live ? c0 + u * tc : glen - 1;       // R of a clamped column is loaded, never used
                        const double2* gp = reinterpret_cast<const double2*>(gbuf + (size_t)c * B);
#pragma unroll
                        for (int q = 0; q < B; q += 2) {
                            const double2 v2 = gp[q / 2];
                            rr[u][q] = v2.x;
                            rr[u][q + 1] = v2.y;
                        }
                        acc[u] = live ? panel[(size_t)(b0 + bw + c) * rpcp + ri] : 0.0;
                    }
#pragma unroll
                    for (int q = 0; q < B; ++q)
#pragma unroll
                        for (int u = 0; u < CU_IL; ++u) acc[u] = fma(-uu[q], rr[u][q], acc[u]);
#pragma unroll
                    for (int u = 0; u < CU_IL; ++u) {
                        const int c = c0 + u * tc;
                        if (c < glen) panel[(size_t)(b0 + bw + c) * rpcp + ri] = dd ? 0.0 : acc[u];
                    }
                }
            }
            __syncthreads();                       // (C) the rest of the panel is up to date
        }
        CP_TICK(6)
    }
#undef CP_TICK
    __syncthreads();
    const bool stopped_all = stop_s != 0;
    if (pivot) {
#pragma unroll
        for (int m = 0; m < RPL; ++m)
            if (ok[m]) p.mu[row0 + rowi[m]] = mu[m];
    }
    if (p.write_u && !stopped_all) {
        for (int e = t; e < nb * rpcp; e += CP_THREADS) {
            const int c = e / rpcp, i = e - c * rpcp;
            if (i < nrows) p.basis[(size_t)(p.t0 + c) * S + row0 + i] = panel[e];
        }
    }
    if (r == 0 && t == 0) {
        if (stopped_all) p.state[0] = 1;
        p.state[1] = p.t0 + done;
        if (p.prof)
            for (int i = 0; i < 8; ++i) p.prof[i] += pa[i];
    }
    cluster.sync();                                // nobody leaves while a peer could still write into its smem
}

// -----------------------------------------------------------------------------------------------------
// R = L^-1 Phi[J, rest]: one warp per trailing column, lane l holds entries l and l + 32
// -----------------------------------------------------------------------------------------------------
constexpr int CS_WARPS = 8;
__global__ void __launch_bounds__(CS_WARPS * 32) car_panel_solve_kernel(const double* __restrict__ basis,
                                                                        const int* __restrict__ piv,
                                                                        const double* __restrict__ lmat,
                                                                        const int* __restrict__ state, int S, int k,
                                                                        int t0, int nb, double* __restrict__ Rt) {
    __shared__ double Lt[CP_NB][CP_NB + 1];        // Lt[c][s] = L[s][c]: lanes read consecutive s
    __shared__ int js[CP_NB];
    if (*(volatile const int*)state != 0) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int e = t; e < CP_NB * CP_NB; e += CS_WARPS * 32) {
        const int s = e / CP_NB, c = e % CP_NB;
        Lt[c][s] = (s < nb && c < s) ? lmat[e] : 0.0;
    }
    if (t < CP_NB) js[t] = t < nb ? piv[t0 + t] : 0;
    __syncthreads();
    const int col = t0 + nb + blockIdx.x * CS_WARPS + warp;
    if (col >= k) return;
    const double* src = basis + (size_t)col * S;
    double g0 = lane < nb ? __ldg(src + js[lane]) : 0.0;
    double g1 = lane + 32 < nb ? __ldg(src + js[lane + 32]) : 0.0;
    for (int s = 0; s < nb; ++s) {
        const double rs = __shfl_sync(0xffffffffu, s < 32 ? g0 : g1, s & 31);
        g0 = fma(-Lt[s][lane], rs, g0);            // Lt[s][l] = L[l][s] = 0 for l <= s: finished entries stay
        g1 = fma(-Lt[s][lane + 32], rs, g1);
    }
    double* dst = Rt + (size_t)(col - (t0 + nb)) * CP_NB;
    dst[lane] = g0;
    dst[lane + 32] = g1;
}

// -----------------------------------------------------------------------------------------------------
// Phi[:, rest] -= U R, zeros in the pivot rows.  64 rows x 64 columns per CTA, 4 x 4 register tile per thread.
// (A 128 x 64 tile with 8 x 4 registers took 31-34 us per launch whatever the grid: every CTA ran its load / FMA /
// read-modify-write phases one after the other with 8 warps -- latency, not throughput.  Four times as many CTAs of a
// quarter of the work each fill the GPU and the phases of different CTAs overlap on an SM.)
// -----------------------------------------------------------------------------------------------------
constexpr int CU_TI = 64, CU_TCOL = 64;
__global__ void __launch_bounds__(256, 2) car_panel_update_kernel(double* __restrict__ basis,
                                                                  const int* __restrict__ piv,
                                                                  const double* __restrict__ Rt,
                                                                  const int* __restrict__ state, int S, int k, int t0,
                                                                  int nb, int c_begin, int c_end) {
    extern __shared__ __align__(16) double su[];
    double* Us = su;                                // [CP_NB][CU_TI]
    double* Rs = su + CP_NB * CU_TI;                // [CU_TCOL][CP_NB + 1]
    __shared__ unsigned char is_piv[CU_TI];
    if (*(volatile const int*)state != 0) return;
    const int t = threadIdx.x;
    const int i0 = blockIdx.x * CU_TI;
    const int c0 = c_begin + blockIdx.y * CU_TCOL;   // columns [c_begin, c_end) of the trailing block
    k = min(k, c_end);
    if (t < CU_TI) is_piv[t] = 0;
    for (int e = t; e < CP_NB * CU_TI; e += 256) {
        const int s = e / CU_TI, i = e % CU_TI;
        Us[e] = (s < nb && i0 + i < S) ? basis[(size_t)(t0 + s) * S + i0 + i] : 0.0;
    }
    for (int e = t; e < CU_TCOL * CP_NB; e += 256) {
        const int c = e / CP_NB, s = e % CP_NB;
        Rs[c * (CP_NB + 1) + s] = (c0 + c < k && s < nb) ? Rt[(size_t)(c0 + c - (t0 + nb)) * CP_NB + s] : 0.0;
    }
    __syncthreads();
    if (t < nb) {
        const int j = piv[t0 + t] - i0;
        if (j >= 0 && j < CU_TI) is_piv[j] = 1;
    }
    // thread (ti, tcx): rows i0 + 32 q + 2 ti + {0, 1} (q < 2: consecutive lanes read consecutive 16-byte chunks of U),
    // columns 4 tcx + jj
    const int ti = t & 15, tcx = t >> 4;
    double acc[4][4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) acc[jj][ii] = 0.0;
#pragma unroll 8
    for (int s = 0; s < CP_NB; ++s) {
        double a[4], b[4];
        const double2* ap = reinterpret_cast<const double2*>(Us + s * CU_TI + ti * 2);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double2 v = ap[q * 16];
            a[2 * q] = v.x;
            a[2 * q + 1] = v.y;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) b[jj] = Rs[(tcx * 4 + jj) * (CP_NB + 1) + s];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) acc[jj][ii] = fma(a[ii], b[jj], acc[jj][ii]);
    }
    __syncthreads();                               // is_piv complete
    double old[4][4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int c = c0 + tcx * 4 + jj;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int il = (ii >> 1) * 32 + ti * 2 + (ii & 1);
            old[jj][ii] = (c < k && i0 + il < S) ? basis[(size_t)c * S + i0 + il] : 0.0;
        }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int c = c0 + tcx * 4 + jj;
        if (c >= k) continue;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int il = (ii >> 1) * 32 + ti * 2 + (ii & 1);
            if (i0 + il < S) basis[(size_t)c * S + i0 + il] = is_piv[il] ? 0.0 : old[jj][ii] - acc[jj][ii];
        }
    }
}

struct PanelPlan {
    int rpl, npw, rpc, rpcp, tr, nb;
    bool single;
    size_t smem(int nb_) const {
        return ((size_t)nb_ * rpcp + 2 * (size_t)CP_P * CP_HB + (size_t)nb_ * CP_B + CP_B * CP_B + 2 * 4 * CP_HB) * 8;
    }
};

static bool plan_panel(int S, int k, int nb_hint, PanelPlan* pl) {
    if (S <= 0 || k <= 0 || k >= S + 1) return false;
    pl->rpc = (S + CP_P - 1) / CP_P;
    if (pl->rpc > CP_THREADS) return false;             // S <= 2048: four pivot warps, two rows per lane
    pl->rpl = pl->rpc <= 32 ? 1 : 2;
    pl->npw = pl->rpc <= 64 ? 1 : (pl->rpc <= 128 ? 2 : 4);
    int tr = 32;
    while (tr < pl->rpc && tr < CP_THREADS) tr *= 2;
    pl->tr = tr;
    pl->rpcp = (pl->rpc + 1) & ~1;
    if (pl->rpcp % 32 == 0) pl->rpcp += 2;
    pl->single = nb_hint <= 0 && pl->smem(k) <= (size_t)CP_SMEM_MAX;
    if (pl->single) {
        pl->nb = k;
        return true;
    }
    int nb = nb_hint > 0 ? nb_hint : CP_NB;
    if (nb > CP_NB) nb = CP_NB;
    while (nb > 8 && pl->smem(nb) > (size_t)CP_SMEM_MAX) nb -= 8;
    if (pl->smem(nb) > (size_t)CP_SMEM_MAX) return false;
    pl->nb = nb;
    return true;
}

// helper stream + events of the lookahead, one set per device, created on first use and never destroyed
struct Lookahead {
    cudaStream_t stream = nullptr;
    cudaEvent_t solved = nullptr, rest_done = nullptr;
    bool ok = false, tried = false;
};
static Lookahead& lookahead_for_device() {
    static Lookahead table[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    Lookahead& la = table[dev];
    if (!la.tried) {
        la.tried = true;
        la.ok = cudaStreamCreateWithFlags(&la.stream, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.solved, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&la.rest_done, cudaEventDisableTiming) == cudaSuccess;
        if (!la.ok) (void)cudaGetLastError();
    }
    return la;
}

}  // namespace sober

using namespace sober;

extern "C" int sober_car_panel_fits(int32_t S, int32_t k) {
    PanelPlan pl;
    return plan_panel(S, k, 0, &pl) ? (pl.single ? 1 : 2) : 0;
}

extern "C" int64_t sober_car_panel_workspace(int32_t S, int32_t k) {
    // state (2 int32, padded) | pivots (k int32, padded to 8 bytes) | L (64 x 64) | Rt (k x 64)
    if (S <= 0 || k <= 0) return -1;
    return 16 + (((int64_t)k * 4 + 15) & ~15ll) + (int64_t)CP_NB * CP_NB * 8 + (int64_t)k * CP_NB * 8;
}

extern "C" int sober_car_panel_profiled(double* basis, int32_t k, int32_t S, double* mu, int32_t nb_hint, int32_t* info,
                                        void* workspace, int64_t workspace_bytes, int64_t* prof, void* stream) {
    if (!basis || !mu || !workspace) return SOBER_ERR_ARG;
    PanelPlan pl;
    if (!plan_panel(S, k, nb_hint, &pl)) return SOBER_ERR_UNSUPPORTED;
    if (workspace_bytes < sober_car_panel_workspace(S, k)) return SOBER_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int* state = (int*)ws;
    int* piv = (int*)(ws + 16);
    double* lmat = (double*)(ws + 16 + (((int64_t)k * 4 + 15) & ~15ll));
    double* Rt = lmat + CP_NB * CP_NB;
    SOBER_CUDA_CHECK(cudaMemsetAsync(state, 0, 16, st));

    void (*kern)(const CarPanelParams) = pl.rpl == 1 ? car_panel_kernel<1, 1>
                                         : (pl.npw == 1 ? car_panel_kernel<2, 1>
                                                        : (pl.npw == 2 ? car_panel_kernel<2, 2> : car_panel_kernel<2, 4>));
    const int kvar = pl.rpl == 1 ? 0 : (pl.npw == 1 ? 1 : (pl.npw == 2 ? 2 : 3));
    const size_t smem_max = pl.smem(pl.single ? k : pl.nb);
    {   // the > 48 KB opt-in is per device and sticky: once per (device, kernel variant), never inside a graph capture
        static int configured[64][4] = {};
        static int configured_upd[64] = {};
        int dev = 0;
        SOBER_CUDA_CHECK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || configured[dev][kvar] < (int)smem_max) {
            SOBER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CP_SMEM_MAX));
            if (dev >= 0 && dev < 64) configured[dev][kvar] = CP_SMEM_MAX;
        }
        if (dev < 0 || dev >= 64 || !configured_upd[dev]) {
            SOBER_CUDA_CHECK(cudaFuncSetAttribute(car_panel_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  (int)((CP_NB * CU_TI + CU_TCOL * (CP_NB + 1)) * 8)));
            if (dev >= 0 && dev < 64) configured_upd[dev] = 1;
        }
    }
    Lookahead& la = lookahead_for_device();
    bool pending = false;
    for (int t0 = 0; t0 < k; t0 += pl.nb) {
        const int nb = (k - t0) < pl.nb ? (k - t0) : pl.nb;
        const bool trailing = t0 + nb < k;
        CarPanelParams p;
        p.basis = basis; p.mu = mu; p.state = state; p.piv = piv; p.lmat = lmat; p.prof = (long long*)prof;
        p.S = S; p.k = k; p.t0 = t0; p.nb = nb;
        p.rpc = pl.rpc; p.rpcp = pl.rpcp; p.tr = pl.tr;
        p.write_u = trailing ? 1 : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(CP_P);
        cfg.blockDim = dim3(CP_THREADS);
        cfg.dynamicSmemBytes = pl.smem(nb);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CP_P;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SOBER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
        if (trailing) {
            // LOOKAHEAD: only the next panel's columns must be up to date before the next panel kernel starts; the
            // rest of the trailing update runs on a helper stream beside it (8 CTAs vs the remaining 140 SMs) and is
            // joined before the next solve, which reads the pivot rows of ALL trailing columns.
            const int rest = k - (t0 + nb);
            const size_t usm = (CP_NB * CU_TI + CU_TCOL * (CP_NB + 1)) * 8;
            if (pending) {
                SOBER_CUDA_CHECK(cudaStreamWaitEvent(st, la.rest_done, 0));
                pending = false;
            }
            car_panel_solve_kernel<<<(unsigned)ceil_div(rest, CS_WARPS), CS_WARPS * 32, 0, st>>>(basis, piv, lmat, state,
                                                                                                   S, k, t0, nb, Rt);
            SOBER_LAUNCH_CHECK("car_panel_solve");
            const int first_end = (t0 + 2 * nb < k) ? t0 + 2 * nb : k;
            dim3 grid1((unsigned)ceil_div(S, CU_TI), (unsigned)ceil_div(first_end - (t0 + nb), CU_TCOL));
            if (first_end < k && la.ok) {
                SOBER_CUDA_CHECK(cudaEventRecord(la.solved, st));
                SOBER_CUDA_CHECK(cudaStreamWaitEvent(la.stream, la.solved, 0));
                dim3 grid2((unsigned)ceil_div(S, CU_TI), (unsigned)ceil_div(k - first_end, CU_TCOL));
                car_panel_update_kernel<<<grid2, 256, usm, la.stream>>>(basis, piv, Rt, state, S, k, t0, nb, first_end, k);
                SOBER_LAUNCH_CHECK("car_panel_update");
                SOBER_CUDA_CHECK(cudaEventRecord(la.rest_done, la.stream));
                pending = true;
                car_panel_update_kernel<<<grid1, 256, usm, st>>>(basis, piv, Rt, state, S, k, t0, nb, t0 + nb, first_end);
            } else {
                dim3 grid((unsigned)ceil_div(S, CU_TI), (unsigned)ceil_div(rest, CU_TCOL));
                car_panel_update_kernel<<<grid, 256, usm, st>>>(basis, piv, Rt, state, S, k, t0, nb, t0 + nb, k);
            }
            SOBER_LAUNCH_CHECK("car_panel_update");
        }
    }
    if (pending) SOBER_CUDA_CHECK(cudaStreamWaitEvent(st, la.rest_done, 0));
    if (info) SOBER_CUDA_CHECK(cudaMemcpyAsync(info, state, 8, cudaMemcpyDeviceToDevice, st));
    return SOBER_OK;
}

extern "C" int sober_car_panel(double* basis, int32_t k, int32_t S, double* mu, int32_t nb_hint, int32_t* info,
                               void* workspace, int64_t workspace_bytes, void* stream) {
    return sober_car_panel_profiled(basis, k, S, mu, nb_hint, info, workspace, workspace_bytes, nullptr, stream);
}
