// K3 (cluster variant) -- the whole Caratheodory reduction of SOBER/_rchq.py:224-270 in ONE kernel on one
// thread-block cluster, all state resident in distributed shared memory:
//
//   phase 1  Householder QR of the design matrix  D = [1 | X]  (S x n')        (replaces the full SVD of :231)
//   phase 2  Q2 = H_1 ... H_n' [0; I_k]  -- an orthonormal basis of null(D^T)   (replaces Vh[-(N-n):] of :234)
//   phase 3  the k dependent elimination steps of :237-266 on Phi = Q2
//
// or, when the caller supplies a basis (parity mode: the rows of torch.linalg.svd's Vh), phase 3 alone.
//
// Why a cluster: the work is tiny (~1e8 flop at b = 200) but strictly sequential -- n' + n' + k steps, each needing
// every row of the matrix.  Through L2 a step costs ~3 us (flag + fence + reload, csrc/car_eliminate.cu); inside a
// cluster an exchange is a DSMEM store plus barrier.cluster (~0.2 us).  Rows are distributed cyclically over the P
// CTAs of the cluster (row i lives in CTA i mod P, so the shrinking active set stays balanced); a step exchanges
// only P partial vectors of length <= max(n', k) (phases 1-2) or the pivot row (phase 3).  Every CTA reduces the
// partials in rank order, so all CTAs hold bit-identical scalars/vectors and no second exchange is needed.
//
// Reflector convention is LAPACK's (dlarfg: beta = -sign(alpha) * norm, v_c = 1), so Q2 equals the trailing columns
// of the complete Householder Q that LAPACK / cuSOLVER produce, up to rounding.
// Elimination arithmetic: EXACT = the reference's unfused (phi_j * v_i) / v_j; otherwise one division per row
// (t_i = v_i / v_j) and an FMA per element.
#include "common.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace sober {

constexpr int CC_THREADS = 256;   // one thread per matrix column: n' <= 256 and k <= 256
constexpr int CC_P = 8;           // CTAs per cluster (portable maximum)

struct CarClusterParams {
    const double* design;  // S x np row-major (ld = np); unused when basis != nullptr
    const double* basis;   // k x S rows (row c = column c of Phi) or nullptr
    double* mu;            // S, in/out
    int* info;             // [0] = elimination steps taken, [1] = 1 if phases 1-2 ran
    long long* prof;       // optional: 16 cycle counters of CTA 0 / thread 0 (diagnostics)
    int S, np, k;
    int rows_max;          // ceil(S / P)
    int W;                 // max(np, k) rounded up to even
};

__device__ __forceinline__ bool ratio_better(double ra, int ia, double rb, int ib) {
    // is (rb, ib) better than (ra, ia)?  first minimum in index order
    if (ib < 0) return false;
    if (ia < 0) return true;
    return rb < ra || (rb == ra && ib < ia);
}

// ---- DSMEM exchange primitives.  A round = every CTA stages its contribution in its own shared memory, then P
// threads push it to the P peers with ONE bulk copy each (cp.async.bulk shared::cta -> shared::cluster) that completes
// a transaction count on the RECEIVER's mbarrier; the receiver posted the bytes it expects and sleeps on that
// mbarrier.  No cluster-wide barrier and no fence on the critical path, one mbarrier update per peer.
__device__ __forceinline__ uint32_t map_rank(uint32_t saddr, uint32_t rank) {
    uint32_t out;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(saddr), "r"(rank));
    return out;
}
__device__ __forceinline__ void bulk_s2c(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dst_cluster),
                 "r"(src_cta), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote_addr),
                 "d"(v), "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Dynamic (CTA-uniform) row access into the register-resident column: a jump table, not RMAX selects.
template <int RMAX>
__device__ __forceinline__ double col_get(const double (&col)[RMAX], int li) {
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
__device__ __forceinline__ void col_zero(double (&col)[RMAX], int li) {
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

// Matrix columns live in REGISTERS: thread j holds column j of its CTA's rows (col[RMAX]); per step only the pivot
// column / Householder vector is broadcast through shared memory, and every inner loop is a branch-free run of
// RMAX (LDS, DFMA) pairs: rows that must not take part are ZERO in the broadcast vector instead of being masked
// (finished rows of the QR are zeroed in the registers -- R itself is never needed -- and eliminated rows of Phi are
// zero by construction).  Measured history of this kernel at S = 400, n' = 200 (clock64 counters, tools/
// profile_car_cluster.py): matrices in shared memory 3.3 us/step; registers + per-row predicates 2.8 us/step
// (2/3 of the issued instructions were ISETP/FSEL); this version: see profiles/.
template <int RMAX, bool EXACT>
__global__ void __launch_bounds__(CC_THREADS, 1) car_cluster_kernel(const CarClusterParams p) {
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int P = CC_P;
    const int r = (int)cluster.block_rank();
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int S = p.S, np = p.np, k = p.k, W = p.W;
    const int Wx = W + 2;                             // slot width: W values + (ratio, idx)

    extern __shared__ __align__(16) double sm[];
    double* xw = sm;                                  // 2 x (P + 1) x Wx  received vectors (+ owner-row slot)
    double* stage = xw + (size_t)2 * (P + 1) * Wx;    // 2 x 2 x Wx        what this CTA sends (vector, owner row)
    double* pc = stage + (size_t)4 * Wx;              // 2 x RMAX          pivot column of my rows (double-buffered)
    double* vb = pc + 2 * RMAX;                       // RMAX              Householder vector of my rows (phase 1)
    double* mu_l = vb + RMAX;                         // RMAX
    double* tau = mu_l + RMAX;                        // np
    double* V = tau + ((np + 1) & ~1);                // RMAX x np         reflectors: v_i (i > c), 1 (i == c), 0 (i < c)
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ int s_best;                            // local row index of this CTA's pivot candidate

    const int nloc = (S - r + P - 1) / P;             // my rows: i = li * P + r   (nloc <= RMAX)
    auto slot = [&](int buf, int q) { return xw + ((size_t)buf * (P + 1) + q) * Wx; };   // q == P: owner-row slot
    auto stg = [&](int buf, int which) { return stage + ((size_t)buf * 2 + which) * Wx; };

    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    double col[RMAX];
    const bool qr = p.basis == nullptr;
    long long pr_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long pr_t = 0;
    const bool prof = p.prof != nullptr && r == 0 && t == 0;
#define CC_TICK(slot_)                                   \
    if (prof) {                                          \
        const long long now_ = clock64();                \
        pr_acc[slot_] += now_ - pr_t;                    \
        pr_t = now_;                                     \
    }
    // ---------------------------------------------------------------------------------------------- load
#pragma unroll
    for (int li = 0; li < RMAX; ++li) {
        const int i = li * P + r;
        double v = 0.0;
        if (li < nloc) {
            if (qr) { if (t < np) v = p.design[(size_t)i * np + t]; }
            else { if (t < k) v = p.basis[(size_t)t * S + i]; }
        }
        col[li] = v;
    }
    for (int li = t; li < RMAX; li += CC_THREADS) mu_l[li] = li < nloc ? p.mu[li * P + r] : 0.0;
    for (int i = t; i < 4 * Wx; i += CC_THREADS) stage[i] = 0.0;
    for (int i = t; i < 3 * RMAX; i += CC_THREADS) pc[i] = 0.0;      // pc (2 x RMAX) and vb
    __syncthreads();
    cluster.sync();   // barriers initialised everywhere before the first remote copy

    const uint32_t bar_addr[2] = {smem_addr(&bars[0]), smem_addr(&bars[1])};
    uint32_t phase_bit0 = 0, phase_bit1 = 0;
    int round = 0;
    // push stage(buf, which)[lo, hi) to slot(buf, dst_slot) of every CTA; lo, hi even (16-byte granules)
    auto push = [&](int buf, int which, int dst_slot, int lo, int hi) {
        if (t < P) {
            bulk_s2c(map_rank(smem_addr(slot(buf, dst_slot) + lo), t), smem_addr(stg(buf, which) + lo),
                     (uint32_t)(hi - lo) * 8u, map_rank(bar_addr[buf], t));
        }
    };
    auto send_all = [&](int buf, const double* local_dst, double v) {   // per-element variant: same offset in every CTA
        const uint32_t a = smem_addr(local_dst);
#pragma unroll
        for (int q = 0; q < P; ++q) st_async_f64(map_rank(a, q), v, map_rank(bar_addr[buf], q));
    };
    auto wait_round = [&](int buf) {
        if (buf == 0) { mbar_wait(&bars[0], phase_bit0); phase_bit0 ^= 1; }
        else { mbar_wait(&bars[1], phase_bit1); phase_bit1 ^= 1; }
    };

    if (qr) {
        // ------------------------------------------------------------------------------------ phase 1: QR
        // One exchange per column c: every CTA sends the partial dot products d_j = sum_{my rows i > c} A[i,c] A[i,j]
        // (j = c is the sum of squares), the owner of row c also sends that row.  Everybody then forms
        // w_j = A[c,j] + scale * d_j  (= v^T A[:, j] with v = [1; scale * A[c+1:, c]]) and applies
        // A[:, j] -= tau w_j v.  Rows <= c are then dead (R is not needed): the owner zeroes row c in its registers.
        if (t == 0) {
#pragma unroll
            for (int li = 0; li < RMAX; ++li) pc[li] = col[li];
            if (r == 0) pc[0] = 0.0;                           // the diagonal entry travels in the owner-row slot
        }
        for (int c = 0; c < np; ++c) {
            const int buf = round & 1;
            ++round;
            const int lc = c / P;                             // local index of row c in its owner
            const int lo = c & ~1;
            const bool own = (c % P == r);
            const double* piv = pc + (c & 1) * RMAX;           // column c of my rows, zero for rows <= c
            if (prof && c == 0) pr_t = clock64();
#ifdef SOBER_CC_STASYNC
            if (t == 0) mbar_expect_tx(&bars[buf], (uint32_t)(P + 1) * (uint32_t)(np - c) * 8u);
#else
            if (t == 0) mbar_expect_tx(&bars[buf], (uint32_t)(P + 1) * (uint32_t)(W - lo) * 8u);
#endif
            __syncthreads();   // pivot column c is in piv
            CC_TICK(0)
            if (t >= c && t < np) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int li = 0; li + 3 < RMAX; li += 4) {
                    s0 = fma(piv[li], col[li], s0);
                    s1 = fma(piv[li + 1], col[li + 1], s1);
                    s2 = fma(piv[li + 2], col[li + 2], s2);
                    s3 = fma(piv[li + 3], col[li + 3], s3);
                }
#ifdef SOBER_CC_STASYNC
                send_all(buf, slot(buf, r) + t, (s0 + s1) + (s2 + s3));
                if (own) send_all(buf, slot(buf, P) + t, col_get<RMAX>(col, lc));
            }
            CC_TICK(1)
#else
                stg(buf, 0)[t] = (s0 + s1) + (s2 + s3);
                if (own) stg(buf, 1)[t] = col_get<RMAX>(col, lc);
            }
            CC_TICK(1)
            fence_proxy_async_smem();   // writers make their generic-proxy stores visible to the bulk-copy engine
            __syncthreads();
            push(buf, 0, r, lo, W);
            if (own) push(buf, 1, P, lo, W);
#endif
            CC_TICK(2)
            wait_round(buf);
            CC_TICK(3)
            double ssq = 0.0;
#pragma unroll
            for (int q = 0; q < P; ++q) ssq += slot(buf, q)[c];
            const double alpha = slot(buf, P)[c];
            double beta = alpha, tc = 0.0, scale = 0.0;
            if (ssq != 0.0) {
                beta = -copysign(sqrt(fma(alpha, alpha, ssq)), alpha);
                tc = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            if (t < RMAX) {
                // the Householder vector of my rows: scale * x_i (i > c), 1 (i == c); also kept for phase 2
                const double v = (own && t == lc) ? 1.0 : scale * piv[t];
                vb[t] = v;
                V[(size_t)t * np + c] = v;
                if (t == 0) tau[c] = tc;
            }
            __syncthreads();
            if (t > c && t < np) {
                double d = 0.0;
#pragma unroll
                for (int q = 0; q < P; ++q) d += slot(buf, q)[t];
                const double w = fma(scale, d, slot(buf, P)[t]);
                const double ntw = -(tc * w);
#pragma unroll
                for (int li = 0; li < RMAX; ++li) col[li] = fma(ntw, vb[li], col[li]);
                if (own) col_zero<RMAX>(col, lc);              // row c is finished
                if (t == c + 1) {   // next pivot column is final: publish it (other buffer), diagonal entry zeroed
                    double* nxt = pc + ((c + 1) & 1) * RMAX;
#pragma unroll
                    for (int li = 0; li < RMAX; ++li) nxt[li] = col[li];
                    if ((c + 1) % P == r) nxt[(c + 1) / P] = 0.0;
                }
            }
            CC_TICK(4)
        }
        __syncthreads();
        // ------------------------------------------------------------------------------------ phase 2: Q2
        // Q <- H_c Q for c = np-1 .. 0, Q = [0; I_k] initially; thread m holds column m of Q.
#pragma unroll
        for (int li = 0; li < RMAX; ++li) col[li] = (li < nloc && (li * P + r) - np == t) ? 1.0 : 0.0;
        for (int c = np - 1; c >= 0; --c) {
            const int buf = round & 1;
            ++round;
            const double ntc = -tau[c];
            if (prof && c == np - 1) pr_t = clock64();
#ifdef SOBER_CC_STASYNC
            if (t == 0) mbar_expect_tx(&bars[buf], (uint32_t)P * (uint32_t)k * 8u);
#else
            if (t == 0) mbar_expect_tx(&bars[buf], (uint32_t)P * (uint32_t)W * 8u);
#endif
            if (t < k) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int li = 0; li + 3 < RMAX; li += 4) {
                    s0 = fma(V[(size_t)li * np + c], col[li], s0);
                    s1 = fma(V[(size_t)(li + 1) * np + c], col[li + 1], s1);
                    s2 = fma(V[(size_t)(li + 2) * np + c], col[li + 2], s2);
                    s3 = fma(V[(size_t)(li + 3) * np + c], col[li + 3], s3);
                }
#ifdef SOBER_CC_STASYNC
                send_all(buf, slot(buf, r) + t, (s0 + s1) + (s2 + s3));
            }
            CC_TICK(5)
#else
                stg(buf, 0)[t] = (s0 + s1) + (s2 + s3);
            }
            CC_TICK(5)
            fence_proxy_async_smem();
            __syncthreads();
            push(buf, 0, r, 0, W);
#endif
            wait_round(buf);
            CC_TICK(6)
            if (t < k) {
                double w = 0.0;
#pragma unroll
                for (int q = 0; q < P; ++q) w += slot(buf, q)[t];
                const double ntw = ntc * w;
#pragma unroll
                for (int li = 0; li < RMAX; ++li) col[li] = fma(ntw, V[(size_t)li * np + c], col[li]);
            }
            CC_TICK(7)
        }
    }

    // ---------------------------------------------------------------------------------------- phase 3: eliminate
    // One exchange per step: every CTA sends its best (ratio, index) AND, speculatively, that candidate's row; the
    // winner's row is then already local everywhere.  Thread m holds column m of Phi.
    __syncthreads();
    if (t == 0) {
#pragma unroll
        for (int li = 0; li < RMAX; ++li) pc[li] = col[li];
    }
    int done = 0;
    for (int s = 0; s < k; ++s) {
        const int buf = round & 1;
        ++round;
        const int lo = s & ~1;
        const double* piv = pc + (s & 1) * RMAX;
        if (prof && s == 0) pr_t = clock64();
        __syncthreads();   // pivot column s is in piv; mu_l of the previous step complete
        if (warp == 0) {
            double br = 0.0;
            int bi = -1;
            for (int li = lane; li < nloc; li += 32) {
                const double v = piv[li];
                if (v > 0.0) {
                    const double ratio = __ddiv_rn(mu_l[li], v);
                    const int i = li * P + r;
                    if (ratio_better(br, bi, ratio, i)) { br = ratio; bi = i; }
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double orr = __shfl_xor_sync(0xffffffffu, br, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ratio_better(br, bi, orr, oi)) { br = orr; bi = oi; }
            }
            if (lane == 0) {
                s_best = bi < 0 ? -1 : bi / P;
#ifdef SOBER_CC_STASYNC
                mbar_expect_tx(&bars[buf], (uint32_t)P * (uint32_t)(k - s + 2) * 8u);
                send_all(buf, slot(buf, r) + W, br);
                send_all(buf, slot(buf, r) + W + 1, (double)bi);
#else
                mbar_expect_tx(&bars[buf], (uint32_t)P * (uint32_t)(Wx - lo) * 8u);
                stg(buf, 0)[W] = br;
                stg(buf, 0)[W + 1] = (double)bi;
#endif
            }
        }
        __syncthreads();
        CC_TICK(8)
#ifdef SOBER_CC_STASYNC
        if (t >= s && t < k) send_all(buf, slot(buf, r) + t, col_get<RMAX>(col, s_best));
#else
        if (t >= s && t < k) stg(buf, 0)[t] = col_get<RMAX>(col, s_best);
        fence_proxy_async_smem();
        __syncthreads();
        push(buf, 0, r, lo, Wx);
#endif
        CC_TICK(9)
        wait_round(buf);
        CC_TICK(10)
        double alpha = 0.0;
        int j = -1, qwin = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const double rq = slot(buf, q)[W];
            const int iq = (int)slot(buf, q)[W + 1];
            if (ratio_better(alpha, j, rq, iq)) { alpha = rq; j = iq; qwin = q; }
        }
        if (j < 0) break;   // no positive entry anywhere: the guard of SOBER/_rchq.py:241-242 (cluster-uniform)
        done = s + 1;
        const double* prow = slot(buf, qwin);
        const double vj = prow[s];
        const int lj = (j % P == r) ? j / P : -1;
        if (t > s && t < k) {
            const double pr = prow[t];
            if (EXACT) {
#pragma unroll
                for (int li = 0; li < RMAX; ++li) col[li] = __dsub_rn(col[li], __ddiv_rn(__dmul_rn(pr, piv[li]), vj));
            } else {
                const double nf = -__ddiv_rn(pr, vj);
#pragma unroll
                for (int li = 0; li < RMAX; ++li) col[li] = fma(nf, piv[li], col[li]);
            }
            if (lj >= 0) col_zero<RMAX>(col, lj);
            if (t == s + 1) {
                double* nxt = pc + ((s + 1) & 1) * RMAX;
#pragma unroll
                for (int li = 0; li < RMAX; ++li) nxt[li] = col[li];
            }
        }
        for (int li = t; li < nloc; li += CC_THREADS) {        // weights of my rows: mu - alpha * v, unfused
            const double v = piv[li];
            mu_l[li] = (li == lj) ? 0.0 : __dsub_rn(mu_l[li], __dmul_rn(alpha, v));
        }
        CC_TICK(11)
    }
    __syncthreads();
    if (prof)
        for (int i = 0; i < 12; ++i) p.prof[i] = pr_acc[i];
    for (int li = t; li < nloc; li += CC_THREADS) p.mu[li * P + r] = mu_l[li];
    if (r == 0 && t == 0 && p.info) {
        p.info[0] = done;
        p.info[1] = qr ? 1 : 0;
    }
    cluster.sync();   // no CTA may exit while peers can still address its shared memory
}

static size_t cluster_smem_bytes(int S, int np, int k, int rmax, int* rows_max_out, int* W_out) {
    const int P = CC_P;
    const int rows_max = (S + P - 1) / P;
    int W = np > k ? np : k;
    W = (W + 1) / 2 * 2;
    if (rows_max_out) *rows_max_out = rows_max;
    if (W_out) *W_out = W;
    const size_t Wx = (size_t)W + 2;
    size_t doubles = 2 * (size_t)(P + 1) * Wx + 4 * Wx + 4 * (size_t)rmax + ((np + 1) & ~1) + (size_t)rmax * np;
    return doubles * 8 + 64;
}

static int pick_rmax(int S, int np, int k, bool have_basis) {
    // register tile height (max local rows) the shape needs; 0 = the cluster kernel does not cover it
    if (k > CC_THREADS || (!have_basis && np > CC_THREADS)) return 0;
    const int rows = (S + CC_P - 1) / CC_P;
    const int rmax = rows <= 32 ? 32 : (rows <= 56 ? 56 : 0);
    if (rmax == 0) return 0;
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    if (cluster_smem_bytes(S, have_basis ? 0 : np, k, rmax, nullptr, nullptr) > (size_t)max_smem) return 0;
    return rmax;
}

}  // namespace sober

using namespace sober;

extern "C" int sober_car_cluster_fits(int32_t S, int32_t np, int32_t have_basis) {
    if (S <= 0 || np <= 0 || np >= S) return 0;
    return pick_rmax(S, np, S - np, have_basis != 0) ? CC_P : 0;
}

extern "C" int sober_car_cluster_profiled(const double* design, const double* basis, int32_t S, int32_t np, double* mu,
                                          int32_t exact, int32_t* info, int64_t* prof, void* stream);

extern "C" int sober_car_cluster(const double* design, const double* basis, int32_t S, int32_t np, double* mu,
                                 int32_t exact, int32_t* info, void* stream) {
    return sober_car_cluster_profiled(design, basis, S, np, mu, exact, info, nullptr, stream);
}

extern "C" int sober_car_cluster_profiled(const double* design, const double* basis, int32_t S, int32_t np, double* mu,
                                          int32_t exact, int32_t* info, int64_t* prof, void* stream) {
    if (S <= 0 || np <= 0 || np >= S || !mu || (!design && !basis)) return SOBER_ERR_ARG;
    const int k = S - np;
    const bool have_basis = basis != nullptr;
    const int rmax = pick_rmax(S, np, k, have_basis);
    if (rmax == 0) return SOBER_ERR_UNSUPPORTED;
    CarClusterParams p;
    p.design = design; p.basis = basis; p.mu = mu; p.info = info; p.prof = (long long*)prof;
    p.S = S; p.np = have_basis ? 0 : np; p.k = k;
    const size_t smem = cluster_smem_bytes(S, p.np, k, rmax, &p.rows_max, &p.W);
    void (*kern)(const CarClusterParams) = nullptr;
    if (rmax == 32) kern = exact ? car_cluster_kernel<32, true> : car_cluster_kernel<32, false>;
    else kern = exact ? car_cluster_kernel<56, true> : car_cluster_kernel<56, false>;
    SOBER_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CC_P);
    cfg.blockDim = dim3(CC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CC_P;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SOBER_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
    return SOBER_OK;
}
