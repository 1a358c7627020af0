// K1 for Tanimoto on bit-packed fingerprints, on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// <x, z> over {0,1}^d is an integer GEMM: 0/1 operands are exact in 8 bits and the counts (<= 2048) in the int32
// accumulators, so `tcgen05.mma.kind::i8` computes popcount(x & z) EXACTLY -- the same integers as the popcount kernel
// (group_bits_kernel), whose 16 POPC per clock per SM bound C4 at 84 ms per step.  Replaces SOBER/_rchq.py:124-136 for
// SOBER/_drug_modelling.py:15-25 exactly like the other K1 kernels (same contract: At, totw, the remainder quirk).
//
// Two kernels.  The DEFAULT is version 2 (group_bits_mma2_kernel, second half of this file: landmark tile resident in
// TMEM as the A operand, loader / expander / MMA / epilogue warps).  Version 1 below (landmark tile resident in shared
// memory) is kept for comparison (variant 5):
//
// CTA = 128 groups (one candidate per group and row) x 64 landmarks.  Warp roles:
//   warps 0-3  EXPANDERS  thread r owns candidate row r of the current row: gathers its bit-packed words from HBM
//                         (128 B per 1024-bit fingerprint: HBM traffic stays that of the packed format) and expands them
//                         to 0/1 bytes straight into the shared-memory operand tile of the MMA (canonical K-major,
//                         no-swizzle layout: 16-byte K-chunks, rows 16 bytes apart), 256 K-elements per pipeline stage;
//   warp  12   MMA        one thread issues 8 x tcgen05.mma (M=128, N=64, K=32) per stage into one of two accumulator
//                         stages in TMEM (64 int32 columns each), tcgen05.commit releases the stage / publishes the tile;
//   warps 4-11 EPILOGUE   (two warps per TMEM lane quarter, 32 columns each) tcgen05.ld of the lane's dot products, FP64 Tanimoto ratio (the arithmetic of
//                         tanimoto_value(), common.cuh), weighted accumulation into 32 FP64 registers per thread.
// The landmark tile (64 x d bytes) is expanded once per CTA and stays resident in shared memory.
#include "common.cuh"

namespace sober {

constexpr int BM_TM = 128;        // candidates (groups) per tile = TMEM lanes
constexpr int BM_TN = 64;         // landmarks per CTA = accumulator columns
constexpr int BM_KB = 256;        // K elements (bits -> bytes) per pipeline stage
constexpr int BM_STAGES = 3;
constexpr int BM_THREADS = 416;   // 4 expander warps + 8 epilogue warps (two per TMEM lane quarter) + 1 MMA warp
constexpr int BM_MMA_WARP = 12;
constexpr int BM_MAXW = 16;       // words per row (1024 bits); wider rows use the popcount kernel

struct BitsMmaParams {
    const uint64_t* X;            // candidate words, row stride W
    const uint64_t* Z;            // landmark words, row stride W
    const double* xn;             // candidate popcounts
    int64_t xn_stride;
    const double* zn;             // landmark popcounts
    const int32_t* idx;
    const double* mu;
    int64_t n_local, pos0, ES;
    int S, L, W;
    double* out;                  // [nsplit][S][L]
    double* totw_out;             // [nsplit][S]
    int64_t row_begin, row_end, rows_per_split;
    double scale;
};

__device__ __forceinline__ void bm_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bm_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bm_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar))
                 : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes (rows 16 B apart);
// LBO = byte distance of the two 16-byte K-chunks of one MMA, SBO = byte distance of consecutive 8-row groups
__device__ __forceinline__ uint64_t bm_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void bm_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bm_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 bits -> 16 bytes of 0/1 without a table: a nibble n becomes (n * 0x00204081) & 0x01010101 (the four shifted copies
// n, n << 7, n << 14, n << 21 do not overlap, bit j of n lands on bit 0 of byte j).  The first version used a 256-entry
// byte -> 8-byte table in shared memory: its bank conflicts doubled the shared-memory wavefronts of the expanders, and
// shared memory is what this kernel runs out of first (the tensor core reads its operands from it as well).
__device__ __forceinline__ uint4 bm_expand16(uint32_t bits) {
    uint4 o;
    o.x = ((bits & 15u) * 0x00204081u) & 0x01010101u;
    o.y = (((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
    o.z = (((bits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
    o.w = (((bits >> 12) & 15u) * 0x00204081u) & 0x01010101u;
    return o;
}

__global__ void __launch_bounds__(BM_THREADS, 1) group_bits_mma_kernel(const BitsMmaParams p) {
    extern __shared__ __align__(128) unsigned char bm_smem[];
    const int K = p.W * 64;                          // K elements = bits
    const int nkb = K / BM_KB;                       // pipeline stages per tile
    unsigned char* Bs = bm_smem;                     // [K/16][64][16]
    unsigned char* As = Bs + (size_t)BM_TN * K;      // [stages][16][128][16]
    double* meta = reinterpret_cast<double*>(As + (size_t)BM_STAGES * BM_TM * BM_KB);        // [4][128][2]  (w, |x|^2)
    double* zn_s = meta + 4 * BM_TM * 2;                                                    // [64]
    __shared__ __align__(8) uint64_t full_bar[BM_STAGES], empty_bar[BM_STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g0 = blockIdx.x * BM_TM;
    const int l0 = blockIdx.y * BM_TN;
    const int64_t r0 = p.row_begin + (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t r1 = min(p.row_end, r0 + p.rows_per_split);
    const int64_t hi = p.pos0 + p.n_local;
    const int ntiles = (int)max((int64_t)0, r1 - r0);

    if (t == 0) {
        for (int s = 0; s < BM_STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        mbar_fence_init();
    }
    if (t < BM_TN) zn_s[t] = (l0 + t < p.L) ? p.zn[l0 + t] : 0.0;
    if (warp == BM_MMA_WARP) {                       // TMEM: 2 accumulator stages x 64 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_addr(&tmem_base_s))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    // resident landmark operand: chunk kc (16 K-elements) of landmark n at Bs[(kc * 64 + n) * 16]
    for (int c = t; c < BM_TN * (K / 16); c += BM_THREADS) {
        const int n = c % BM_TN, kc = c / BM_TN;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (l0 + n < p.L) {
            const uint64_t word = __ldg(p.Z + (int64_t)(l0 + n) * p.W + (kc >> 2));
            val = bm_expand16((uint32_t)(word >> (16 * (kc & 3))) & 0xffffu);
        }
        *reinterpret_cast<uint4*>(Bs + ((size_t)kc * BM_TN + n) * 16) = val;
    }
    bm_fence_async();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
        // ================= EXPANDERS =================
        // Software pipeline over the tiles: the alive-list entry of tile it + 2 and the bit rows of tile it + 1 are in
        // flight while tile it is expanded (the chain idx -> row -> words is two dependent L2 round trips; issued on
        // demand it cost ~5000 cycles per tile and the tensor core sat idle).
        const int r = t;                              // row of the tile
        const int g = g0 + r;
        auto locate = [&](int it, int64_t& row, double& w, bool& ok) {
            const int64_t pos = (r0 + it) * p.S + g;
            ok = it < ntiles && g < p.S && pos >= p.pos0 && pos < hi;
            row = 0;
            w = 0.0;
            if (ok) {
                const int64_t loc = pos - p.pos0;
                row = p.idx ? (int64_t)__ldg(p.idx + loc) : loc;
                w = p.mu ? __ldg(p.mu + loc) : 1.0;
            }
        };
        uint64_t cur[BM_MAXW], nxt[BM_MAXW];
        auto fetch = [&](uint64_t (&dst)[BM_MAXW], int64_t row, bool ok) {
            const ulonglong2* src = reinterpret_cast<const ulonglong2*>(p.X + row * p.W);
#pragma unroll
            for (int wd = 0; wd < BM_MAXW; wd += 2) {
                ulonglong2 v = make_ulonglong2(0ull, 0ull);
                if (ok && wd < p.W) v = __ldg(src + wd / 2);
                dst[wd] = v.x;
                dst[wd + 1] = v.y;
            }
        };
        int64_t row0_, row1_, row2_;
        double w0_, w1_, w2_;
        bool ok0_, ok1_, ok2_;
        locate(0, row0_, w0_, ok0_);
        locate(1, row1_, w1_, ok1_);
        fetch(cur, row0_, ok0_);
        double xn0_ = ok0_ ? __ldg(p.xn + row0_ * p.xn_stride) : 0.0;
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < ntiles; ++it) {
            fetch(nxt, row1_, ok1_);                                     // bit rows of tile it + 1
            const double xn1_ = ok1_ ? __ldg(p.xn + row1_ * p.xn_stride) : 0.0;
            locate(it + 2, row2_, w2_, ok2_);                            // alive-list entry of tile it + 2
            double* m = meta + ((size_t)(it & 3) * BM_TM + r) * 2;
            m[0] = w0_;
            m[1] = xn0_;
#pragma unroll
            for (int kb = 0; kb < BM_MAXW / 4; ++kb) {
                if (kb < nkb) {
                    mbar_wait_sleep(&empty_bar[stage], phase ^ 1u);
                    unsigned char* dst = As + (size_t)stage * BM_TM * BM_KB;
#pragma unroll
                    for (int wd = 0; wd < BM_KB / 64; ++wd) {
                        const uint64_t word = cur[kb * (BM_KB / 64) + wd];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            *reinterpret_cast<uint4*>(dst + ((size_t)(wd * 4 + c) * BM_TM + r) * 16) =
                                bm_expand16((uint32_t)(word >> (16 * c)) & 0xffffu);
                        }
                    }
                    bm_fence_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
                    bm_arrive(&full_bar[stage]);
                    if (++stage == BM_STAGES) { stage = 0; phase ^= 1u; }
                }
            }
#pragma unroll
            for (int wd = 0; wd < BM_MAXW; ++wd) cur[wd] = nxt[wd];
            row0_ = row1_; w0_ = w1_; ok0_ = ok1_; xn0_ = xn1_;
            row1_ = row2_; w1_ = w2_; ok1_ = ok2_;
        }
    } else if (warp == BM_MMA_WARP) {
        // ================= MMA ISSUER =================
        if (lane == 0) {
            // instruction descriptor: D = s32, A = B = unsigned 8-bit, both K-major, N = 64, M = 128
            const uint32_t idesc = (2u << 4) | ((uint32_t)(BM_TN >> 3) << 17) | ((uint32_t)(BM_TM >> 4) << 24);
            const uint32_t a_base = smem_addr(As), b_base = smem_addr(Bs);
            int stage = 0;
            uint32_t phase = 0, tphase[2] = {0, 0};
            for (int it = 0; it < ntiles; ++it) {
                const int acc = it & 1;
                mbar_wait_sleep(&tempty_bar[acc], tphase[acc] ^ 1u);     // epilogue has drained this accumulator stage
                tphase[acc] ^= 1u;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * BM_TN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait_sleep(&full_bar[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < BM_KB / 32; ++j) {
                        const uint64_t adesc = bm_desc(a_base + (uint32_t)stage * BM_TM * BM_KB + (uint32_t)j * 2u * BM_TM * 16u,
                                                       BM_TM * 16u, 128u);
                        const uint64_t bdesc = bm_desc(b_base + (uint32_t)(kb * (BM_KB / 16) + 2 * j) * BM_TN * 16u,
                                                       BM_TN * 16u, 128u);
                        bm_mma_i8(d_tmem, adesc, bdesc, idesc, (kb | j) != 0 ? 1u : 0u);
                    }
                    bm_commit(&empty_bar[stage]);     // the stage is free once these MMAs have read it
                    if (++stage == BM_STAGES) { stage = 0; phase ^= 1u; }
                }
                bm_commit(&tfull_bar[acc]);           // accumulator of this tile complete
            }
        }
        __syncwarp();
    } else {
        // ================= EPILOGUE =================
        // warps 4-7 take accumulator columns 0-31, warps 8-11 columns 32-63; a warp may access TMEM lanes
        // [32 q, 32 q + 32) with q = warp % 4
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int r = 32 * q + lane;                  // row of the tile = TMEM lane
        const int g = g0 + r;
        constexpr int HN = BM_TN / 2;
        double acc[HN];
#pragma unroll
        for (int l = 0; l < HN; ++l) acc[l] = 0.0;
        double tw = 0.0;
        uint32_t tphase[2] = {0, 0};
        for (int it = 0; it < ntiles; ++it) {
            const int a = it & 1;
            mbar_wait_sleep(&tfull_bar[a], tphase[a]);
            tphase[a] ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            bm_tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)a * BM_TN + (uint32_t)half * HN, v);
            // the accumulator stage is in registers: hand it back to the MMA warp before the FP64 work
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bm_arrive(&tempty_bar[a]);
            const double* m = meta + ((size_t)(it & 3) * BM_TM + r) * 2;
            const double w = m[0], exn = m[1];
            const int64_t pos = (r0 + it) * p.S + g;
            if (w != 0.0) {
                if (half == 0 && blockIdx.y == 0 && pos < p.ES) tw += w;
#pragma unroll
                for (int l = 0; l < HN; ++l)
                    acc[l] = fma(tanimoto_bits_value(__hiloint2double(0x43300000, (int)v[l]) - 4503599627370496.0, exn, zn_s[half * HN + l] + 1e-6), w, acc[l]);
            }
        }
        if (g < p.S) {
            double* out = p.out + ((int64_t)blockIdx.z * p.S + g) * p.L + l0 + half * HN;
#pragma unroll
            for (int l = 0; l < HN; ++l)
                if (l0 + half * HN + l < p.L) out[l] = acc[l] * p.scale;
            if (half == 0 && blockIdx.y == 0) p.totw_out[(int64_t)blockIdx.z * p.S + g] = tw;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == BM_MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
    }
}

// =====================================================================================================
// Version 2: the LANDMARK tile lives in TMEM as the A operand of the MMA.
//
// Version 1 (above) runs out of shared memory first: per 128 x 64 tile the expanders store 128 KB and the tensor core
// itself reads 192 KB of operands (A 4 KB + B 2 KB per MMA).  Here A = 128 landmarks x d bytes is expanded ONCE per CTA
// into tensor memory (tcgen05.st: lane = landmark, one 32-bit column = 4 K-elements = one nibble of the bit row; 256
// columns at d = 1024) and only the streamed candidate tile (B: 64 candidates x 256 K-elements per stage, 16 KB) passes
// through shared memory: 64 KB stored + 64 KB read by the tensor core per 128 x 64 tile, and each candidate row is
// expanded by L / 128 CTAs instead of L / 64.  The epilogue thread is now a LANDMARK (TMEM lane) and walks the tile's
// candidates (accumulator columns): 16 epilogue warps (4 column groups per lane quarter), 16 FP64 accumulators per
// thread, stores to At coalesced along the landmarks.  TMEM: 256 columns A + 2 x 64 columns of int32 accumulators (512
// allocated).  Warps: 0-7 expanders (4 threads per candidate row), 8-23 epilogue, 24 MMA (whole warp runs the loop, an
// elected lane issues), 25-26 loaders (lane = candidate row, cp.async staging ring).  Two 64 KB operand slots, one
// fence.proxy.async per tile and thread.  What each design step bought, with the ncu readings: profiles/r02_bits_tcgen05.txt
// =====================================================================================================
constexpr int B2_TN = 64;         // candidates (groups) per tile = accumulator columns
constexpr int B2_TM = 128;        // landmarks per CTA = TMEM lanes
constexpr int B2_SLOTS = 2;                     // operand slots: one tile's K-blocks each (64 rows x d bytes, <= 64 KB)
constexpr int B2_SLOT_BYTES = B2_TN * BM_MAXW * 64;
constexpr int B2_META = 16;       // ring of per-tile candidate metadata: the epilogue lags the expanders by at most 7 tiles (d = 256)
// Epilogue warps: B2_EG column groups per TMEM lane quarter.  One warp issues an FP64 instruction only every ~13 cycles
// (measured: 2 epilogue warps per scheduler left the FP64 pipe at 30 % with every warp stalled on it), so the epilogue
// wants thread-level parallelism: 4 warps per scheduler, 16 candidates (accumulator columns) of the tile each.
constexpr int B2_EG = 4;
constexpr int B2_HN = B2_TN / B2_EG;
constexpr int B2_EPI_WARPS = 4 * B2_EG;
#ifndef SOBER_B2_XW
#define SOBER_B2_XW 8
#endif
constexpr int B2_XW = SOBER_B2_XW;              // expander warps (a multiple of 4, so the epilogue warps keep warp % 4 = lane quarter)
constexpr int B2_XP = B2_XW * 32 / B2_TN;       // expander threads per candidate row
constexpr int B2_WPT = 4 / B2_XP;               // 64-bit words per expander thread and 256-bit K-block
constexpr int B2_MMA_WARP = B2_XW + B2_EPI_WARPS;
constexpr int B2_DEPTH = 3;                     // tiles of global-load lookahead in the loaders
constexpr int B2_RING = 8;                      // staging slots for raw words / weights / popcounts (> B2_DEPTH + 1, power of 2)
constexpr int B2_IRING = 8;                     // staging slots for alive-list entries (>= 2 B2_DEPTH + 1, power of 2)
constexpr int B2_THREADS = (B2_MMA_WARP + 1 + B2_TN / 32) * 32;   // + the loader warps (one lane per candidate row)

__device__ __forceinline__ void bm_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool bm_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int BYTES>
__device__ __forceinline__ void bm_cp_async(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr(dst_smem)), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void bm_tmem_ldn(uint32_t taddr, uint32_t (&v)[16]) { bm_tmem_ld16(taddr, v); }
__device__ __forceinline__ void bm_tmem_ldn(uint32_t taddr, uint32_t (&v)[32]) { bm_tmem_ld32(taddr, v); }

__device__ __forceinline__ void bm_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
                 : "memory");
}
__device__ __forceinline__ void bm_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(B2_THREADS, 1) group_bits_mma2_kernel(const BitsMmaParams p) {
    extern __shared__ __align__(128) unsigned char bm_smem[];
    const int K = p.W * 64;
    const int nkb = K / BM_KB;
    unsigned char* Bs = bm_smem;                                                     // [slots][4 K-blocks][16][64][16]
    double* meta = reinterpret_cast<double*>(Bs + (size_t)B2_SLOTS * B2_SLOT_BYTES);  // [B2_META][64][2]  (w, |x|^2)
    uint64_t* raw = reinterpret_cast<uint64_t*>(meta + (size_t)B2_META * B2_TN * 2); // [B2_RING][BM_MAXW][64] staged words
    double* wraw = reinterpret_cast<double*>(raw + (size_t)B2_RING * BM_MAXW * B2_TN);   // [B2_RING][64]
    double* xraw = wraw + B2_RING * B2_TN;                                           // [B2_RING][64]
    int* irow = reinterpret_cast<int*>(xraw + B2_RING * B2_TN);                      // [B2_IRING][64]
    __shared__ __align__(8) uint64_t full_bar[B2_SLOTS], empty_bar[B2_SLOTS], tfull_bar[2], tempty_bar[2];
    __shared__ __align__(8) uint64_t raw_full[B2_RING], raw_empty[B2_RING];
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, lane = t & 31, warp = __shfl_sync(0xffffffffu, t >> 5, 0);   // provably warp-uniform
    const int g0 = blockIdx.x * B2_TN;
    const int l0 = blockIdx.y * B2_TM;
    const int64_t r0 = p.row_begin + (int64_t)blockIdx.z * p.rows_per_split;
    const int64_t r1 = min(p.row_end, r0 + p.rows_per_split);
    const int64_t hi = p.pos0 + p.n_local;
    const int ntiles = (int)max((int64_t)0, r1 - r0);
    auto alive_at = [&](int it, int g, int64_t& loc) {
        const int64_t pos = (r0 + it) * p.S + g;
        loc = pos - p.pos0;
        return it >= 0 && it < ntiles && g < p.S && pos >= p.pos0 && pos < hi;
    };

    if (t == 0) {
        for (int s = 0; s < B2_SLOTS; ++s) { mbar_init(&full_bar[s], B2_XW * 32); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], B2_EPI_WARPS); }
        for (int s = 0; s < B2_RING; ++s) { mbar_init(&raw_full[s], B2_TN); mbar_init(&raw_empty[s], B2_XW); }
        mbar_fence_init();
    }
    if (warp == B2_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tmem_base_s))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    const uint32_t tmem_a = tmem_base;                 // columns [0, K / 4): the landmark operand
    const uint32_t tmem_d = tmem_base + 256;           // columns [256, 384): two accumulator stages
    // resident A operand: 4 epilogue warps (TMEM lane quarters 0-3), thread = landmark, 32 columns (= 32 nibbles) per store
    if (warp >= B2_XW && warp < B2_XW + 4) {
        const int q = warp & 3;
        const int l = l0 + 32 * q + lane;
        for (int c0 = 0; c0 < K / 4; c0 += 32) {       // 32 columns = 128 K-elements = 2 words
            uint32_t v[32];
            uint64_t w0 = 0ull, w1 = 0ull;
            if (l < p.L) {
                w0 = __ldg(p.Z + (int64_t)l * p.W + c0 / 16);
                w1 = __ldg(p.Z + (int64_t)l * p.W + c0 / 16 + 1);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                v[j] = (((uint32_t)(w0 >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;
                v[16 + j] = (((uint32_t)(w1 >> (4 * j)) & 15u) * 0x00204081u) & 0x01010101u;
            }
            bm_tmem_st32(tmem_a + ((uint32_t)(32 * q) << 16) + (uint32_t)c0, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (warp > B2_MMA_WARP) {
        // ================= LOADERS: lane = candidate row of the tile =================
        // Global latency stays out of everybody else's loop: cp.async into staging slots, B2_DEPTH tiles ahead for the
        // fingerprint words / weight / popcount and 2 B2_DEPTH tiles ahead for the alive-list entry they depend on.  The
        // loaders never fence (the expanders' fence.proxy.async compiles to a MEMBAR that would wait for every copy in
        // flight), and hand a slot over with an mbarrier once cp.async.wait_group says it has landed.
        constexpr int D = B2_DEPTH;
        const int r = t - (B2_MMA_WARP + 1) * 32;
        const int g = g0 + r;
        double tw = 0.0;
        for (int it = -2 * D; it < ntiles; ++it) {
            asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
            int64_t loc;
            if (p.idx && alive_at(it + 2 * D, g, loc)) bm_cp_async<4>(&irow[((it + 2 * D) & (B2_IRING - 1)) * B2_TN + r], p.idx + loc);
            const int tb = it + D;
            if (tb >= 0 && tb < ntiles) {
                const int slot = tb & (B2_RING - 1);
                mbar_wait_backoff(&raw_empty[slot], ((uint32_t)(tb / B2_RING) & 1u) ^ 1u, 256);
                if (alive_at(tb, g, loc)) {
                    const int64_t row = p.idx ? (int64_t)irow[(tb & (B2_IRING - 1)) * B2_TN + r] : loc;
                    const uint64_t* src = p.X + row * p.W;
                    for (int w = 0; w < p.W; ++w) bm_cp_async<8>(&raw[((size_t)slot * BM_MAXW + w) * B2_TN + r], src + w);
                    if (p.mu) bm_cp_async<8>(&wraw[slot * B2_TN + r], p.mu + loc);
                    bm_cp_async<8>(&xraw[slot * B2_TN + r], p.xn + row * p.xn_stride);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (it < 0) continue;
            const bool ok = alive_at(it, g, loc);
            const int slot = it & (B2_RING - 1);
            const double w = ok ? (p.mu ? wraw[slot * B2_TN + r] : 1.0) : 0.0;
            double* m = meta + ((size_t)(it & (B2_META - 1)) * B2_TN + r) * 2;
            m[0] = w;
            m[1] = ok ? xraw[slot * B2_TN + r] : 0.0;
            if ((r0 + it) * p.S + g < p.ES) tw += w;
            bm_arrive(&raw_full[slot]);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (blockIdx.y == 0 && g < p.S) p.totw_out[(int64_t)blockIdx.z * p.S + g] = tw;
    } else if (warp < B2_XW) {
        // ================= EXPANDERS: thread (row r, part h) expands word h of each of the tile's 4-word K-blocks =================
        // ONE fence.proxy.async per tile and thread (it costs a MEMBAR.ALL.CTA): the tile's K-blocks go to one of two
        // 64 KB operand slots, so the MMA warp works on a slot while the next is being filled.
        const int r = t & (B2_TN - 1), h = t / B2_TN;
        const int g = g0 + r;
        for (int it = 0; it < ntiles; ++it) {
            const int slot = it & (B2_RING - 1);
            int64_t loc;
            const bool ok = alive_at(it, g, loc);
            mbar_wait_backoff(&raw_full[slot], (uint32_t)(it / B2_RING) & 1u, 128);
            uint64_t word[BM_MAXW / 4][B2_WPT];
#pragma unroll
            for (int kb = 0; kb < BM_MAXW / 4; ++kb)
#pragma unroll
                for (int wd = 0; wd < B2_WPT; ++wd)
                    word[kb][wd] = (ok && kb < nkb) ? raw[((size_t)slot * BM_MAXW + kb * 4 + h * B2_WPT + wd) * B2_TN + r] : 0ull;
            __syncwarp();
            if (lane == 0) bm_arrive(&raw_empty[slot]);
            const int bs = it & (B2_SLOTS - 1);
            mbar_wait_backoff(&empty_bar[bs], ((uint32_t)(it / B2_SLOTS) & 1u) ^ 1u, 128);
            unsigned char* dst = Bs + (size_t)bs * B2_SLOT_BYTES;
#pragma unroll
            for (int kb = 0; kb < BM_MAXW / 4; ++kb) {
                if (kb < nkb) {
#pragma unroll
                    for (int wd = 0; wd < B2_WPT; ++wd) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<uint4*>(dst + (size_t)kb * B2_TN * BM_KB +
                                                      ((size_t)((B2_WPT * h + wd) * 4 + c) * B2_TN + r) * 16) =
                                bm_expand16((uint32_t)(word[kb][wd] >> (16 * c)) & 0xffffu);
                    }
                }
            }
            bm_fence_async();
            bm_arrive(&full_bar[bs]);
        }
    } else if (warp == B2_MMA_WARP) {
        // ================= MMA ISSUER =================
        // The whole warp runs the loop and one ELECTED lane issues: every operand is then warp-uniform for the compiler
        // (warp index and TMEM base come through a shuffle from lane 0), so each tcgen05.mma is one UIADD3 + UTCIMMA.
        // Inside an `if (lane == 0)` the same code compiled to a 12-instruction waterfall loop per MMA (R2UR, ELECT,
        // BRA.U.ANY) whose ~130-cycle latency -- not the 32-cycle MMA -- set the tile time (tensor pipe 26 % active).
        const uint32_t idesc = (2u << 4) | ((uint32_t)(B2_TN >> 3) << 17) | ((uint32_t)(B2_TM >> 4) << 24);
        const uint64_t desc0 = bm_desc(smem_addr(Bs), B2_TN * 16u, 128u);
        for (int it = 0; it < ntiles; ++it) {
            const int acc = it & 1, bs = it & (B2_SLOTS - 1);
            mbar_wait_backoff(&tempty_bar[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u, 32);
            mbar_wait_backoff(&full_bar[bs], (uint32_t)(it / B2_SLOTS) & 1u, 32);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_d + (uint32_t)acc * B2_TN;
            const uint64_t desc_t = desc0 + (uint64_t)((uint32_t)bs * (B2_SLOT_BYTES >> 4));
            if (bm_elect_one()) {
                for (int kb = 0; kb < nkb; ++kb) {
#pragma unroll
                    for (int j = 0; j < BM_KB / 32; ++j)      // A: 8 columns (32 K-elements), B: two 16-byte K-chunks per MMA
                        bm_mma_i8_ts(d_tmem, tmem_a + (uint32_t)(kb * (BM_KB / 4) + 8 * j),
                                     desc_t + (uint64_t)((kb * B2_TN * BM_KB + j * 2 * B2_TN * 16) >> 4), idesc, (kb | j) != 0 ? 1u : 0u);
                }
                bm_commit(&empty_bar[bs]);
                bm_commit(&tfull_bar[acc]);
            }
            __syncwarp();
        }
    } else {
        // ================= EPILOGUE: thread = landmark (TMEM lane), B2_HN candidates (columns) of the tile =================
        const int q = warp & 3, cg = (warp - B2_XW) >> 2;
        const int l = l0 + 32 * q + lane;
        const double zne = (l < p.L ? __ldg(p.zn + l) : 0.0) + 1e-6;
        double acc[B2_HN];
#pragma unroll
        for (int j = 0; j < B2_HN; ++j) acc[j] = 0.0;
        for (int it = 0; it < ntiles; ++it) {
            const int a = it & 1;
            mbar_wait_backoff(&tfull_bar[a], (uint32_t)(it >> 1) & 1u, 256);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[B2_HN];
            bm_tmem_ldn(tmem_d + ((uint32_t)(32 * q) << 16) + (uint32_t)a * B2_TN + (uint32_t)cg * B2_HN, v);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bm_arrive(&tempty_bar[a]);
            const double2* m = reinterpret_cast<const double2*>(meta + ((size_t)(it & (B2_META - 1)) * B2_TN + cg * B2_HN) * 2);
#pragma unroll
            for (int j = 0; j < B2_HN; ++j) {
                const double2 we = m[j];                  // (w, |x|^2) of candidate j: broadcast load
                const double dot = __hiloint2double(0x43300000, (int)v[j]) - 4503599627370496.0;
                acc[j] = fma(tanimoto_bits_value(dot, we.y, zne), we.x, acc[j]);
            }
        }
        if (l < p.L) {
#pragma unroll
            for (int j = 0; j < B2_HN; ++j) {
                const int g = g0 + cg * B2_HN + j;
                if (g < p.S) p.out[((int64_t)blockIdx.z * p.S + g) * p.L + l] = acc[j] * p.scale;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == B2_MMA_WARP) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

size_t bits_mma2_smem() {
    return (size_t)B2_SLOTS * B2_SLOT_BYTES + (size_t)B2_META * B2_TN * 2 * 8 +
           (size_t)B2_RING * BM_MAXW * B2_TN * 8 + 2 * (size_t)B2_RING * B2_TN * 8 + (size_t)B2_IRING * B2_TN * 4 + 128;
}

size_t bits_mma_smem(int W) {
    const size_t K = (size_t)W * 64;
    return (size_t)BM_TN * K + (size_t)BM_STAGES * BM_TM * BM_KB + 256 * 8 + 4 * BM_TM * 2 * 8 + BM_TN * 8 + 128;
}

bool bits_mma_supported(int W) { return W >= 4 && W <= BM_MAXW && W % 4 == 0; }

int launch_bits_mma2(const BitsMmaParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = bits_mma2_smem();
    static int configured[64] = {};
    int dev = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        SOBER_CUDA_CHECK(cudaFuncSetAttribute(group_bits_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = 1;
    }
    group_bits_mma2_kernel<<<grid, B2_THREADS, smem, st>>>(p);
    SOBER_LAUNCH_CHECK("group_bits_mma2");
    return SOBER_OK;
}

int launch_bits_mma(const BitsMmaParams& p, dim3 grid, cudaStream_t st) {
    const size_t smem = bits_mma_smem(p.W);
    static int configured[64] = {};
    int dev = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || configured[dev] < (int)smem) {
        SOBER_CUDA_CHECK(cudaFuncSetAttribute(group_bits_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) configured[dev] = (int)smem;
    }
    group_bits_mma_kernel<<<grid, BM_THREADS, smem, st>>>(p);
    SOBER_LAUNCH_CHECK("group_bits_mma");
    return SOBER_OK;
}

}  // namespace sober
