// Shared device/host helpers for the sober_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/sober_b200.h"

namespace sober {

void set_cuda_error(cudaError_t e, const char* where);

#define SOBER_CUDA_CHECK(expr)                                  \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) {                                \
            ::sober::set_cuda_error(_e, #expr);                 \
            return SOBER_ERR_CUDA;                              \
        }                                                       \
    } while (0)

#define SOBER_LAUNCH_CHECK(name)                                \
    do {                                                        \
        cudaError_t _e = cudaGetLastError();                    \
        if (_e != cudaSuccess) {                                \
            ::sober::set_cuda_error(_e, name);                  \
            return SOBER_ERR_CUDA;                              \
        }                                                       \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

int sm_count();
int stream_sm_count(void* stream);   // csrc/partition.cu: the SM-partitioned stream sees fewer SMs
int partition_sm_count();             // 0 = no partition on this device

// ---------------------------------------------------------------------------------------------------
// FP64 math tuned for the FP64 pipe (B200: 64 DFMA/clk/SM, no FP64 SFU).  The library exp()/sqrt()/division
// cost ~3x more issue slots than FP64 slots (special-case branches, integer fix-ups, register moves); K1 is
// bound by exactly those, so the three primitives are restated with the minimum number of FP64 instructions:
//
//   exp_neg(s) = exp(-s), s >= 0      6 FP64 + ~6 integer/LDS    max rel err 8.1e-13 + 1.1e-16 * s (checked against
//                                      50-digit arithmetic, tools/check_fastmath.py)
//   sqrt_pos(a), a >= 1e-30            3 FP64 + 1 MUFU + 1 int    rel err <= 8.5e-14
//   div_pos(a, b), b > 0               7 FP64 + 1 MUFU            <= 1 ulp
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Table size 2^B and polynomial degree of exp_neg.  |r| <= ln2 / 2^(B+1), truncation error r^(deg+1) / (deg+1)!:
//   B = 6,  deg 5: 3.5e-17  (round 1)        B = 8,  deg 3: 1.4e-13
//   B = 11, deg 2: 8.1e-13  (default: 3 DFMA fewer per kernel value; the tolerance on kernel values is 1e-10)
// checked against 50-digit arithmetic by tools/check_fastmath.py.
#ifndef SOBER_EXP_TAB_BITS
#define SOBER_EXP_TAB_BITS 11
#endif
#ifndef SOBER_EXP_DEG
#define SOBER_EXP_DEG 2
#endif
#ifndef SOBER_SQRT_INTHALF
#define SOBER_SQRT_INTHALF 1
#endif
constexpr int EXP_TAB_BITS = SOBER_EXP_TAB_BITS;
constexpr int EXP_TAB_SIZE = 1 << EXP_TAB_BITS;
static_assert(EXP_TAB_BITS >= 4 && EXP_TAB_BITS <= 11, "exp table: 16 .. 2048 entries");
static __device__ const double EXP2_TAB[2048] = {  // 2^(j/2048), correctly rounded (tools/gen_exp2_table.py)
#include "exp2_tab.inc"
};

// 64-bit constants of the routines below.  Read as constant-bank operands (DFMA ..., c[0x3][..]) they cost no
// instruction; as literals the compiler rebuilds each one with two UMOV/IMAD.MOV per use (measured in the K1 SASS:
// 6 of 48 instructions per kernel value).
static __constant__ double MATH_C[6] = {
    -(double)EXP_TAB_SIZE / 0.69314718055994530942,   // [0] -2^B / ln2
    0.69314718055994530942 / (double)EXP_TAB_SIZE,    // [1] ln2 / 2^B (a power-of-two scaling of ln2 rounded to double)
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,   // [2..4] exp polynomial
    1.0 / 3.0};               // [5] Matern-5/2

__device__ __forceinline__ void load_exp_table(double* tab_smem, int tid, int nthreads) {
    for (int j = tid; j < EXP_TAB_SIZE; j += nthreads) tab_smem[j] = EXP2_TAB[j << (11 - EXP_TAB_BITS)];
}

__device__ __forceinline__ double lds_f64(uint32_t saddr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
    return v;
}

// `tab` is the SHARED-SPACE address of the 64-entry table (smem_addr(tab_smem)): a generic pointer would cost a
// generic->shared conversion (S2UR SR_CgaCtaId + ULEA) at every use
__device__ __forceinline__ double exp_neg(double s_in, uint32_t tab) {
    // exp(-s) = 2^m * 2^(j/T) * exp(r),  -s = (T m + j) ln2/T + r,  |r| <= ln2/(2T),  T = EXP_TAB_SIZE
    const double MAGIC = 6755399441055744.0;            // 1.5 * 2^52: adds round-to-nearest-integer
    const double NEG_L2ET = MATH_C[0];                  // -T / ln2
    const double LN2_T = MATH_C[1];                     // ln2/T rounded to double
    // clamp s at ~700 through the integer pipe (s >= 0: high words order like ints): below 1e-304 the exponent
    // trick at the end would wrap
    const double s = __hiloint2double(min(__double2hiint(s_in), 0x4085E000), __double2loint(s_in));
    const double kd = fma(s, NEG_L2ET, MAGIC);
    const int k = __double2loint(kd);
    const double kf = kd - MAGIC;
    // one FMA: the product kf * LN2_T is exact inside the FMA, so r carries only the rounding of the constant
    // (|s| * 1.1e-16 absolute) -- no hi/lo split needed
    const double r = fma(kf, -LN2_T, -s);
#if SOBER_EXP_DEG >= 5
    double q = fma(MATH_C[2], r, MATH_C[3]);
    q = fma(q, r, MATH_C[4]);
    q = fma(q, r, 0.5);
#elif SOBER_EXP_DEG == 4
    double q = fma(MATH_C[3], r, MATH_C[4]);
    q = fma(q, r, 0.5);
#elif SOBER_EXP_DEG == 3
    double q = fma(MATH_C[4], r, 0.5);
#else
    double q = 0.5;
#endif
    q = fma(q, r, 1.0);
    const double p = fma(q, r, 1.0);
    const double v = lds_f64(tab + ((k & (EXP_TAB_SIZE - 1)) << 3)) * p;   // in [1, 2)
    // scale by 2^m on the exponent field (ALU pipe): hi += (k >> B) << 20 == (k & ~(T-1)) << (20 - B)
    const int hi = __double2hiint(v) + ((k & ~(EXP_TAB_SIZE - 1)) << (20 - EXP_TAB_BITS));
    return __hiloint2double(hi, __double2loint(v));
}

__device__ __forceinline__ double sqrt_pos(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));   // MUFU.RSQ64H, rel err 2^-22
    // one Goldschmidt step: relative error 1.5 * (2^-22)^2 = 8.5e-14 (the tolerance on kernel values is 1e-10)
#if SOBER_SQRT_INTHALF
    // y / 2 on the exponent field (integer pipe; y = rsqrt(a) <= 1e15 for a >= 1e-30: no underflow) -- one FP64 less
    const double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    const double g = a * y;
#else
    const double g = a * y, h = 0.5 * y;
#endif
    const double e = fma(-g, h, 0.5);
    return fma(g, e, g);
}

__device__ __forceinline__ double div_pos(double a, double b) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));     // MUFU.RCP64H
    double e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    e = fma(-b, y, 1.0);
    y = fma(y, e, y);
    const double q = a * y;
    return fma(fma(-b, q, a), y, q);
}

__device__ __forceinline__ double clamp_min_pos(double a, double floor_) {
    // max(a, floor_) for floor_ > 0 through the integer pipe: for non-negative doubles the high words order
    // like signed ints, and a (slightly) negative a -- cancellation noise -- has a negative high word.
    const int hi = max(__double2hiint(a), __double2hiint(floor_));
    return __hiloint2double(hi, __double2loint(a));
}

// ---------------------------------------------------------------------------------------------------
// Kernel nonlinearities.  `dot` is sum_k x_k * zt_k where zt already carries the factor -2 for the
// stationary families, so the squared distance is xn + zn + dot.  The host folds the family's constant
// into the lengthscale (coordinates are multiplied by 1/sqrt2, 1, sqrt3, sqrt5 for RBF / Matern-1/2, -3/2,
// -5/2), so here   RBF = exp(-d2),  M12 = exp(-r),  M32 = (1+r) exp(-r),  M52 = (1 + r + d2/3) exp(-r)
// with r = sqrt(d2).  The output scale is applied once per output entry by the caller (linearity).
// ---------------------------------------------------------------------------------------------------
template <int FAM>
__device__ __forceinline__ double stationary_value(double d2, uint32_t tab) {
    if (FAM == SOBER_RBF) {
        return exp_neg(clamp_min_pos(d2, 0.0), tab);
    }
    const double a = clamp_min_pos(d2, 1e-30);
    const double r = sqrt_pos(a);
    const double ex = exp_neg(r, tab);
    if (FAM == SOBER_MATERN12) return ex;
    if (FAM == SOBER_MATERN32) return (1.0 + r) * ex;
    return fma(a, MATH_C[5], 1.0 + r) * ex;
}

__device__ __forceinline__ double tanimoto_value(double dot, double xn, double zn) {
    const double eps = 1e-6;
    const double v = div_pos(dot + eps, (eps + xn) + (zn - dot));
    return fmax(v, 0.0);
}

// Tanimoto on bit-packed rows: dot, |x|^2, |z|^2 are exact small integers with dot <= min(|x|^2, |z|^2), so
// m = |x|^2 + |z|^2 - dot is an exact integer >= 0 and the ratio (dot + eps) / (m + eps) needs no clamp.  zne = |z|^2 + eps
// is formed once per landmark.  Reciprocal: MUFU seed + ONE Newton step, 3.3e-13 measured against the oracle's division,
// far below the 1e-10 the kernel matrix has to match to; the second step cost 2 of the 10 FP64 instructions per
// pair that bound the tcgen05 epilogue.  Shared by the popcount kernel and the tcgen05 kernels: their Grams are bitwise equal.
__device__ __forceinline__ double tanimoto_bits_value(double dot, double xn, double zne) {
    const double den = (xn + zne) - dot;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
    y = fma(y, fma(-den, y, 1.0), y);
    return (dot + 1e-6) * y;
}

template <int FAM>
__device__ __forceinline__ double kernel_value(double dot, double xn, double zn, uint32_t tab) {
    if (FAM == SOBER_TANIMOTO) return tanimoto_value(dot, xn, zn);
    return stationary_value<FAM>((xn + zn) + dot, tab);
}

// ---------------------------------------------------------------------------------------------------
// TMA 1-D bulk copy + mbarrier (PTX; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
// same with a suspend-time hint (ns): a warp that waits for a whole pipeline stage sleeps in hardware instead of
// re-issuing the probe (in the tcgen05 Tanimoto kernel 15 % of all issued instructions were such probes)
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity), "r"(20000u)
        : "memory");
}
// the same with a back-off: a failed try_wait puts the warp to sleep instead of polling the barrier's shared-memory word
// (16 epilogue warps polling one barrier were a quarter of the shared-memory wavefronts of the tcgen05 Tanimoto kernel)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns = 64) {
    uint32_t done;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

}  // namespace sober
