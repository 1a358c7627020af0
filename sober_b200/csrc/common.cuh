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

// ---------------------------------------------------------------------------------------------------
// Kernel nonlinearities.  `dot` is sum_k x_k * zt_k where zt already carries the factor -2 for the
// stationary families, so d2 = xn + zn + dot.  The output scale is applied once per output entry by
// the caller (linearity), not per pair.
// ---------------------------------------------------------------------------------------------------
template <int FAM>
__device__ __forceinline__ double kernel_value(double dot, double xn, double zn) {
    if (FAM == SOBER_TANIMOTO) {
        const double eps = 1e-6;
        double v = (dot + eps) / (eps + xn + zn - dot);
        return fmax(v, 0.0);
    }
    double d2 = fmax(xn + zn + dot, 0.0);
    if (FAM == SOBER_RBF) {
        return exp(-0.5 * d2);
    }
    double r = sqrt(fmax(d2, 1e-30));
    if (FAM == SOBER_MATERN12) {
        return exp(-r);
    }
    if (FAM == SOBER_MATERN32) {
        double s = 1.7320508075688772 * r;
        return (1.0 + s) * exp(-s);
    }
    // Matern-5/2
    double s = 2.23606797749979 * r;
    return (1.0 + s + (5.0 / 3.0) * (r * r)) * exp(-s);
}

}  // namespace sober
