// Library-wide helpers: error text, device queries, ABI version.
#include "common.cuh"

#include <stdio.h>
#include <string.h>

namespace sober {

static thread_local char g_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    (void)cudaGetLastError();  // clear the sticky-less error state
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace sober

extern "C" int sober_abi_version(void) { return SOBER_B200_ABI_VERSION; }
extern "C" const char* sober_last_cuda_error(void) { return sober::g_err; }
extern "C" int sober_sm_count(int* out) {
    if (!out) return SOBER_ERR_ARG;
    int dev = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    SOBER_CUDA_CHECK(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return SOBER_OK;
}
