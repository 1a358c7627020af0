// SM partition for overlapping the first K1 pass with the latency-bound tail of the Nystrom range finder.
//
// The tail (eigh + refinement sweeps) is ~2.5 ms of one-CTA kernels; the first K1 pass is ~2.3 ms of a grid that
// fills every SM for ~0.4 ms per CTA.  Run side by side on two ordinary streams they do not overlap: a tail kernel
// only gets an SM when a K1 CTA retires, and those retire in waves.  A CUDA green context confines the K1 stream to
// all SMs but a few; the tail keeps its stream of the primary context and always finds the reserved SMs free.
//
// The driver entry points are resolved through cudaGetDriverEntryPoint, so the library does not link libcuda (it must
// load on machines without a driver: the CPU-side ABI tests).
#include <cuda.h>

#include "common.cuh"

namespace {

template <typename F>
bool driver_fn(const char* name, F* out) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess ||
        !fn) {
        (void)cudaGetLastError();
        return false;
    }
    *out = reinterpret_cast<F>(fn);
    return true;
}

struct Partition {
    CUgreenCtx ctx = nullptr;
    CUstream stream = nullptr;
    int sms = 0;
    bool tried = false;
};
Partition g_part[64];

}  // namespace

namespace sober {
// SMs a launch on ``stream`` can use: the partition's share when it is this device's SM-partitioned stream, else all.
// K1 sizes its grid in whole waves of resident CTAs, and the first (largest) pass runs on the partition.
int stream_sm_count(void* stream) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return sm_count();
    const Partition& part = g_part[dev];
    return (stream && part.stream && (void*)part.stream == stream && part.sms > 0) ? part.sms : sm_count();
}
int partition_sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    return g_part[dev].stream ? g_part[dev].sms : 0;
}
}  // namespace sober

// Stream confined to (all SMs - reserve_sms) of the current device, created on first use and cached per device.
// *stream = NULL (and SOBER_OK) when the driver cannot partition the device: the caller then simply does not overlap.
extern "C" int sober_partition_stream(int32_t reserve_sms, void** stream, int32_t* sm_count_out) {
    if (!stream || reserve_sms <= 0) return SOBER_ERR_ARG;
    *stream = nullptr;
    if (sm_count_out) *sm_count_out = 0;
    int dev = 0;
    SOBER_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return SOBER_OK;
    Partition& part = g_part[dev];
    if (!part.tried) {
        part.tried = true;
        SOBER_CUDA_CHECK(cudaFree(0));   // primary context up
        decltype(&cuDeviceGet) p_device_get;
        decltype(&cuDeviceGetDevResource) p_get_res;
        decltype(&cuDevSmResourceSplitByCount) p_split;
        decltype(&cuDevResourceGenerateDesc) p_desc;
        decltype(&cuGreenCtxCreate) p_create;
        decltype(&cuGreenCtxStreamCreate) p_stream;
        if (driver_fn("cuDeviceGet", &p_device_get) && driver_fn("cuDeviceGetDevResource", &p_get_res) &&
            driver_fn("cuDevSmResourceSplitByCount", &p_split) && driver_fn("cuDevResourceGenerateDesc", &p_desc) &&
            driver_fn("cuGreenCtxCreate", &p_create) && driver_fn("cuGreenCtxStreamCreate", &p_stream)) {
            CUdevice cu_dev;
            CUdevResource all, big, rest;
            CUdevResourceDesc desc;
            unsigned int groups = 1;
            if (p_device_get(&cu_dev, dev) == CUDA_SUCCESS &&
                p_get_res(cu_dev, &all, CU_DEV_RESOURCE_TYPE_SM) == CUDA_SUCCESS &&
                (int)all.sm.smCount > 2 * reserve_sms &&
                p_split(&big, &groups, &all, &rest, 0, all.sm.smCount - (unsigned)reserve_sms) == CUDA_SUCCESS &&
                groups == 1 && big.sm.smCount < all.sm.smCount && p_desc(&desc, &big, 1) == CUDA_SUCCESS &&
                p_create(&part.ctx, desc, cu_dev, CU_GREEN_CTX_DEFAULT_STREAM) == CUDA_SUCCESS) {
                if (p_stream(&part.stream, part.ctx, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS)
                    part.sms = (int)big.sm.smCount;
                else
                    part.stream = nullptr;
            }
        }
        (void)cudaGetLastError();
    }
    *stream = part.stream;
    if (sm_count_out) *sm_count_out = part.sms;
    return SOBER_OK;
}
