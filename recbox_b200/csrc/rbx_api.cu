// Error plumbing and device queries behind the C ABI (include/recbox_b200.h).
#include <stdarg.h>
#include <string.h>
#include "rbx_common.cuh"

static thread_local char g_err[512] = "";

int rbx_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int rbx_sm_count() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

extern "C" {

int rbx_version(void) { return 10000 * 0 + 100 * 1 + 0; }

const char* rbx_last_error(void) { return g_err; }

int rbx_device_sm_count(void) {
    int dev = 0, n = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "rbx_device_sm_count: %s", cudaGetErrorString(e));
    return n;
}

// L2 set-aside for persisting (evict_last) lines: without it the evict_last hints of the fused kernels
// have nothing to live in.  bytes is clamped to the device maximum; returns the size in effect (<0 = error).
long long rbx_l2_set_persisting_bytes(long long bytes) {
    int dev = 0, max_bytes = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&max_bytes, cudaDevAttrMaxPersistingL2CacheSize, dev);
    if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "rbx_l2_set_persisting_bytes: %s", cudaGetErrorString(e));
    if (bytes < 0) bytes = 0;
    if (bytes > max_bytes) bytes = max_bytes;
    e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)bytes);
    if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "rbx_l2_set_persisting_bytes: %s", cudaGetErrorString(e));
    return bytes;
}

}  // extern "C"
