// Shared device helpers and host-side error plumbing for librecbox_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/recbox_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librecbox_b200 is written for sm_100a (B200) only"
#endif

int rbx_fail(int code, const char* fmt, ...);

#define RBX_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) return rbx_fail(RBX_ERR_ARG, __VA_ARGS__);  \
    } while (0)

#define RBX_LAUNCH_CHECK(name)                                                              \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess)                                                             \
            return rbx_fail(RBX_ERR_CUDA, "%s: launch failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

// NVTX range around an entry point (tracing): visible in nsys / ncu --nvtx timelines, a few
// nanoseconds when no tool is attached.
#include <nvtx3/nvToolsExt.h>
struct RbxRange {
    explicit RbxRange(const char* name) { nvtxRangePushA(name); }
    ~RbxRange() { nvtxRangePop(); }
};
#define RBX_RANGE(name) RbxRange rbx_range__(name)

// SM count of the current device, cached per device.
int rbx_sm_count();

// tensor-core pass A of the top-k search (csrc/gemm.cu), used by csrc/topk.cu
int rbx_topk_filter_tc(const float* q, const float* items, int64_t U, int64_t n0, int64_t n1, int D, const float* tau, int* count,
                       unsigned long long* queue, int64_t qcap, cudaStream_t st);

static inline cudaStream_t rbx_cast_stream(rbx_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// 16-byte read-only load that is allowed to live in L1 (hot embedding rows repeat under Zipf).
__device__ __forceinline__ float4 ld_row_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// 16-byte streaming load: read once, do not pollute L1 (upstream grads, saved activations).
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float ld_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// 16-byte streaming store (written once, consumed by a later kernel from L2/HBM).
__device__ __forceinline__ void st_stream_f4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// Vector reduction into global memory: one 16-byte red instead of four scalar ones (sm_90+).
__device__ __forceinline__ void red_add_f4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void red_add_f1(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- L2 eviction-priority hints (createpolicy + .L2::cache_hint) --------------------------------
// The random-access working set (embedding table, its gradient table) is about the size of the
// 126 MB L2; the per-step streams (E, dE: 163 MB each) are larger.  Marking the streams evict_first
// and the tables evict_last keeps table lines resident across the launch (and across steps), so
// repeated rows are served from L2 instead of HBM.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ld_row_f4_hint(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4_hint(const float* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_stream_f4_hint(float* p, float4 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void red_add_f4_hint(float* p, float4 v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) {
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 f4_scale(float4 a, float s) {
    return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float4 f4_fma(float4 a, float s, float4 c) {  // a*s + c
    return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}
__device__ __forceinline__ float4 f4_sqacc(float4 a, float4 q) {  // q + a*a
    return make_float4(fmaf(a.x, a.x, q.x), fmaf(a.y, a.y, q.y), fmaf(a.z, a.z, q.z), fmaf(a.w, a.w, q.w));
}

// Sum over the LANES consecutive lanes of a sub-warp group (LANES power of two <= 32).
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
