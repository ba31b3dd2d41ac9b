// e (replicas): in-switch all-reduce of the fused gradient buffer over NVLink 5 / NVSwitch multicast memory (NVLS).
//
// Replicas of a table that fits every GPU synchronise its dense gradient once per step (what
// DistributedDataParallel does for the reference's RecBole / rechub trainers, third_party/recbole/trainer/trainer.py:60-64).
// NCCL's all-reduce of the 68 MB buffer takes 284 us on 8 GPUs (425 GB/s bus bandwidth, measured r2ac) -- more than the
// forward + backward it follows (190 us).  Here every rank owns 1/world of the buffer and runs ONE kernel over its slice:
//     v = multimem.ld_reduce.add(mc + i)      the switch reads the 16 bytes from every GPU's copy and adds them
//     multimem.st(mc + i, v)                  the switch writes the sum into every GPU's copy
// so a GPU sends and receives the buffer once (1/world of it as reduce requests, all of it as the broadcast it receives).
// The buffer is CUDA-VMM symmetric memory with a multicast mapping (torch.distributed._symmetric_memory provides the
// allocation, the multicast pointer and the cross-rank barriers: plumbing); the data path is this kernel.
#include "rbx_common.cuh"

namespace {

__device__ __forceinline__ float4 mc_ld_reduce(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// four reduce requests in flight per thread before the first broadcast store (a request crosses the switch twice)
__global__ void __launch_bounds__(512) k_nvls_allreduce(float* __restrict__ mc, int64_t n4_begin, int64_t n4_end) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = n4_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4_end; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = mc_ld_reduce(mc + 4 * (i + u * stride));
#pragma unroll
        for (int u = 0; u < 4; ++u) mc_st(mc + 4 * (i + u * stride), v[u]);
    }
    for (; i < n4_end; i += stride) mc_st(mc + 4 * i, mc_ld_reduce(mc + 4 * i));
}

}  // namespace

extern "C" {

// mc = multicast address of the symmetric buffer (n floats, n % 4 == 0, 16-byte aligned); every rank calls it between two
// cross-rank barriers (all contributions written before, all sums visible after).  Rank r reduces floats [r*n/world, (r+1)*n/world).
int rbx_nvls_allreduce_f32(float* mc, int64_t n, int rank, int world, rbx_stream_t stream) {
    const char* who = "rbx_nvls_allreduce_f32";
    RBX_RANGE(who);
    RBX_REQUIRE(mc && n >= 0 && (n & 3) == 0 && ((uintptr_t)mc & 15) == 0, "%s: buffer must be a 16-byte aligned multiple of 4 floats", who);
    RBX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "%s: bad rank / world", who);
    const int64_t n4 = n / 4, per = (n4 + world - 1) / world;
    const int64_t b = per * rank < n4 ? per * rank : n4, e = b + per < n4 ? b + per : n4;
    if (e <= b) return RBX_OK;
    int64_t grid = (e - b + 511) / 512;
    const int64_t cap = (int64_t)rbx_sm_count() * 2;
    if (grid > cap) grid = cap;
    k_nvls_allreduce<<<(unsigned)grid, 512, 0, rbx_cast_stream(stream)>>>(mc, b, e);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
