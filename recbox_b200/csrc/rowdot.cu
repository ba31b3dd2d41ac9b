// a10: two-tower scores y[b,k] = <u[b,:], v[b,k,:]> and their backward.  HBM-bound
// (4D(1+K) + 4K bytes per user); warp-level reductions, no tensor cores.
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int capped_grid(int64_t warps_needed, int ctas_per_sm) {
    int64_t ctas = (warps_needed + kThreads / 32 - 1) / (kThreads / 32);
    const int64_t cap = (int64_t)rbx_sm_count() * ctas_per_sm;
    if (ctas > cap) ctas = cap;
    return ctas < 1 ? 1 : (int)ctas;
}

// group of LPR lanes per (b,k) pair
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_rowdot_fwd_vec(const float* __restrict__ u, const float* __restrict__ v,
                                                            float* __restrict__ y, int64_t B, int K) {
    constexpr int D = 4 * LPR, PPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int64_t P = B * K;
    for (int64_t base = warp0 * PPW; base < P; base += nwarps * PPW) {
        const int64_t pr = base + gi;
        float acc = 0.f;
        if (pr < P) {
            const int64_t b = pr / K;
            const float4 a = ld_row_f4(u + (size_t)b * D + 4 * lig);
            const float4 c = ld_stream_f4(v + (size_t)pr * D + 4 * lig);
            acc = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, a.w * c.w)));
        }
        acc = group_sum<LPR>(acc);
        if (pr < P && lig == 0) y[pr] = acc;
    }
}

__global__ void __launch_bounds__(kThreads) k_rowdot_fwd_any(const float* __restrict__ u, const float* __restrict__ v,
                                                            float* __restrict__ y, int64_t B, int K, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int64_t P = B * K;
    for (int64_t pr = warp0; pr < P; pr += nwarps) {
        const int64_t b = pr / K;
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(__ldg(u + (size_t)b * D + d), __ldg(v + (size_t)pr * D + d), acc);
        acc = group_sum<32>(acc);
        if (lane == 0) y[pr] = acc;
    }
}

// backward: group per user b walks its K items: dv[b,k,:] = dy[b,k] u[b,:]; du[b,:] = sum_k dy[b,k] v[b,k,:]
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_rowdot_bwd_vec(const float* __restrict__ u, const float* __restrict__ v,
                                                            const float* __restrict__ dy, float* __restrict__ du,
                                                            float* __restrict__ dv, int64_t B, int K) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        if (b >= B) continue;
        const float4 a = ld_row_f4(u + (size_t)b * D + 4 * lig);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float g = __ldg(dy + b * K + k);
            const size_t o = ((size_t)b * K + k) * D + 4 * lig;
            if (du) acc = f4_fma(ld_stream_f4(v + o), g, acc);
            if (dv) st_stream_f4(dv + o, f4_scale(a, g));
        }
        if (du) *reinterpret_cast<float4*>(du + (size_t)b * D + 4 * lig) = acc;
    }
}

__global__ void __launch_bounds__(kThreads) k_rowdot_bwd_any(const float* __restrict__ u, const float* __restrict__ v,
                                                            const float* __restrict__ dy, float* __restrict__ du,
                                                            float* __restrict__ dv, int64_t B, int K, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        for (int d = lane; d < D; d += 32) {
            const float a = __ldg(u + (size_t)b * D + d);
            float acc = 0.f;
            for (int k = 0; k < K; ++k) {
                const float g = __ldg(dy + b * K + k);
                const size_t o = ((size_t)b * K + k) * D + d;
                if (du) acc = fmaf(__ldg(v + o), g, acc);
                if (dv) dv[o] = a * g;
            }
            if (du) du[(size_t)b * D + d] = acc;
        }
    }
}

inline bool vec_ok(int D, const void* a, const void* b, const void* c, const void* d) {
    return D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 &&
           (uintptr_t)c % 16 == 0 && (uintptr_t)d % 16 == 0;
}

}  // namespace

#define RBX_DISPATCH_LPR(D, CALL)                         \
    switch ((D) / 4) {                                    \
        case 1: { constexpr int LPR = 1; CALL; } break;   \
        case 2: { constexpr int LPR = 2; CALL; } break;   \
        case 4: { constexpr int LPR = 4; CALL; } break;   \
        case 8: { constexpr int LPR = 8; CALL; } break;   \
        case 16: { constexpr int LPR = 16; CALL; } break; \
        default: { constexpr int LPR = 32; CALL; } break; \
    }

extern "C" {

int rbx_rowdot_fwd(const float* u, const float* v, float* y, int64_t B, int K, int D, rbx_stream_t stream) {
    const char* who = "rbx_rowdot_fwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && K >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    if (B == 0 || K == 0) return RBX_OK;
    RBX_REQUIRE(u && v && y, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec_ok(D, u, v, nullptr, nullptr)) {
        RBX_DISPATCH_LPR(D, (k_rowdot_fwd_vec<LPR><<<capped_grid((B * K + 32 / LPR - 1) / (32 / LPR), 8), kThreads, 0, st>>>(u, v, y, B, K)));
    } else {
        k_rowdot_fwd_any<<<capped_grid(B * K, 8), kThreads, 0, st>>>(u, v, y, B, K, D);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_rowdot_bwd(const float* u, const float* v, const float* dy, float* du, float* dv, int64_t B, int K, int D,
                   rbx_stream_t stream) {
    const char* who = "rbx_rowdot_bwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && K >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    if (B == 0 || (!du && !dv)) return RBX_OK;
    RBX_REQUIRE(u && v && (dy || K == 0), "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec_ok(D, u, v, du, dv)) {
        RBX_DISPATCH_LPR(D, (k_rowdot_bwd_vec<LPR><<<capped_grid((B + 32 / LPR - 1) / (32 / LPR), 8), kThreads, 0, st>>>(u, v, dy, du, dv, B, K)));
    } else {
        k_rowdot_bwd_any<<<capped_grid(B, 8), kThreads, 0, st>>>(u, v, dy, du, dv, B, K, D);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
