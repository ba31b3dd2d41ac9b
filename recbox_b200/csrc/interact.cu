// a6: the four InnerProductInteraction modes on a materialised E [B,F,D]
// (ranking/pytorch/layers/interactions/inner_product.py:40-56) and their backward.
//   0 product_sum / 1 bi_interaction : streaming, group-of-lanes per sample, sums in registers
//   2 inner_product / 3 elementwise_product : one CTA per sample, E[b] staged in shared memory
//     (row stride D+1 floats -> conflict-free when lanes read different fields)
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int capped_grid(int64_t ctas, int ctas_per_sm) {
    const int64_t cap = (int64_t)rbx_sm_count() * ctas_per_sm;
    if (ctas > cap) ctas = cap;
    return ctas < 1 ? 1 : (int)ctas;
}

// ------------------------------------------------------------------ modes 0/1, vector path
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_sumsq_fwd_vec(const float* __restrict__ E, float* __restrict__ out, int64_t B,
                                                           int F, int mode) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        const bool valid = b < B;
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;
        if (valid) {
            const float* e = E + (size_t)b * F * D + 4 * lig;
#pragma unroll 8
            for (int f = 0; f < F; ++f) {
                const float4 v = ld_stream_f4(e + (size_t)f * D);
                S = f4_add(S, v);
                Q = f4_sqacc(v, Q);
            }
        }
        const float4 bi = make_float4((S.x * S.x - Q.x) * 0.5f, (S.y * S.y - Q.y) * 0.5f, (S.z * S.z - Q.z) * 0.5f,
                                      (S.w * S.w - Q.w) * 0.5f);
        if (mode == 1) {
            if (valid) *reinterpret_cast<float4*>(out + (size_t)b * D + 4 * lig) = bi;
        } else {
            const float s = group_sum<LPR>((bi.x + bi.y) + (bi.z + bi.w));
            if (valid && lig == 0) out[b] = s;
        }
    }
}

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_sumsq_bwd_vec(const float* __restrict__ E, const float* __restrict__ dout,
                                                           float* __restrict__ dE, int64_t B, int F, int mode) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        if (b >= B) continue;
        const float* e = E + (size_t)b * F * D + 4 * lig;
        float* de = dE + (size_t)b * F * D + 4 * lig;
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int f = 0; f < F; ++f) S = f4_add(S, ld_row_f4(e + (size_t)f * D));
        float4 g;
        if (mode == 1) g = ld_stream_f4(dout + (size_t)b * D + 4 * lig);
        else { const float s = __ldg(dout + b); g = make_float4(s, s, s, s); }
#pragma unroll 8
        for (int f = 0; f < F; ++f) {
            const float4 v = ld_row_f4(e + (size_t)f * D);   // second touch: L1/L2 hit
            st_stream_f4(de + (size_t)f * D, make_float4(g.x * (S.x - v.x), g.y * (S.y - v.y), g.z * (S.z - v.z), g.w * (S.w - v.w)));
        }
    }
}

// ------------------------------------------------------------------ modes 0/1, any D: warp per sample
__global__ void __launch_bounds__(kThreads) k_sumsq_fwd_any(const float* __restrict__ E, float* __restrict__ out, int64_t B,
                                                           int F, int D, int mode) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        float tot = 0.f;
        for (int d = lane; d < D; d += 32) {
            float S = 0.f, Q = 0.f;
            for (int f = 0; f < F; ++f) {
                const float v = __ldg(E + ((size_t)b * F + f) * D + d);
                S += v;
                Q = fmaf(v, v, Q);
            }
            const float bi = (S * S - Q) * 0.5f;
            if (mode == 1) out[(size_t)b * D + d] = bi;
            tot += bi;
        }
        if (mode == 0) {
            tot = group_sum<32>(tot);
            if (lane == 0) out[b] = tot;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_sumsq_bwd_any(const float* __restrict__ E, const float* __restrict__ dout,
                                                           float* __restrict__ dE, int64_t B, int F, int D, int mode) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        for (int d = lane; d < D; d += 32) {
            float S = 0.f;
            for (int f = 0; f < F; ++f) S += __ldg(E + ((size_t)b * F + f) * D + d);
            const float g = mode == 1 ? __ldg(dout + (size_t)b * D + d) : __ldg(dout + b);
            for (int f = 0; f < F; ++f) {
                const size_t o = ((size_t)b * F + f) * D + d;
                dE[o] = g * (S - __ldg(E + o));
            }
        }
    }
}

// ------------------------------------------------------------------ modes 2/3: CTA per sample
__device__ __forceinline__ int pair_index(int i, int j, int F) {  // i < j, row-major upper triangle
    return i * (2 * F - i - 1) / 2 + (j - i - 1);
}

// dynamic smem: sE[F][D+1] floats, then sPair[P] packed (i<<16|j) for the forward
__global__ void __launch_bounds__(kThreads) k_pairs_fwd(const float* __restrict__ E, float* __restrict__ out, int64_t B, int F,
                                                       int D, int mode) {
    extern __shared__ float smem[];
    const int ldE = D + 1, P = F * (F - 1) / 2;
    float* sE = smem;
    int* sPair = reinterpret_cast<int*>(smem + (size_t)F * ldE);
    for (int i = threadIdx.x; i < F; i += kThreads)
        for (int j = i + 1; j < F; ++j) sPair[pair_index(i, j, F)] = (i << 16) | j;
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < F * D; t += kThreads) sE[(t / D) * ldE + (t % D)] = ld_stream_f1(E + (size_t)b * F * D + t);
        __syncthreads();
        if (mode == 2) {
            for (int p = threadIdx.x; p < P; p += kThreads) {
                const int ij = sPair[p];
                const float* a = sE + (ij >> 16) * ldE;
                const float* c = sE + (ij & 0xffff) * ldE;
                float acc = 0.f;
                for (int d = 0; d < D; ++d) acc = fmaf(a[d], c[d], acc);
                out[(size_t)b * P + p] = acc;
            }
        } else {
            const int64_t tot = (int64_t)P * D;
            for (int64_t t = threadIdx.x; t < tot; t += kThreads) {
                const int p = (int)(t / D), d = (int)(t % D);
                const int ij = sPair[p];
                out[(size_t)b * tot + t] = sE[(ij >> 16) * ldE + d] * sE[(ij & 0xffff) * ldE + d];
            }
        }
    }
}

// backward: thread <-> (i,d): dE[b,i,d] = sum_{j != i} w(i,j)[d] * e[j,d];
// mode 2: w = dout[b,p(i,j)] (staged in smem) ; mode 3: w = dout[b,p(i,j),d] (global, each read twice)
__global__ void __launch_bounds__(kThreads) k_pairs_bwd(const float* __restrict__ E, const float* __restrict__ dout,
                                                       float* __restrict__ dE, int64_t B, int F, int D, int mode) {
    extern __shared__ float smem[];
    const int ldE = D + 1, P = F * (F - 1) / 2;
    float* sE = smem;
    float* sG = smem + (size_t)F * ldE;   // mode 2 only: dout[b, :]
    for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        for (int t = threadIdx.x; t < F * D; t += kThreads) sE[(t / D) * ldE + (t % D)] = ld_stream_f1(E + (size_t)b * F * D + t);
        if (mode == 2)
            for (int p = threadIdx.x; p < P; p += kThreads) sG[p] = ld_stream_f1(dout + (size_t)b * P + p);
        __syncthreads();
        for (int t = threadIdx.x; t < F * D; t += kThreads) {
            const int i = t / D, d = t % D;
            float acc = 0.f;
            for (int j = 0; j < F; ++j) {
                if (j == i) continue;
                const int p = i < j ? pair_index(i, j, F) : pair_index(j, i, F);
                const float w = mode == 2 ? sG[p] : __ldg(dout + ((size_t)b * P + p) * D + d);
                acc = fmaf(w, sE[j * ldE + d], acc);
            }
            dE[(size_t)b * F * D + t] = acc;
        }
    }
}

inline bool vec_ok(int D, const void* a, const void* b, const void* c) {
    return D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 &&
           (uintptr_t)c % 16 == 0;
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

int rbx_interact_fwd(const float* E, float* out, int64_t B, int F, int D, int mode, rbx_stream_t stream) {
    const char* who = "rbx_interact_fwd";
    RBX_REQUIRE(B >= 0 && F >= 1 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode >= 0 && mode <= 3, "%s: InnerProductInteraction mode %d is not supported", who, mode);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(E && out, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (mode <= 1) {
        if (vec_ok(D, E, mode == 1 ? out : nullptr, nullptr)) {
            RBX_DISPATCH_LPR(D, (k_sumsq_fwd_vec<LPR><<<capped_grid((B + 32 / LPR * 8 - 1) / (32 / LPR * 8), 8), kThreads, 0, st>>>(E, out, B, F, mode)));
        } else {
            k_sumsq_fwd_any<<<capped_grid((B + 7) / 8, 8), kThreads, 0, st>>>(E, out, B, F, D, mode);
        }
    } else {
        if (F < 2) return RBX_OK;
        RBX_REQUIRE(F < 32768, "%s: F too large", who);
        const size_t P = (size_t)F * (F - 1) / 2;
        const size_t smem = ((size_t)F * (D + 1) + P) * 4;
        RBX_REQUIRE(smem <= 200 * 1024, "%s: F=%d D=%d needs %zu B shared memory", who, F, D, smem);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pairs_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pairs_fwd<<<capped_grid(B, 6), kThreads, smem, st>>>(E, out, B, F, D, mode);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_interact_bwd(const float* E, const float* dout, float* dE, int64_t B, int F, int D, int mode,
                     rbx_stream_t stream) {
    const char* who = "rbx_interact_bwd";
    RBX_REQUIRE(B >= 0 && F >= 1 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode >= 0 && mode <= 3, "%s: InnerProductInteraction mode %d is not supported", who, mode);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(E && dE && (dout || F < 2), "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (mode <= 1) {
        if (vec_ok(D, E, dE, mode == 1 ? dout : nullptr)) {
            RBX_DISPATCH_LPR(D, (k_sumsq_bwd_vec<LPR><<<capped_grid((B + 32 / LPR * 8 - 1) / (32 / LPR * 8), 8), kThreads, 0, st>>>(E, dout, dE, B, F, mode)));
        } else {
            k_sumsq_bwd_any<<<capped_grid((B + 7) / 8, 8), kThreads, 0, st>>>(E, dout, dE, B, F, D, mode);
        }
    } else {
        if (F < 2) {
            cudaMemsetAsync(dE, 0, (size_t)B * F * D * 4, st);
            return RBX_OK;
        }
        const size_t P = (size_t)F * (F - 1) / 2;
        const size_t smem = ((size_t)F * (D + 1) + (mode == 2 ? P : 0)) * 4;
        RBX_REQUIRE(smem <= 200 * 1024, "%s: F=%d D=%d needs %zu B shared memory", who, F, D, smem);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pairs_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pairs_bwd<<<capped_grid(B, 6), kThreads, smem, st>>>(E, dout, dE, B, F, D, mode);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
