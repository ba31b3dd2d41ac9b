// a5 / a9: plain row gather, scatter-add, and the pooled sequence gather (sum / masked average)
// that never materialises [B, L, D].  Same lane mapping as embed_fm.cu: LPR = D/4 lanes per row
// on the vector path; a generic lane-strided path covers every other D <= RBX_MAX_DIM.
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;

inline int capped_grid(int64_t warps_needed, int ctas_per_sm) {
    int64_t ctas = (warps_needed + kThreads / 32 - 1) / (kThreads / 32);
    const int64_t cap = (int64_t)rbx_sm_count() * ctas_per_sm;
    if (ctas > cap) ctas = cap;
    return ctas < 1 ? 1 : (int)ctas;
}

inline bool vec_ok(int D, const void* a, const void* b, const void* c) {
    return D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 &&
           (uintptr_t)c % 16 == 0;
}

// ----------------------------------------------------------------------------------------- gather
template <int LPR, int U>
__global__ void __launch_bounds__(kThreads) k_gather_vec(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                                        float* __restrict__ out, int64_t N) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * (RPW * U); base < N; base += nwarps * (RPW * U)) {
        int32_t r[U];
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            r[u] = i < N ? __ldg(ids + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r[u] >= 0) v[u] = ld_row_f4(table + (size_t)r[u] * D + 4 * lig);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            if (r[u] >= 0) st_stream_f4(out + (size_t)i * D + 4 * lig, v[u]);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_gather_any(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                                        float* __restrict__ out, int64_t N, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int32_t r = __ldg(ids + i);
        for (int d = lane; d < D; d += 32) out[(size_t)i * D + d] = r >= 0 ? __ldg(table + (size_t)r * D + d) : 0.f;
    }
}

// ------------------------------------------------------------------------------------ scatter-add
template <int LPR, int U>
__global__ void __launch_bounds__(kThreads) k_scatter_vec(const float* __restrict__ g, const int32_t* __restrict__ ids,
                                                         int32_t pad_row, float* __restrict__ gt, int64_t N) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * (RPW * U); base < N; base += nwarps * (RPW * U)) {
        int32_t r[U];
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            r[u] = i < N ? __ldg(ids + i) : -1;
            if (r[u] == pad_row) r[u] = -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            if (r[u] >= 0) v[u] = ld_stream_f4(g + (size_t)i * D + 4 * lig);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r[u] >= 0) red_add_f4(gt + (size_t)r[u] * D + 4 * lig, v[u]);
    }
}

__global__ void __launch_bounds__(kThreads) k_scatter_any(const float* __restrict__ g, const int32_t* __restrict__ ids,
                                                         int32_t pad_row, float* __restrict__ gt, int64_t N, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int32_t r = __ldg(ids + i);
        if (r < 0 || r == pad_row) continue;
        for (int d = lane; d < D; d += 32) red_add_f1(gt + (size_t)r * D + d, g[(size_t)i * D + d]);
    }
}

// ---------------------------------------------------------------------------------- pooled gather
// group (LPR lanes) per sample; L ids walked U at a time.  mode 1 divides by the number of
// positions whose gathered row has a non-zero element sum (sequence.py:10 / pooling.py:29).
// ------------------------------------------------------------------------- deterministic scatter-add
// a reproducible form of aten::embedding_dense_backward.  The ids arrive stably SORTED (sorted_ids, with
// `order` = the original positions, ascending within equal ids); a warp takes position i and works only if i starts a run of
// equal ids: it walks the run in position order, sums the gradient rows in registers and adds the total to the row with a
// plain read-modify-write -- exactly one warp owns a row, no atomics, the same bits on every run.
__global__ void __launch_bounds__(kThreads) k_segment_sum(const float* __restrict__ g, const int64_t* __restrict__ order,
                                                          const int32_t* __restrict__ sorted_ids, int32_t pad_row,
                                                          float* __restrict__ g_table, int64_t N, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t i = warp0; i < N; i += nwarps) {
        const int32_t r = __ldg(sorted_ids + i);
        if (r < 0 || r == pad_row) continue;
        if (i > 0 && __ldg(sorted_ids + i - 1) == r) continue;          // not the first position of its run
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int d = d0 + lane;
            float acc = 0.0f;
            for (int64_t j = i; j < N && __ldg(sorted_ids + j) == r; ++j)
                if (d < D) acc += __ldg(g + (size_t)__ldg(order + j) * D + d);
            if (d < D) g_table[(size_t)r * D + d] += acc;
        }
    }
}

template <int LPR, int U>
__global__ void __launch_bounds__(kThreads) k_pooled_fwd_vec(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                                            int64_t ids_ld, float* __restrict__ out, int64_t out_ld,
                                                            float* __restrict__ cnt, int64_t B, int L, int mode) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        const bool valid = b < B;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float n_nonzero = 0.f;
        for (int l0 = 0; l0 < L; l0 += U) {
            int32_t r[U];
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) r[u] = (valid && l0 + u < L) ? __ldg(ids + b * ids_ld + l0 + u) : -1;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r[u] >= 0) v[u] = ld_row_f4(table + (size_t)r[u] * D + 4 * lig);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc = f4_add(acc, v[u]);
                if (mode == 1) {
                    const float rs = group_sum<LPR>((v[u].x + v[u].y) + (v[u].z + v[u].w));
                    n_nonzero += (rs != 0.f) ? 1.f : 0.f;
                }
            }
        }
        if (valid) {
            float4 o = acc;
            if (mode == 1) {
                const float den = n_nonzero + 1e-12f;
                o = make_float4(acc.x / den, acc.y / den, acc.z / den, acc.w / den);
            }
            *reinterpret_cast<float4*>(out + (size_t)b * out_ld + 4 * lig) = o;
            if (cnt && lig == 0) cnt[b] = n_nonzero;
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_pooled_fwd_any(const float* __restrict__ table, const int32_t* __restrict__ ids,
                                                            int64_t ids_ld, float* __restrict__ out, int64_t out_ld,
                                                            float* __restrict__ cnt, int64_t B, int L, int D, int mode) {
    constexpr int KD = RBX_MAX_DIM / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        float acc[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc[k] = 0.f;
        float n_nonzero = 0.f;
        for (int l = 0; l < L; ++l) {
            const int32_t r = __ldg(ids + b * ids_ld + l);
            float rs = 0.f;
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                const int d = lane + 32 * k;
                if (d < D && r >= 0) {
                    const float e = __ldg(table + (size_t)r * D + d);
                    acc[k] += e;
                    rs += e;
                }
            }
            if (mode == 1) {
                rs = group_sum<32>(rs);
                n_nonzero += (rs != 0.f) ? 1.f : 0.f;
            }
        }
        const float den = n_nonzero + 1e-12f;
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            const int d = lane + 32 * k;
            if (d < D) out[(size_t)b * out_ld + d] = mode == 1 ? acc[k] / den : acc[k];
        }
        if (cnt && lane == 0) cnt[b] = n_nonzero;
    }
}

template <int LPR, int U>
__global__ void __launch_bounds__(kThreads) k_pooled_bwd_vec(const float* __restrict__ g, int64_t g_ld,
                                                            const int32_t* __restrict__ ids, int64_t ids_ld,
                                                            const float* __restrict__ cnt, int32_t pad_row,
                                                            float* __restrict__ gt, int64_t B, int L, int mode) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        if (b >= B) continue;
        float4 gv = ld_stream_f4(g + (size_t)b * g_ld + 4 * lig);
        if (mode == 1) {
            const float den = __ldg(cnt + b) + 1e-12f;
            gv = make_float4(gv.x / den, gv.y / den, gv.z / den, gv.w / den);
        }
        for (int l0 = 0; l0 < L; l0 += U) {
            int32_t r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                r[u] = (l0 + u < L) ? __ldg(ids + b * ids_ld + l0 + u) : -1;
                if (r[u] == pad_row) r[u] = -1;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (r[u] >= 0) red_add_f4(gt + (size_t)r[u] * D + 4 * lig, gv);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_pooled_bwd_any(const float* __restrict__ g, int64_t g_ld,
                                                            const int32_t* __restrict__ ids, int64_t ids_ld,
                                                            const float* __restrict__ cnt, int32_t pad_row,
                                                            float* __restrict__ gt, int64_t B, int L, int D, int mode) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float den = mode == 1 ? __ldg(cnt + b) + 1e-12f : 1.f;
        for (int l = 0; l < L; ++l) {
            const int32_t r = __ldg(ids + b * ids_ld + l);
            if (r < 0 || r == pad_row) continue;
            for (int d = lane; d < D; d += 32) {
                const float x = g[(size_t)b * g_ld + d];
                red_add_f1(gt + (size_t)r * D + d, mode == 1 ? x / den : x);
            }
        }
    }
}

// ------------------------------------------------------------------ pooling of a materialised [B,L,D]
// (MaskedSumPooling / MaskedAveragePooling called directly on an embedding tensor: sequence.py:4-20,
// pooling.py:22-40).  One warp per sample, lane d owns columns d, d+32, ...; mask (uint8 [B,L]) is
// optional -- without it the reference's rule applies (row element-sum != 0).
__global__ void __launch_bounds__(kThreads) k_pool_fwd(const float* __restrict__ emb, const uint8_t* __restrict__ mask,
                                                      float* __restrict__ out, float* __restrict__ cnt, int64_t B, int L,
                                                      int D, int mode) {
    constexpr int KD = RBX_MAX_DIM / 32;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        float acc[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) acc[k] = 0.f;
        float n = 0.f;
        for (int l = 0; l < L; ++l) {
            const float* row = emb + ((size_t)b * L + l) * D;
            float rs = 0.f;
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                const int d = lane + 32 * k;
                if (d < D) {
                    const float e = ld_stream_f1(row + d);
                    acc[k] += e;
                    rs += e;
                }
            }
            if (mode == 1) {
                if (mask) n += mask[b * L + l] ? 1.f : 0.f;
                else n += (group_sum<32>(rs) != 0.f) ? 1.f : 0.f;
            }
        }
        const float den = n + 1e-12f;
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            const int d = lane + 32 * k;
            if (d < D) out[(size_t)b * D + d] = mode == 1 ? acc[k] / den : acc[k];
        }
        if (cnt && lane == 0) cnt[b] = n;
    }
}

// d_emb[b,l,:] = g[b,:] / (cnt[b] + 1e-12)  (mode 1)  |  g[b,:]  (mode 0): the sum's gradient reaches
// every position, masked or not, exactly as autograd does for sequence.py:8-12
__global__ void __launch_bounds__(kThreads) k_pool_bwd(const float* __restrict__ g, const float* __restrict__ cnt,
                                                      float* __restrict__ d_emb, int64_t B, int L, int D, int mode) {
    const int64_t n = B * (int64_t)L * D;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const int64_t b = i / ((int64_t)L * D);
        const int d = (int)(i % D);
        const float x = __ldg(g + b * D + d);
        d_emb[i] = mode == 1 ? x / (__ldg(cnt + b) + 1e-12f) : x;
    }
}

// Vector forms of the two materialised pooling kernels (D = 4*LPR a power of two <= 128, 16-byte aligned).  The scalar
// kernels above walk the L rows one at a time with one 4-byte load per lane and a full-warp shuffle reduction per row
// (0.24 of the HBM roofline on [8192,200,64], profiles/r1u_kernel_rooflines.md).  Here LPR lanes own a row, a warp step
// covers 32/LPR rows and kU steps are in flight; the row-sum test of sequence.py:10 is a log2(LPR)-step shuffle.
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_pool_fwd_vec(const float* __restrict__ emb, const uint8_t* __restrict__ mask,
                                                          float* __restrict__ out, float* __restrict__ cnt, int64_t B, int L,
                                                          int mode) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR, kU = 4;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), ri = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        const float* eb = emb + (size_t)b * L * D + 4 * lig;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float n = 0.f;
        for (int l0 = 0; l0 < L; l0 += kU * RPW) {
            float4 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int l = l0 + u * RPW + ri;
                v[u] = l < L ? ld_stream_f4(eb + (size_t)l * D) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int l = l0 + u * RPW + ri;
                acc = f4_add(acc, v[u]);
                if (mode == 1) {
                    if (mask) {
                        if (l < L && lig == 0 && mask[b * L + l]) n += 1.f;
                    } else {
                        const float rs = group_sum<LPR>((v[u].x + v[u].y) + (v[u].z + v[u].w));
                        if (l < L && lig == 0 && rs != 0.f) n += 1.f;
                    }
                }
            }
        }
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        n = group_sum<32>(n);
        if (ri == 0) {
            const float den = n + 1e-12f;                 // sequence.py:11 divides (no reciprocal-multiply)
            float4 r = acc;
            if (mode == 1) r = make_float4(acc.x / den, acc.y / den, acc.z / den, acc.w / den);
            *reinterpret_cast<float4*>(out + (size_t)b * D + 4 * lig) = r;
        }
        if (cnt && lane == 0) cnt[b] = n;
    }
}

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_pool_bwd_vec(const float* __restrict__ g, const float* __restrict__ cnt,
                                                          float* __restrict__ d_emb, int64_t B, int L, int mode) {
    constexpr int D = 4 * LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1);
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const int n4 = L * LPR;                       // float4 per sample; lane t, t+32, ... all hit column chunk `lig`
    for (int64_t b = warp0; b < B; b += nwarps) {
        float4 x = ld_row_f4(g + (size_t)b * D + 4 * lig);
        if (mode == 1) {
            const float den = __ldg(cnt + b) + 1e-12f;
            x = make_float4(x.x / den, x.y / den, x.z / den, x.w / den);
        }
        float* db = d_emb + (size_t)b * L * D;
#pragma unroll 4
        for (int t = lane; t < n4; t += 32) st_stream_f4(db + 4 * (size_t)t, x);
    }
}

}  // namespace

#define RBX_DISPATCH_LPR(D, CALL)                   \
    switch ((D) / 4) {                              \
        case 1: { constexpr int LPR = 1; CALL; } break;   \
        case 2: { constexpr int LPR = 2; CALL; } break;   \
        case 4: { constexpr int LPR = 4; CALL; } break;   \
        case 8: { constexpr int LPR = 8; CALL; } break;   \
        case 16: { constexpr int LPR = 16; CALL; } break; \
        default: { constexpr int LPR = 32; CALL; } break; \
    }

extern "C" {

int rbx_gather_rows(const float* table, const int32_t* ids, float* out, int64_t N, int D, rbx_stream_t stream) {
    const char* who = "rbx_gather_rows";
    RBX_RANGE(who);
    RBX_REQUIRE(N >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    if (N == 0) return RBX_OK;
    RBX_REQUIRE(table && ids && out, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec_ok(D, table, out, nullptr)) {
        RBX_DISPATCH_LPR(D, (k_gather_vec<LPR, 4><<<capped_grid((N + (32 / LPR) * 4 - 1) / ((32 / LPR) * 4), 8), kThreads, 0, st>>>(
                                table, ids, out, N)));
    } else {
        k_gather_any<<<capped_grid(N, 8), kThreads, 0, st>>>(table, ids, out, N, D);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_scatter_add_rows(const float* g, const int32_t* ids, int32_t pad_row, float* g_table, int64_t N, int D,
                         rbx_stream_t stream) {
    const char* who = "rbx_scatter_add_rows";
    RBX_RANGE(who);
    RBX_REQUIRE(N >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    if (N == 0) return RBX_OK;
    RBX_REQUIRE(g && ids && g_table, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec_ok(D, g, g_table, nullptr)) {
        RBX_DISPATCH_LPR(D, (k_scatter_vec<LPR, 4><<<capped_grid((N + (32 / LPR) * 4 - 1) / ((32 / LPR) * 4), 8), kThreads, 0, st>>>(
                                g, ids, pad_row, g_table, N)));
    } else {
        k_scatter_any<<<capped_grid(N, 8), kThreads, 0, st>>>(g, ids, pad_row, g_table, N, D);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_segment_sum_rows(const float* g, const int64_t* order, const int32_t* sorted_ids, int32_t pad_row, float* g_table,
                         int64_t N, int D, rbx_stream_t stream) {
    const char* who = "rbx_segment_sum_rows";
    RBX_RANGE(who);
    RBX_REQUIRE(N >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    if (N == 0) return RBX_OK;
    RBX_REQUIRE(g && order && sorted_ids && g_table, "%s: null pointer", who);
    k_segment_sum<<<capped_grid((N + kThreads / 32 - 1) / (kThreads / 32), 8), kThreads, 0, rbx_cast_stream(stream)>>>(g, order, sorted_ids,
                                                                                                              pad_row, g_table, N, D);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_pooled_gather_fwd(const float* table, const int32_t* ids, int64_t ids_ld, float* out, int64_t out_ld,
                          float* cnt, int64_t B, int L, int D, int mode, rbx_stream_t stream) {
    const char* who = "rbx_pooled_gather_fwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && L >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode == 0 || mode == 1, "%s: mode %d", who, mode);
    RBX_REQUIRE(ids_ld >= L && out_ld >= D, "%s: leading dimension too small", who);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(table && ids && out, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec_ok(D, table, out, nullptr) && out_ld % 4 == 0) {
        RBX_DISPATCH_LPR(D, (k_pooled_fwd_vec<LPR, 8><<<capped_grid((B + 32 / LPR - 1) / (32 / LPR), 4), kThreads, 0, st>>>(
                                table, ids, ids_ld, out, out_ld, cnt, B, L, mode)));
    } else {
        k_pooled_fwd_any<<<capped_grid(B, 4), kThreads, 0, st>>>(table, ids, ids_ld, out, out_ld, cnt, B, L, D, mode);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_pooled_gather_bwd(const float* g, int64_t g_ld, const int32_t* ids, int64_t ids_ld, const float* cnt,
                          int32_t pad_row, float* g_table, int64_t B, int L, int D, int mode, rbx_stream_t stream) {
    const char* who = "rbx_pooled_gather_bwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && L >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode == 0 || mode == 1, "%s: mode %d", who, mode);
    RBX_REQUIRE(mode == 0 || cnt, "%s: cnt (saved by the forward) required for masked average", who);
    RBX_REQUIRE(ids_ld >= L && g_ld >= D, "%s: leading dimension too small", who);
    if (B == 0 || L == 0) return RBX_OK;
    RBX_REQUIRE(g && ids && g_table, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec_ok(D, g, g_table, nullptr) && g_ld % 4 == 0) {
        RBX_DISPATCH_LPR(D, (k_pooled_bwd_vec<LPR, 8><<<capped_grid((B + 32 / LPR - 1) / (32 / LPR), 4), kThreads, 0, st>>>(
                                g, g_ld, ids, ids_ld, cnt, pad_row, g_table, B, L, mode)));
    } else {
        k_pooled_bwd_any<<<capped_grid(B, 4), kThreads, 0, st>>>(g, g_ld, ids, ids_ld, cnt, pad_row, g_table, B, L, D, mode);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_pool_fwd(const float* emb, const uint8_t* mask, float* out, float* cnt, int64_t B, int L, int D, int mode,
                 rbx_stream_t stream) {
    const char* who = "rbx_pool_fwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && L >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode == 0 || mode == 1, "%s: mode %d", who, mode);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE((emb || L == 0) && out, "%s: null pointer", who);
    if (L > 0 && vec_ok(D, emb, out, nullptr)) {
        RBX_DISPATCH_LPR(D, (k_pool_fwd_vec<LPR><<<capped_grid(B, 8), kThreads, 0, rbx_cast_stream(stream)>>>(emb, mask, out, cnt, B, L, mode)));
    } else {
        k_pool_fwd<<<capped_grid(B, 8), kThreads, 0, rbx_cast_stream(stream)>>>(emb, mask, out, cnt, B, L, D, mode);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_pool_bwd(const float* g, const float* cnt, float* d_emb, int64_t B, int L, int D, int mode, rbx_stream_t stream) {
    const char* who = "rbx_pool_bwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && L >= 0 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode == 0 || (mode == 1 && cnt), "%s: mode %d / cnt", who, mode);
    if (B == 0 || L == 0) return RBX_OK;
    RBX_REQUIRE(g && d_emb, "%s: null pointer", who);
    if (vec_ok(D, g, d_emb, nullptr)) {
        RBX_DISPATCH_LPR(D, (k_pool_bwd_vec<LPR><<<capped_grid(B, 8), kThreads, 0, rbx_cast_stream(stream)>>>(g, cnt, d_emb, B, L, mode)));
        RBX_LAUNCH_CHECK(who);
        return RBX_OK;
    }
    const int64_t n = B * (int64_t)L * D;
    int64_t grid = (n + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)rbx_sm_count() * 16;
    if (grid > cap) grid = cap;
    k_pool_bwd<<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(g, cnt, d_emb, B, L, D, mode);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
