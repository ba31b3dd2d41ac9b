// a6: the four InnerProductInteraction modes on a materialised E [B,F,D]
// (ranking/pytorch/layers/interactions/inner_product.py:40-56) and their backward.
//   0 product_sum / 1 bi_interaction : streaming, group-of-lanes per sample, sums in registers
//   2 inner_product / 3 elementwise_product : one CTA per sample, E[b] staged in shared memory
//     (row stride D+1 floats -> conflict-free when lanes read different fields)
#include <stdlib.h>
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

// ------------------------------------------------------------------ modes 0/1, register-resident form
// A sample is spread over LPR (column chunks) x FG (field groups) lanes; lane (c, fg) holds the 16-byte chunk c of the
// fields fg, fg + FG, ... in registers (R of them), so E is read from HBM exactly ONCE also in the backward (the plain
// kernels above re-read it for the second pass: 0.46 of the roofline) and all of a lane's loads are independent.
template <int LPR>
struct SumsqMap {
    static constexpr int FG = (32 / LPR) < 4 ? (32 / LPR) : 4;     // field groups per sample
    static constexpr int SPW = 32 / (LPR * FG);                    // samples per warp
};

template <int LPR, int R>
__global__ void __launch_bounds__(kThreads) k_sumsq_reg(const float* __restrict__ E, const float* __restrict__ dout,
                                                       float* __restrict__ out, int64_t B, int F, int mode, int bwd) {
    constexpr int D = 4 * LPR, FG = SumsqMap<LPR>::FG, SPW = SumsqMap<LPR>::SPW;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), fg = (lane / LPR) & (FG - 1), sg = lane / (LPR * FG);
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + sg;
        const bool valid = b < B;
        const float* e = E + (size_t)(valid ? b : 0) * F * D + 4 * lig;
        float4 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int f = i * FG + fg;
            v[i] = (valid && f < F) ? ld_stream_f4(e + (size_t)f * D) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;
#pragma unroll
        for (int i = 0; i < R; ++i) {
            S = f4_add(S, v[i]);
            Q = f4_sqacc(v[i], Q);
        }
#pragma unroll
        for (int o = LPR; o < LPR * FG; o <<= 1) {                // sum over the field groups (same column chunk)
            S.x += __shfl_xor_sync(0xffffffffu, S.x, o); S.y += __shfl_xor_sync(0xffffffffu, S.y, o);
            S.z += __shfl_xor_sync(0xffffffffu, S.z, o); S.w += __shfl_xor_sync(0xffffffffu, S.w, o);
            Q.x += __shfl_xor_sync(0xffffffffu, Q.x, o); Q.y += __shfl_xor_sync(0xffffffffu, Q.y, o);
            Q.z += __shfl_xor_sync(0xffffffffu, Q.z, o); Q.w += __shfl_xor_sync(0xffffffffu, Q.w, o);
        }
        if (!bwd) {
            const float4 bi = make_float4((S.x * S.x - Q.x) * 0.5f, (S.y * S.y - Q.y) * 0.5f, (S.z * S.z - Q.z) * 0.5f,
                                          (S.w * S.w - Q.w) * 0.5f);
            if (mode == 1) {
                if (valid && fg == 0) *reinterpret_cast<float4*>(out + (size_t)b * D + 4 * lig) = bi;
            } else {
                const float s = group_sum<LPR>((bi.x + bi.y) + (bi.z + bi.w));
                if (valid && fg == 0 && lig == 0) out[b] = s;
            }
        } else if (valid) {
            float4 g;
            if (mode == 1) g = ld_stream_f4(dout + (size_t)b * D + 4 * lig);
            else { const float s = __ldg(dout + b); g = make_float4(s, s, s, s); }
            float* de = out + (size_t)b * F * D + 4 * lig;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int f = i * FG + fg;
                if (f < F)
                    st_stream_f4(de + (size_t)f * D, make_float4(g.x * (S.x - v[i].x), g.y * (S.y - v[i].y), g.z * (S.z - v[i].z),
                                                                 g.w * (S.w - v[i].w)));
            }
        }
    }
}

// launch the register-resident kernel when the sample fits (F <= FG * 16); returns false otherwise
template <int LPR>
bool launch_sumsq_reg(const float* E, const float* dout, float* out, int64_t B, int F, int mode, int bwd, cudaStream_t st) {
    constexpr int FG = SumsqMap<LPR>::FG, SPW = SumsqMap<LPR>::SPW;
    const int need = (F + FG - 1) / FG;
    const int grid = capped_grid((B + SPW * 8 - 1) / (SPW * 8), 8);
    if (need <= 4) k_sumsq_reg<LPR, 4><<<grid, kThreads, 0, st>>>(E, dout, out, B, F, mode, bwd);
    else if (need <= 8) k_sumsq_reg<LPR, 8><<<grid, kThreads, 0, st>>>(E, dout, out, B, F, mode, bwd);
    else if (need <= 12) k_sumsq_reg<LPR, 12><<<grid, kThreads, 0, st>>>(E, dout, out, B, F, mode, bwd);
    else if (need <= 16) k_sumsq_reg<LPR, 16><<<grid, kThreads, 0, st>>>(E, dout, out, B, F, mode, bwd);
    else return false;
    return true;
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

// ------------------------------------------------------------------ modes 2/3, vector path: WARP per sample
// (D = 4*LPR a power of two <= 128).  The CTA-per-sample kernels below synchronise the whole CTA twice per sample,
// divide by D per element and re-derive pair indices in the inner loop; they ran at 0.04-0.3 of the HBM roofline
// (profiles/r1u_kernel_rooflines.md).  Here every warp owns a sample: E[b] is staged in the warp's private slice of
// shared memory with coalesced 16-byte loads, and the pair work is register-tiled so shared-memory bandwidth stays
// below the HBM time.  No CTA-wide barrier after the setup.
//
// Slice layout: row r of the sample at float offset r*D + 4*(r>>2) -- the extra 16 B per 4-row block spreads the
// blocks over the eight 16-byte bank groups, so an LDS.128 of column chunk c of rows {4t+k : t = lane's block} is
// conflict-free.  Rows F .. 4*ceil(F/4)-1 are zero padding (tiles read them, results are not written).
__device__ __forceinline__ int row_off(int r, int D) { return r * D + ((r >> 2) << 2); }
__host__ __device__ inline int slice_floats(int F, int D) { const int FB = (F + 3) >> 2; return 4 * FB * D + 4 * FB; }

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); return fmaf(a.w, b.w, acc);
}
// Blackwell's packed fp32 FMA (FFMA2, fma.rn.f32x2): the two halves of a 16-byte chunk go through ONE issue slot.
// dot4x2 keeps an (even-k, odd-k) pair of partial sums; the pair is folded once per output at the end.
__device__ __forceinline__ float2 dot4x2(float4 a, float4 b, float2 acc) {
    acc = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), acc);
    return __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), acc);
}
__device__ __forceinline__ float4 f4_fma_x2(float4 a, float s, float4 c) {  // a*s + c, two FFMA2
    const float2 ss = make_float2(s, s);
    const float2 lo = __ffma2_rn(make_float2(a.x, a.y), ss, make_float2(c.x, c.y));
    const float2 hi = __ffma2_rn(make_float2(a.z, a.w), ss, make_float2(c.z, c.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_fma4(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

template <int LPR>
__device__ __forceinline__ void stage_sample(float* sE, const float* __restrict__ Eb, int F, int lane) {
    constexpr int D = 4 * LPR;
    for (int t = lane; t < F * LPR; t += 32) {
        const int r = t / LPR, c = t & (LPR - 1);
        *reinterpret_cast<float4*>(sE + row_off(r, D) + 4 * c) = ld_stream_f4(Eb + 4 * (size_t)t);
    }
}

template <int LPR>
__device__ __forceinline__ void zero_pad_rows(float* sE, int F, int lane) {
    constexpr int D = 4 * LPR;
    const int FB = (F + 3) >> 2;
    for (int t = F * LPR + lane; t < 4 * FB * LPR; t += 32)
        *reinterpret_cast<float4*>(sE + row_off(t / LPR, D) + 4 * (t & (LPR - 1))) = make_float4(0.f, 0.f, 0.f, 0.f);
}

// mode 2 forward: out[b, p(i,j)] = <e_i, e_j>.  A lane owns a 4x4 tile of the upper triangle (tile table shared by the
// CTA): 8 LDS.128 per 64 FMA.
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_ip_fwd_warp(const float* __restrict__ E, float* __restrict__ out, int64_t B, int F) {
    constexpr int D = 4 * LPR;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int FB = (F + 3) >> 2, NT = FB * (FB + 1) / 2, P = F * (F - 1) / 2;
    int* sTile = reinterpret_cast<int*>(smem);                       // [NT] (ti << 8) | tj, ti <= tj
    float* sE = smem + ((NT + 3) & ~3) + (size_t)warp * slice_floats(F, D);
    for (int ti = threadIdx.x; ti < FB; ti += blockDim.x)
        for (int tj = ti; tj < FB; ++tj) sTile[ti * FB - ti * (ti - 1) / 2 + (tj - ti)] = (ti << 8) | tj;
    zero_pad_rows<LPR>(sE, F, lane);
    __syncthreads();
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp, nw = (int64_t)gridDim.x * wpc;
    for (int64_t b = gw; b < B; b += nw) {
        __syncwarp();
        stage_sample<LPR>(sE, E + (size_t)b * F * D, F, lane);
        __syncwarp();
        float* ob = out + (size_t)b * P;
        for (int t = lane; t < NT; t += 32) {
            const int ti = sTile[t] >> 8, tj = sTile[t] & 0xff;
            float2 acc2[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc2[x][y] = make_float2(0.f, 0.f);
            const float* pa = sE + row_off(4 * ti, D);
            const float* pb = sE + row_off(4 * tj, D);
#pragma unroll 4
            for (int c = 0; c < LPR; ++c) {
                float4 a[4], bb[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    a[k] = *reinterpret_cast<const float4*>(pa + k * D + 4 * c);
                    bb[k] = *reinterpret_cast<const float4*>(pb + k * D + 4 * c);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc2[x][y] = dot4x2(a[x], bb[y], acc2[x][y]);
            }
            float acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = acc2[x][y].x + acc2[x][y].y;
            // pair_index(i, j) = base(i) + j with base(i) = i (2F - i - 1) / 2 - i - 1
            if (tj > ti && 4 * tj + 3 < F) {
                // interior tile: all 16 pairs exist, no tests (the epilogue was half of the kernel's instructions, ncu r1z)
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = 4 * ti + x;
                    float* o = ob + (i * (2 * F - i - 1) / 2 - i - 1) + 4 * tj;
#pragma unroll
                    for (int y = 0; y < 4; ++y) o[y] = acc[x][y];
                }
            } else {
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int i = 4 * ti + x;
                    if (i >= F - 1) continue;
                    const int base = i * (2 * F - i - 1) / 2 - i - 1;
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int j = 4 * tj + y;
                        if (j > i && j < F) ob[base + j] = acc[x][y];
                    }
                }
            }
        }
    }
}

// (Round 2, measured and dropped: packing the FMAs of this kernel and of the backward into FFMA2 changed nothing -- 122.9 /
// 192.5 us before and after, r2am --, and a variant with two samples per warp and 8x8 register tiles, i.e. half the LDS.128
// wavefronts per sample and all lanes busy in one round, was SLOWER: 190 us at 128 registers, r2an.  The kernel is bound by the
// latency of its dependent LDS -> FMA -> STG chains at the occupancy its shared-memory slices allow, not by FMA issue or
// shared-memory bandwidth.)

// mode 2 backward: dE[b,i,:] = sum_j G[i,j] e_j with G the symmetric, zero-diagonal matrix of dout[b, p(i,j)].
// G is scattered into the warp's slice as sG[j][i] (row stride GS).  A lane owns RPT consecutive rows i x one 16-byte
// column chunk, RPT = ceil(F / (32/LPR)) so that ONE round covers the sample with all 32 lanes busy (F = 39, D = 16:
// 8 row blocks of 5 x 4 chunks); per j it reads RPT scalars of G (lanes of a row block share the address) and one
// 16-byte chunk of e_j for 4*RPT FMAs.  (The first version used 4-row tasks: 40 tasks = two rounds, the second with
// 8 lanes, and two LDS.128 per 16 FMAs -- 0.28 of the HBM roofline.)
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Per-warp slice: sE[2] (double-buffered sample rows) | sW (the sample's P upstream gradients, linear) | sG (F x GS).
// While sample b is in the FMA loop, the rows and gradients of the warp's NEXT sample stream in with cp.async (the first
// version loaded, then computed, with nothing in flight: `long_scoreboard` was its top stall and it ran at 0.41 of the
// roofline, profiles/r1_ncu_summary.md).
template <int LPR, int RPT>
__global__ void __launch_bounds__(kThreads) k_ip_bwd_warp(const float* __restrict__ E, const float* __restrict__ dout,
                                                         float* __restrict__ dE, int64_t B, int F) {
    constexpr int D = 4 * LPR;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int P = F * (F - 1) / 2;
    const int NB = (F + RPT - 1) / RPT;               // row blocks
    const int GS = NB * RPT + 1;                      // odd-ish stride: the transposed scatter below spreads over banks
    const int SL = slice_floats(F, D), P4 = (P + 3) & ~3, G4 = (F * GS + 3) & ~3;
    int* sPair = reinterpret_cast<int*>(smem);                        // [P] (i << 16) | j
    float* sE0 = smem + P4 + (size_t)warp * (2 * SL + P4 + G4);
    float* sE1 = sE0 + SL;
    float* sW = sE1 + SL;
    float* sG = sW + P4;
    for (int i = threadIdx.x; i < F; i += blockDim.x)
        for (int j = i + 1; j < F; ++j) sPair[pair_index(i, j, F)] = (i << 16) | j;
    zero_pad_rows<LPR>(sE0, F, lane);
    zero_pad_rows<LPR>(sE1, F, lane);
    for (int t = lane; t < F * GS; t += 32) sG[t] = 0.f;               // diagonal and padding stay zero
    __syncthreads();
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp, nw = (int64_t)gridDim.x * wpc;
    const int NTask = NB * LPR;
    auto prefetch = [&](int64_t bn, float* sEn) {
        const float* Eb = E + (size_t)bn * F * D;
        for (int t = lane; t < F * LPR; t += 32) cp_async_16(sEn + row_off(t / LPR, D) + 4 * (t & (LPR - 1)), Eb + 4 * (size_t)t);
        const float* wb = dout + (size_t)bn * P;
        for (int p = lane; p < P; p += 32) cp_async_4(sW + p, wb + p);
    };
    if (gw < B) prefetch(gw, sE0);
    cp_async_commit_group();
    int cur = 0;
    for (int64_t b = gw; b < B; b += nw, cur ^= 1) {
        cp_async_wait_all();
        __syncwarp();                                  // sample b has landed; everyone is done with sample b - nw
        const float* sE = cur ? sE1 : sE0;
        for (int p = lane; p < P; p += 32) {
            const int ij = sPair[p], i = ij >> 16, j = ij & 0xffff;
            const float w = sW[p];
            sG[j * GS + i] = w;
            sG[i * GS + j] = w;
        }
        __syncwarp();                                  // sW is free again, sG is complete
        if (b + nw < B) prefetch(b + nw, cur ? sE0 : sE1);
        cp_async_commit_group();
        float* db = dE + (size_t)b * F * D;
        for (int t = lane; t < NTask; t += 32) {
            const int ib = t / LPR, c = t & (LPR - 1);
            float4 acc[RPT];
#pragma unroll
            for (int x = 0; x < RPT; ++x) acc[x] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* gcol = sG + ib * RPT;
            const float* ecol = sE + 4 * c;
#pragma unroll 2
            for (int j = 0; j < F; ++j) {
                const float4 e = *reinterpret_cast<const float4*>(ecol + row_off(j, D));
#pragma unroll
                for (int x = 0; x < RPT; ++x) acc[x] = f4_fma_x2(e, gcol[j * GS + x], acc[x]);
            }
#pragma unroll
            for (int x = 0; x < RPT; ++x)
                if (ib * RPT + x < F) st_stream_f4(db + (size_t)(ib * RPT + x) * D + 4 * c, acc[x]);
        }
    }
    cp_async_wait_all();
}

// ------------------------------------------------------------------ mode 2 on the tensor cores (D = 16, 25 <= F <= 40)
// The per-sample work of mode 2 is a tiny matrix product -- forward: the upper triangle of the Gram matrix E E^T
// ([F,16] x [16,F]); backward: dE = G E with G the symmetric zero-diagonal [F,F] matrix of the upstream gradients -- and
// the SIMT kernels above spend their time moving operands from shared memory to the FMA pipe (389 / 530 shared-memory
// wavefronts and 1 153 / 1 832 warp instructions per sample at F = 39, profiles/r1_ncu_summary.md).  Here a warp still
// owns a sample, but the products run as warp-level mma.sync.m16n8k8 in 3xTF32 (x = hi + lo, hi = tf32(x),
// lo = tf32(x - hi); a_lo b_hi + a_hi b_lo + a_hi b_hi in fp32 accumulators: fp32-level accuracy, the 1e-5 contract) with
// every operand element read from shared memory ONCE per lane, as a conflict-free LDS.32 into its fragment register:
//   fragments (g = lane >> 2, t = lane & 3):  A 16x8: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)
//                                             B 8x8 : b0 (k = t, n = g) b1 (k = t+4, n = g)
//                                             C 16x8: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// These are per-sample 40 x 40 x 16 products, far below a tcgen05 tile (M = 128 per instruction, operands through
// descriptors): the warp-level form is the one that fits, the kernels stay HBM-bound by design.
// forward : rows padded to 40 (5 row blocks of 8).  x[rb][ks][h] = E[8 rb + g][8 ks + 4 h + t] serves as the A fragment
//           of m-tile i (row blocks 2i, 2i+1) AND as the B fragment of n-tile rb, because B = E^T.  Only the 9 (m, n) tile
//           pairs that touch the upper triangle are computed: 9 x 2 k-steps x 3 = 54 mma per sample.  C goes to a dense
//           [40][40] tile in shared memory (STS.64, conflict-free at stride 40), from where the lanes write the P outputs
//           in order with coalesced 4-byte stores; the (i, j) -> offset table of a lane's 25 outputs sits in registers.
// backward: dE^T [16 x F] = E^T [16 x 40] G [40 x F]: one m-tile, 5 n-tiles, 5 k-steps: 75 mma per sample (the
//           untransposed form needs 3 m-tiles of which the last is half empty: 90).  G is expanded from the staged
//           gradients into a dense symmetric [40][44] tile (diagonal and padding stay zero), B fragments are plain LDS.32.
// The next sample's rows (and gradients) stream in with cp.async as soon as the current sample's fragments are in
// registers, into the same buffers.
constexpr int kIpmRows = 40;          // padded fields
constexpr int kIpmES = 20;            // forward: row stride of the staged sample (bank = 20 g + t: conflict-free)
constexpr int kIpmCS = 40;            // forward: row stride of the C tile (STS.64 of (g, 2t): conflict-free)
constexpr int kIpmBS = 24;            // backward: row stride of the staged sample (bank = 24 t + g: conflict-free)
constexpr int kIpmGS = 44;            // backward: row stride of the G tile (bank = 12 g + t: conflict-free)
constexpr int kIpmMaxK = 25;          // ceil(40 * 39 / 2 / 32) outputs per lane
constexpr int kIpmTab = 400;          // forward: floats reserved for the CTA's pair table (780 x uint16)
constexpr int kIpmTabB = 800;         // backward: 780 x uint32
constexpr int kIpmFwdWarps = 4, kIpmBwdWarps = 3;
constexpr int kIpmFwdSlice = kIpmRows * kIpmES + kIpmRows * kIpmCS;           // 2 400 floats per warp
constexpr int kIpmW = 784;                                                    // staged gradients: P + 3 (alignment slack), /4
constexpr int kIpmBwdSlice = kIpmRows * kIpmBS + kIpmW + kIpmRows * kIpmGS;   // 3 504 floats per warp

// x = hi + lo for 3xTF32.  tf32 = the upper 19 bits of an fp32; the mma reads exactly those and ignores the low 13, so
// rounding (to nearest, ties away -- what cvt.rna.tf32.f32 does) is "+ 0x1000" on the bit pattern.  hi is also masked
// because x - hi must be exact; lo only carries the rounding increment (its low bits are ignored by the mma).  Four
// integer / FP instructions per element; ptxas expands the two cvt.rna of the textbook form into nine.
__device__ __forceinline__ void ipm_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) + 0x1000u;
}
__device__ __forceinline__ void ipm_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kIpmFwdWarps * 32, 5) k_ip_fwd_mma(const float* __restrict__ E, float* __restrict__ out, int64_t B, int F) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int P = F * (F - 1) / 2;
    unsigned short* sPair = reinterpret_cast<unsigned short*>(smem);                  // [P] -> i * CS + j
    float* sE = smem + kIpmTab + (size_t)warp * kIpmFwdSlice;
    float* sC = sE + kIpmRows * kIpmES;
    for (int i = threadIdx.x; i < F; i += blockDim.x)
        for (int j = i + 1; j < F; ++j) sPair[pair_index(i, j, F)] = (unsigned short)(i * kIpmCS + j);
    for (int q = F * kIpmES + lane; q < kIpmRows * kIpmES; q += 32) sE[q] = 0.f;      // rows F .. 39: zero, never refilled
    __syncthreads();
    uint32_t tab[(kIpmMaxK + 1) / 2];
#pragma unroll
    for (int k = 0; k < (kIpmMaxK + 1) / 2; ++k) {
        const int p0 = lane + 64 * k, p1 = p0 + 32;
        const uint32_t lo = p0 < P ? sPair[p0] : 0u, hi = p1 < P ? sPair[p1] : 0u;
        tab[k] = lo | (hi << 16);
    }
    const int64_t gw = (int64_t)blockIdx.x * kIpmFwdWarps + warp, nw = (int64_t)gridDim.x * kIpmFwdWarps;
    // chunk q = lane + 32 k of the sample's F * 4 sixteen-byte chunks -> row q >> 2, column 4 (q & 3): lane-constant
    // addresses plus k * 512 B (global) / k * 8 rows (shared)
    float* const pfDst = sE + (lane >> 2) * kIpmES + 4 * (lane & 3);
    const int nChunk = F * 4;
    auto prefetch = [&](int64_t bn) {
        const float* src = E + (size_t)bn * F * 16 + 4 * lane;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (lane + 32 * k < nChunk) cp_async_16(pfDst + 8 * k * kIpmES, src + 128 * k);
    };
    if (gw < B) prefetch(gw);
    cp_async_commit_group();
    const float* xE = sE + g * kIpmES + t;
    float* cW = sC + g * kIpmCS + 2 * t;
    for (int64_t b = gw; b < B; b += nw) {
        cp_async_wait_all();
        __syncwarp();                                  // sample b has landed; every lane is done with the previous C tile
        uint32_t xh[5][2][2], xl[5][2][2];
#pragma unroll
        for (int rb = 0; rb < 5; ++rb)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int h = 0; h < 2; ++h) ipm_split(xE[8 * rb * kIpmES + 8 * ks + 4 * h], xh[rb][ks][h], xl[rb][ks][h]);
        __syncwarp();                                  // the staged rows are in registers everywhere: refill the buffer
        if (b + nw < B) prefetch(b + nw);
        cp_async_commit_group();
        // tile (i, j): rows 16 i .. 16 i + 15 (row blocks 2i, 2i+1; block 5 does not exist: zeros), columns 8 j .. 8 j + 7
#define RBX_IPM_TILE(i, j)                                                                                              \
        {                                                                                                               \
            float c[4] = {0.f, 0.f, 0.f, 0.f};                                                                          \
            _Pragma("unroll") for (int ks = 0; ks < 2; ++ks) {                                                          \
                const uint32_t ah0 = xh[2 * (i)][ks][0], ah2 = xh[2 * (i)][ks][1], al0 = xl[2 * (i)][ks][0], al2 = xl[2 * (i)][ks][1]; \
                const uint32_t ah1 = (i) < 2 ? xh[(i) < 2 ? 2 * (i) + 1 : 0][ks][0] : 0u, ah3 = (i) < 2 ? xh[(i) < 2 ? 2 * (i) + 1 : 0][ks][1] : 0u; \
                const uint32_t al1 = (i) < 2 ? xl[(i) < 2 ? 2 * (i) + 1 : 0][ks][0] : 0u, al3 = (i) < 2 ? xl[(i) < 2 ? 2 * (i) + 1 : 0][ks][1] : 0u; \
                ipm_mma(c, al0, al1, al2, al3, xh[j][ks][0], xh[j][ks][1]);                                             \
                ipm_mma(c, ah0, ah1, ah2, ah3, xl[j][ks][0], xl[j][ks][1]);                                             \
                ipm_mma(c, ah0, ah1, ah2, ah3, xh[j][ks][0], xh[j][ks][1]);                                             \
            }                                                                                                           \
            *reinterpret_cast<float2*>(cW + 16 * (i) * kIpmCS + 8 * (j)) = make_float2(c[0], c[1]);                     \
            if ((i) < 2) *reinterpret_cast<float2*>(cW + (16 * (i) + 8) * kIpmCS + 8 * (j)) = make_float2(c[2], c[3]);  \
        }
        RBX_IPM_TILE(0, 0) RBX_IPM_TILE(0, 1) RBX_IPM_TILE(0, 2) RBX_IPM_TILE(0, 3) RBX_IPM_TILE(0, 4)
        RBX_IPM_TILE(1, 2) RBX_IPM_TILE(1, 3) RBX_IPM_TILE(1, 4) RBX_IPM_TILE(2, 4)
#undef RBX_IPM_TILE
        __syncwarp();                                  // the C tile is complete
        float* ob = out + (size_t)b * P;
#pragma unroll
        for (int k = 0; k < kIpmMaxK; ++k) {
            const int p = lane + 32 * k;
            const uint32_t off = (k & 1) ? (tab[k >> 1] >> 16) : (tab[k >> 1] & 0xffffu);
            if (p < P) __stcs(ob + p, sC[off]);
        }
    }
    cp_async_wait_all();
}

__global__ void __launch_bounds__(kIpmBwdWarps * 32, 5) k_ip_bwd_mma(const float* __restrict__ E, const float* __restrict__ dout,
                                                                    float* __restrict__ dE, int64_t B, int F) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int P = F * (F - 1) / 2;
    const int64_t total = B * (int64_t)P;
    uint32_t* sPair = reinterpret_cast<uint32_t*>(smem);          // [P] -> float offsets (i * GS + j) | (j * GS + i) << 16
    float* sE = smem + kIpmTabB + (size_t)warp * kIpmBwdSlice;
    float* sW = sE + kIpmRows * kIpmBS;
    float* sG = sW + kIpmW;
    for (int i = threadIdx.x; i < F; i += blockDim.x)
        for (int j = i + 1; j < F; ++j) sPair[pair_index(i, j, F)] = (uint32_t)(i * kIpmGS + j) | ((uint32_t)(j * kIpmGS + i) << 16);
    for (int q = F * kIpmBS + lane; q < kIpmRows * kIpmBS; q += 32) sE[q] = 0.f;      // rows F .. 39 (multiplied by G's zero columns)
    for (int q = lane; q < kIpmRows * kIpmGS; q += 32) sG[q] = 0.f;                   // diagonal and padding stay zero
    __syncthreads();
    const int64_t gw = (int64_t)blockIdx.x * kIpmBwdWarps + warp, nw = (int64_t)gridDim.x * kIpmBwdWarps;
    float* const pfDst = sE + (lane >> 2) * kIpmBS + 4 * (lane & 3);
    const int nChunk = F * 4;
    auto prefetch = [&](int64_t bn) {
        const float* src = E + (size_t)bn * F * 16 + 4 * lane;
#pragma unroll
        for (int k = 0; k < 5; ++k)
            if (lane + 32 * k < nChunk) cp_async_16(pfDst + 8 * k * kIpmBS, src + 128 * k);
        // the sample's P gradients start at float bn * P, 16-byte aligned only when (bn * P) % 4 == 0: copy the aligned
        // 16-byte chunks that cover them (element p lands at sW[mis + p]); a chunk reaching past the end of dout goes float by float
        const int64_t first = bn * (int64_t)P;
        const int mis = (int)(first & 3);
        const int64_t base = first - mis + 4 * lane;
        const int nq = (mis + P + 3) >> 2;
        const float* wsrc = dout + base;
        float* wdst = sW + 4 * lane;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            if (lane + 32 * k < nq) {
                if (base + 128 * k + 4 <= total) {
                    cp_async_16(wdst + 128 * k, wsrc + 128 * k);
                } else {
                    for (int e = 0; e < 4; ++e)
                        if (base + 128 * k + e < total) cp_async_4(wdst + 128 * k + e, wsrc + 128 * k + e);
                }
            }
        }
    };
    if (gw < B) prefetch(gw);
    cp_async_commit_group();
    for (int64_t b = gw; b < B; b += nw) {
        cp_async_wait_all();
        __syncwarp();                                  // sample b has landed; every lane is done with the previous G tile
        // A = E^T: a0 = E[8 ks + t][g], a1 = E[8 ks + t][g + 8], a2 = E[8 ks + t + 4][g], a3 = E[8 ks + t + 4][g + 8]
        uint32_t ah[5][4], al[5][4];
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) {
            const float* r0 = sE + (8 * ks + t) * kIpmBS + g;
            ipm_split(r0[0], ah[ks][0], al[ks][0]);
            ipm_split(r0[8], ah[ks][1], al[ks][1]);
            ipm_split(r0[4 * kIpmBS], ah[ks][2], al[ks][2]);
            ipm_split(r0[4 * kIpmBS + 8], ah[ks][3], al[ks][3]);
        }
        const float* wv = sW + (int)((b * (int64_t)P) & 3);
#pragma unroll 5
        for (int p = lane; p < P; p += 32) {
            const uint32_t oo = sPair[p];
            const float w = wv[p];
            sG[oo & 0xffffu] = w;
            sG[oo >> 16] = w;
        }
        __syncwarp();                                  // G is complete; the staging buffers are free: refill them
        if (b + nw < B) prefetch(b + nw);
        cp_async_commit_group();
        float* db = dE + (size_t)b * F * 16;
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
            // B = G (symmetric): b0 = G[8 nt + g][8 ks + t], b1 = G[8 nt + g][8 ks + t + 4]
            const float* gr = sG + (8 * nt + g) * kIpmGS + t;
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 5; ++ks) {
                uint32_t bh0, bl0, bh1, bl1;
                ipm_split(gr[8 * ks], bh0, bl0);
                ipm_split(gr[8 * ks + 4], bh1, bl1);
                ipm_mma(c, al[ks][0], al[ks][1], al[ks][2], al[ks][3], bh0, bh1);
                ipm_mma(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bl0, bl1);
                ipm_mma(c, ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bh0, bh1);
            }
            // C = dE^T: c0 = dE[8 nt + 2t][g], c1 = dE[8 nt + 2t + 1][g], c2 / c3 the same rows at column g + 8
            const int r0 = 8 * nt + 2 * t;
            if (r0 < F) {
                __stcs(db + (size_t)r0 * 16 + g, c[0]);
                __stcs(db + (size_t)r0 * 16 + g + 8, c[2]);
            }
            if (r0 + 1 < F) {
                __stcs(db + (size_t)(r0 + 1) * 16 + g, c[1]);
                __stcs(db + (size_t)(r0 + 1) * 16 + g + 8, c[3]);
            }
        }
    }
    cp_async_wait_all();
}

// mode 3 forward: out[b, p, :] = e_i * e_j.  LPR lanes per pair, 32/LPR pairs per step, 512-byte coalesced stores.
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_ew_fwd_warp(const float* __restrict__ E, float* __restrict__ out, int64_t B, int F) {
    constexpr int D = 4 * LPR, PPW = 32 / LPR;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int P = F * (F - 1) / 2;
    int* sPair = reinterpret_cast<int*>(smem);
    float* sE = smem + ((P + 3) & ~3) + (size_t)warp * slice_floats(F, D);
    for (int i = threadIdx.x; i < F; i += blockDim.x)
        for (int j = i + 1; j < F; ++j) sPair[pair_index(i, j, F)] = (i << 16) | j;
    __syncthreads();
    const int ps = lane / LPR, c = lane & (LPR - 1);
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp, nw = (int64_t)gridDim.x * wpc;
    for (int64_t b = gw; b < B; b += nw) {
        __syncwarp();
        stage_sample<LPR>(sE, E + (size_t)b * F * D, F, lane);
        __syncwarp();
        float* ob = out + (size_t)b * P * D + 4 * c;
#pragma unroll 4
        for (int p = ps; p < P; p += PPW) {
            const int ij = sPair[p];
            const float4 a = *reinterpret_cast<const float4*>(sE + row_off(ij >> 16, D) + 4 * c);
            const float4 bb = *reinterpret_cast<const float4*>(sE + row_off(ij & 0xffff, D) + 4 * c);
            st_stream_f4(ob + (size_t)p * D, f4_mul(a, bb));
        }
    }
}

// mode 3 backward: dE[b,i,:] = sum_{j != i} dout[b, p(i,j), :] * e_j.  dout is streamed once in pair order; a step is
// (row i, 32/LPR consecutive j): the lane of (j, chunk c) adds w*e_j into its register accumulator of row i and w*e_i
// into the shared accumulator of row j (one owner per (j, c) within a row i; __syncwarp between rows).
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_ew_bwd_warp(const float* __restrict__ E, const float* __restrict__ dout,
                                                         float* __restrict__ dE, int64_t B, int F) {
    constexpr int D = 4 * LPR, PPW = 32 / LPR;
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int P = F * (F - 1) / 2;
    const int per_warp = 2 * slice_floats(F, D);
    float* sE = smem + (size_t)warp * per_warp;
    float* sD = sE + slice_floats(F, D);
    const int ps = lane / LPR, c = lane & (LPR - 1);
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp, nw = (int64_t)gridDim.x * wpc;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t b = gw; b < B; b += nw) {
        __syncwarp();
        stage_sample<LPR>(sE, E + (size_t)b * F * D, F, lane);
        for (int t = lane; t < F * LPR; t += 32) *reinterpret_cast<float4*>(sD + row_off(t / LPR, D) + 4 * (t & (LPR - 1))) = zero;
        __syncwarp();
        const float* gb = dout + (size_t)b * P * D + 4 * c;
        int pbase = 0;                                   // pair_index(i, i+1)
        for (int i = 0; i + 1 < F; ++i) {
            const float4 ei = *reinterpret_cast<const float4*>(sE + row_off(i, D) + 4 * c);
            float4 acc = zero;
            const int nj = F - 1 - i;
            // loads of the whole row first (they do not depend on shared memory), then the accumulation
            for (int q0 = 0; q0 < nj; q0 += 4 * PPW) {
                float4 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + u * PPW + ps;
                    w[u] = q < nj ? ld_stream_f4(gb + (size_t)(pbase + q) * D) : zero;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int q = q0 + u * PPW + ps;
                    if (q < nj) {
                        const int j = i + 1 + q;
                        float4* dj = reinterpret_cast<float4*>(sD + row_off(j, D) + 4 * c);
                        acc = f4_fma4(w[u], *reinterpret_cast<const float4*>(sE + row_off(j, D) + 4 * c), acc);
                        *dj = f4_fma4(w[u], ei, *dj);
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
            if (ps == 0) {
                float4* di = reinterpret_cast<float4*>(sD + row_off(i, D) + 4 * c);
                *di = f4_add(*di, acc);
            }
            pbase += nj;
            __syncwarp();
        }
        float* db = dE + (size_t)b * F * D;
        for (int t = lane; t < F * LPR; t += 32)
            st_stream_f4(db + 4 * (size_t)t, *reinterpret_cast<const float4*>(sD + row_off(t / LPR, D) + 4 * (t & (LPR - 1))));
    }
}

// warps per CTA so that `fixed + wpc * per_warp` floats fit the shared-memory budget; 0 = does not fit at all
inline int warps_for(size_t fixed_floats, size_t per_warp_floats, size_t* smem_bytes) {
    const size_t budget = 200 * 1024;
    int wpc = kThreads / 32;
    while (wpc > 0 && (fixed_floats + wpc * per_warp_floats) * 4 > budget) --wpc;
    *smem_bytes = (fixed_floats + (size_t)wpc * per_warp_floats) * 4;
    return wpc;
}

template <typename K>
inline int warp_grid(K kernel, int wpc, size_t smem, int64_t B) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, wpc * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return capped_grid((B + wpc - 1) / wpc, per_sm);
}

// ------------------------------------------------------------------ f4: power sums over the fields (InteractionMachine)
// P[b,k-1,:] = sum_f e_f^k for k = 1..order (order <= 5), the only pass over [B,F,D] that InteractionMachine.forward
// (ranking/pytorch/layers/interactions/interaction_machine.py:44-70) needs: the reference walks E `order` times and
// materialises Q = Q * X each time.  Backward: dE[b,f,:] = sum_k k * e_f^(k-1) * dP[b,k-1,:]  (one read of E, one write).
constexpr int kMaxOrder = 5;

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_power_sums_fwd_vec(const float* __restrict__ E, float* __restrict__ P, int64_t B,
                                                                int F, int order) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        if (b >= B) continue;
        float4 acc[kMaxOrder];
#pragma unroll
        for (int k = 0; k < kMaxOrder; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* e = E + (size_t)b * F * D + 4 * lig;
#pragma unroll 4
        for (int f = 0; f < F; ++f) {
            const float4 v = ld_stream_f4(e + (size_t)f * D);
            float4 q = v;
#pragma unroll
            for (int k = 0; k < kMaxOrder; ++k) {
                if (k < order) {
                    acc[k] = f4_add(acc[k], q);
                    q = f4_mul(q, v);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kMaxOrder; ++k)
            if (k < order) *reinterpret_cast<float4*>(P + ((size_t)b * order + k) * D + 4 * lig) = acc[k];
    }
}

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_power_sums_bwd_vec(const float* __restrict__ E, const float* __restrict__ dP,
                                                                float* __restrict__ dE, int64_t B, int F, int order) {
    constexpr int D = 4 * LPR, SPW = 32 / LPR;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        if (b >= B) continue;
        float4 g[kMaxOrder];                              // k * dP_k
#pragma unroll
        for (int k = 0; k < kMaxOrder; ++k) {
            g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < order) g[k] = f4_scale(ld_row_f4(dP + ((size_t)b * order + k) * D + 4 * lig), (float)(k + 1));
        }
        const float* e = E + (size_t)b * F * D + 4 * lig;
        float* de = dE + (size_t)b * F * D + 4 * lig;
#pragma unroll 4
        for (int f = 0; f < F; ++f) {
            const float4 v = ld_stream_f4(e + (size_t)f * D);
            float4 r = g[0], q = v;
#pragma unroll
            for (int k = 1; k < kMaxOrder; ++k) {
                if (k < order) {
                    r = f4_fma4(g[k], q, r);
                    q = f4_mul(q, v);
                }
            }
            st_stream_f4(de + (size_t)f * D, r);
        }
    }
}

// Lane mapping of k_sumsq_reg: a sample is spread over LPR column chunks x FG field groups, every lane's loads are
// independent (the first version walked the 39 fields serially per lane: 0.45 / 0.60 of the roofline).
template <int LPR, int R>
__global__ void __launch_bounds__(kThreads) k_power_sums_reg(const float* __restrict__ E, const float* __restrict__ dP,
                                                            float* __restrict__ out, int64_t B, int F, int order, int bwd) {
    constexpr int D = 4 * LPR, FG = SumsqMap<LPR>::FG, SPW = SumsqMap<LPR>::SPW;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), fg = (lane / LPR) & (FG - 1), sg = lane / (LPR * FG);
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * SPW; base < B; base += nwarps * SPW) {
        const int64_t b = base + sg;
        const bool valid = b < B;
        const float* e = E + (size_t)(valid ? b : 0) * F * D + 4 * lig;
        float4 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int f = i * FG + fg;
            v[i] = (valid && f < F) ? ld_stream_f4(e + (size_t)f * D) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!bwd) {
            float4 acc[kMaxOrder];
#pragma unroll
            for (int k = 0; k < kMaxOrder; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < R; ++i) {
                float4 q = v[i];
#pragma unroll
                for (int k = 0; k < kMaxOrder; ++k)
                    if (k < order) {
                        acc[k] = f4_add(acc[k], q);
                        q = f4_mul(q, v[i]);
                    }
            }
#pragma unroll
            for (int k = 0; k < kMaxOrder; ++k) {
                if (k < order) {
#pragma unroll
                    for (int o = LPR; o < LPR * FG; o <<= 1) {
                        acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o); acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
                        acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o); acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
                    }
                    if (valid && fg == 0) *reinterpret_cast<float4*>(out + ((size_t)b * order + k) * D + 4 * lig) = acc[k];
                }
            }
        } else if (valid) {
            float4 g[kMaxOrder];                              // k * dP_k
#pragma unroll
            for (int k = 0; k < kMaxOrder; ++k) {
                g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < order) g[k] = f4_scale(ld_row_f4(dP + ((size_t)b * order + k) * D + 4 * lig), (float)(k + 1));
            }
            float* de = out + (size_t)b * F * D + 4 * lig;
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int f = i * FG + fg;
                if (f < F) {
                    float4 r = g[0], q = v[i];
#pragma unroll
                    for (int k = 1; k < kMaxOrder; ++k)
                        if (k < order) {
                            r = f4_fma4(g[k], q, r);
                            q = f4_mul(q, v[i]);
                        }
                    st_stream_f4(de + (size_t)f * D, r);
                }
            }
        }
    }
}

template <int LPR>
bool launch_power_sums_reg(const float* E, const float* dP, float* out, int64_t B, int F, int order, int bwd, cudaStream_t st) {
    constexpr int FG = SumsqMap<LPR>::FG, SPW = SumsqMap<LPR>::SPW;
    const int need = (F + FG - 1) / FG;
    const int grid = capped_grid((B + SPW * 8 - 1) / (SPW * 8), 8);
    if (need <= 4) k_power_sums_reg<LPR, 4><<<grid, kThreads, 0, st>>>(E, dP, out, B, F, order, bwd);
    else if (need <= 8) k_power_sums_reg<LPR, 8><<<grid, kThreads, 0, st>>>(E, dP, out, B, F, order, bwd);
    else if (need <= 12) k_power_sums_reg<LPR, 12><<<grid, kThreads, 0, st>>>(E, dP, out, B, F, order, bwd);
    else return false;
    return true;
}

// any D: warp per sample, lane owns columns d, d + 32, ...
__global__ void __launch_bounds__(kThreads) k_power_sums_any(const float* __restrict__ E, const float* __restrict__ dP,
                                                            float* __restrict__ out, int64_t B, int F, int D, int order, int bwd) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t b = warp0; b < B; b += nwarps) {
        for (int d = lane; d < D; d += 32) {
            float acc[kMaxOrder];
#pragma unroll
            for (int k = 0; k < kMaxOrder; ++k) acc[k] = (bwd && k < order) ? (k + 1) * __ldg(dP + ((size_t)b * order + k) * D + d) : 0.f;
            for (int f = 0; f < F; ++f) {
                const size_t o = ((size_t)b * F + f) * D + d;
                const float v = __ldg(E + o);
                float q = v, r = acc[0];
#pragma unroll
                for (int k = 0; k < kMaxOrder; ++k) {
                    if (k < order) {
                        if (!bwd) acc[k] += q;
                        else if (k >= 1) r = fmaf(acc[k], q, r);
                        if (!bwd || k >= 1) q *= v;
                    }
                }
                if (bwd) out[o] = r;
            }
            if (!bwd)
#pragma unroll
                for (int k = 0; k < kMaxOrder; ++k)
                    if (k < order) out[((size_t)b * order + k) * D + d] = acc[k];
        }
    }
}

// mode 2 on the warp-level tensor-core path: D = 16 and 25 <= F <= 40 (the Criteo-shaped configs; smaller F leave most of
// the padded 40 x 40 product empty and stay on the SIMT tiles).  RBX_IP_ENGINE=0 selects the SIMT kernels everywhere.
inline bool ipm_covers(int F, int D) {
    if (D != 16 || F < 25 || F > kIpmRows) return false;
    const char* e = getenv("RBX_IP_ENGINE");
    return !(e && *e == '0');
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
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && F >= 1 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode >= 0 && mode <= 3, "%s: InnerProductInteraction mode %d is not supported", who, mode);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(E && out, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (mode <= 1) {
        if (vec_ok(D, E, mode == 1 ? out : nullptr, nullptr)) {
            bool done = false;
            RBX_DISPATCH_LPR(D, (done = launch_sumsq_reg<LPR>(E, nullptr, out, B, F, mode, 0, st)));
            if (!done)
                RBX_DISPATCH_LPR(D, (k_sumsq_fwd_vec<LPR><<<capped_grid((B + 32 / LPR * 8 - 1) / (32 / LPR * 8), 8), kThreads, 0, st>>>(E, out, B, F, mode)));
        } else {
            k_sumsq_fwd_any<<<capped_grid((B + 7) / 8, 8), kThreads, 0, st>>>(E, out, B, F, D, mode);
        }
    } else {
        if (F < 2) return RBX_OK;
        RBX_REQUIRE(F < 32768, "%s: F too large", who);
        const size_t P = (size_t)F * (F - 1) / 2;
        if (mode == 2 && ipm_covers(F, D) && (uintptr_t)E % 16 == 0) {       // warp-level 3xTF32 mma (RBX_IP_ENGINE=0: SIMT tiles)
            const size_t msmem = ((size_t)kIpmTab + (size_t)kIpmFwdWarps * kIpmFwdSlice) * 4;
            k_ip_fwd_mma<<<warp_grid(k_ip_fwd_mma, kIpmFwdWarps, msmem, B), kIpmFwdWarps * 32, msmem, st>>>(E, out, B, F);
            RBX_LAUNCH_CHECK(who);
            return RBX_OK;
        }
        if (vec_ok(D, E, mode == 3 ? out : nullptr, nullptr) && F <= 255) {
            const int FB = (F + 3) / 4;
            size_t wsmem = 0;
            const size_t fixed = mode == 2 ? (size_t)((FB * (FB + 1) / 2 + 3) & ~3) : ((P + 3) & ~(size_t)3);
            const int wpc = warps_for(fixed, slice_floats(F, D), &wsmem);
            if (wpc > 0) {
                if (mode == 2) {
                    RBX_DISPATCH_LPR(D, (k_ip_fwd_warp<LPR><<<warp_grid(k_ip_fwd_warp<LPR>, wpc, wsmem, B), wpc * 32, wsmem, st>>>(E, out, B, F)));
                } else {
                    RBX_DISPATCH_LPR(D, (k_ew_fwd_warp<LPR><<<warp_grid(k_ew_fwd_warp<LPR>, wpc, wsmem, B), wpc * 32, wsmem, st>>>(E, out, B, F)));
                }
                RBX_LAUNCH_CHECK(who);
                return RBX_OK;
            }
        }
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
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && F >= 1 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(mode >= 0 && mode <= 3, "%s: InnerProductInteraction mode %d is not supported", who, mode);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(E && dE && (dout || F < 2), "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (mode <= 1) {
        if (vec_ok(D, E, dE, mode == 1 ? dout : nullptr)) {
            bool done = false;
            RBX_DISPATCH_LPR(D, (done = launch_sumsq_reg<LPR>(E, dout, dE, B, F, mode, 1, st)));
            if (!done)
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
        if (mode == 2 && ipm_covers(F, D) && (uintptr_t)E % 16 == 0 && (uintptr_t)dout % 16 == 0) {
            const size_t msmem = ((size_t)kIpmTabB + (size_t)kIpmBwdWarps * kIpmBwdSlice) * 4;
            k_ip_bwd_mma<<<warp_grid(k_ip_bwd_mma, kIpmBwdWarps, msmem, B), kIpmBwdWarps * 32, msmem, st>>>(E, dout, dE, B, F);
            RBX_LAUNCH_CHECK(who);
            return RBX_OK;
        }
        if (vec_ok(D, E, dE, mode == 3 ? dout : nullptr) && F <= 255) {
            const int FB = (F + 3) / 4;
            size_t wsmem = 0;
            (void)FB;
            // mode 2: rows per lane so that one round of 32 lanes covers the sample (capped at 8: more rounds beyond)
            int rpt = (F + (32 / (D / 4)) - 1) / (32 / (D / 4));
            rpt = rpt < 1 ? 1 : (rpt > 8 ? 8 : rpt);
            const int NBk = (F + rpt - 1) / rpt, GS = NBk * rpt + 1;
            const size_t fixed = mode == 2 ? ((P + 3) & ~(size_t)3) : 0;
            const size_t per_warp = mode == 2 ? 2 * (size_t)slice_floats(F, D) + ((P + 3) & ~(size_t)3) + (((size_t)F * GS + 3) & ~(size_t)3)
                                              : 2 * (size_t)slice_floats(F, D);
            int wpc = warps_for(fixed, per_warp, &wsmem);
            if (mode == 2 && wpc > 0) {       // most resident warps per SM (227 KB, ~1 KB reserved per CTA), not per CTA
                int best = wpc, best_warps = 0;
                for (int w = wpc; w >= 1; --w) {
                    const size_t cta = (fixed + (size_t)w * per_warp) * 4 + 1024;
                    const int warps = (int)((227 * 1024) / cta) * w;
                    if (warps > best_warps) { best_warps = warps; best = w; }
                }
                wpc = best;
                wsmem = (fixed + (size_t)wpc * per_warp) * 4;
            }
            if (wpc > 0) {
                if (mode == 2) {
#define RBX_IPB(R) RBX_DISPATCH_LPR(D, (k_ip_bwd_warp<LPR, R><<<warp_grid(k_ip_bwd_warp<LPR, R>, wpc, wsmem, B), wpc * 32, wsmem, st>>>(E, dout, dE, B, F)))
                    switch (rpt) {
                        case 1: RBX_IPB(1); break;
                        case 2: RBX_IPB(2); break;
                        case 3: RBX_IPB(3); break;
                        case 4: RBX_IPB(4); break;
                        case 5: RBX_IPB(5); break;
                        case 6: RBX_IPB(6); break;
                        case 7: RBX_IPB(7); break;
                        default: RBX_IPB(8); break;
                    }
#undef RBX_IPB
                } else {
                    RBX_DISPATCH_LPR(D, (k_ew_bwd_warp<LPR><<<warp_grid(k_ew_bwd_warp<LPR>, wpc, wsmem, B), wpc * 32, wsmem, st>>>(E, dout, dE, B, F)));
                }
                RBX_LAUNCH_CHECK(who);
                return RBX_OK;
            }
        }
        const size_t smem = ((size_t)F * (D + 1) + (mode == 2 ? P : 0)) * 4;
        RBX_REQUIRE(smem <= 200 * 1024, "%s: F=%d D=%d needs %zu B shared memory", who, F, D, smem);
        if (smem > 48 * 1024) cudaFuncSetAttribute(k_pairs_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pairs_bwd<<<capped_grid(B, 6), kThreads, smem, st>>>(E, dout, dE, B, F, D, mode);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_power_sums_fwd(const float* E, float* P, int64_t B, int F, int D, int order, rbx_stream_t stream) {
    const char* who = "rbx_power_sums_fwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && F >= 1 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(order >= 1 && order <= kMaxOrder, "%s: order=%d is not supported", who, order);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(E && P, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    bool done = false;
    if (vec_ok(D, E, P, nullptr)) RBX_DISPATCH_LPR(D, (done = launch_power_sums_reg<LPR>(E, nullptr, P, B, F, order, 0, st)));
    if (done) {
    } else if (vec_ok(D, E, P, nullptr)) {
        RBX_DISPATCH_LPR(D, (k_power_sums_fwd_vec<LPR><<<capped_grid((B + 32 / LPR * 8 - 1) / (32 / LPR * 8), 8), kThreads, 0, st>>>(E, P, B, F, order)));
    } else {
        k_power_sums_any<<<capped_grid((B + 7) / 8, 8), kThreads, 0, st>>>(E, nullptr, P, B, F, D, order, 0);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_power_sums_bwd(const float* E, const float* dP, float* dE, int64_t B, int F, int D, int order, rbx_stream_t stream) {
    const char* who = "rbx_power_sums_bwd";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && F >= 1 && D >= 1 && D <= RBX_MAX_DIM, "%s: bad size", who);
    RBX_REQUIRE(order >= 1 && order <= kMaxOrder, "%s: order=%d is not supported", who, order);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(E && dP && dE, "%s: null pointer", who);
    cudaStream_t st = rbx_cast_stream(stream);
    bool done = false;
    if (vec_ok(D, E, dP, dE)) RBX_DISPATCH_LPR(D, (done = launch_power_sums_reg<LPR>(E, dP, dE, B, F, order, 1, st)));
    if (done) {
    } else if (vec_ok(D, E, dP, dE)) {
        RBX_DISPATCH_LPR(D, (k_power_sums_bwd_vec<LPR><<<capped_grid((B + 32 / LPR * 8 - 1) / (32 / LPR * 8), 8), kThreads, 0, st>>>(E, dP, dE, B, F, order)));
    } else {
        k_power_sums_any<<<capped_grid((B + 7) / 8, 8), kThreads, 0, st>>>(E, dP, dE, B, F, D, order, 1);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
