// K1/K2/K3: fused multi-slot embedding gather + FM product_sum + LR (forward) and the gradient
// scatter-add (backward).  See include/recbox_b200.h for the contract and the reference lines
// each entry point replaces; DESIGN.md "Kernels" for the roofline of each.
//
// Thread mapping (vector path, D % 4 == 0, D <= 128): a row of D floats is D/4 float4; LPR = D/4
// consecutive lanes own one SAMPLE and walk its slots, so a warp covers 32/LPR samples and one
// warp-level LDG.128 fetches 32/LPR complete rows.  The FM sums S = sum_f e_f and Q = sum_f e_f^2
// stay in registers; sum over d is a log2(LPR)-step shuffle.  Slots are processed U at a time with
// all U row loads issued before the first use (memory-level parallelism: U x 16 B per lane in
// flight; at 16 resident warps/SM that is 64 KB/SM outstanding, enough to cover HBM latency).
// Scalar path (any D <= 512): one warp per sample, lane d owns columns d, d+32, ...
#include "rbx_common.cuh"

namespace {

struct SlotMeta {
    int16_t cat_pos[RBX_MAX_SLOTS];
    int16_t num_pos[RBX_MAX_SLOTS];
};

struct FwdParams {
    const float* table;
    const float* table_lr;
    const int32_t* rows;
    const float* dense_x;
    const float* dense_w;
    const float* dense_w_lr;
    const float* lr_bias;
    float* E;
    float* S;
    float* fm_out;
    float* lr_out;
    int64_t B;
    int64_t R;
    int F, Fn, D;
    SlotMeta meta;
};

struct BwdParams {
    const float* table;
    const int32_t* rows;
    const float* dense_x;
    const float* dense_w;
    const float* E;
    const float* S;
    const float* dE;
    const float* d_fm;
    const float* d_lr;
    float* g_table;
    float* g_table_lr;
    float* g_dense_w;
    float* g_dense_w_lr;
    float* g_lr_bias;
    int64_t B;
    int64_t R;
    int F, Fn, D;
    SlotMeta meta;
    int32_t pad_row[RBX_MAX_SLOTS];
};

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// forward, vector path
// ---------------------------------------------------------------------------------------------
template <int LPR, int U>
__global__ void __launch_bounds__(kThreads) k_embed_fm_fwd_vec(const __grid_constant__ FwdParams p) {
    constexpr int D = 4 * LPR;
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int lig = lane & (LPR - 1);
    const int gi = lane / LPR;
    const int F = p.F, Fn = p.Fn, Ft = p.F + p.Fn;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const float bias = p.lr_bias ? __ldg(p.lr_bias) : 0.f;

    for (int64_t base = warp0 * SPW; base < p.B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        const bool valid = b < p.B;
        const int32_t* rb = p.rows + b * F;
        float* Eb = p.E ? p.E + (size_t)b * Ft * D + 4 * lig : nullptr;
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;

        for (int f0 = 0; f0 < F; f0 += U) {
            int32_t r[U];
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                r[u] = -1;
                if (valid && f0 + u < F) r[u] = __ldg(rb + f0 + u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((uint32_t)r[u] < (uint64_t)p.R) v[u] = ld_row_f4(p.table + (size_t)r[u] * D + 4 * lig);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (valid && f0 + u < F) {
                    S = f4_add(S, v[u]);
                    Q = f4_sqacc(v[u], Q);
                    if (Eb) st_stream_f4(Eb + (size_t)p.meta.cat_pos[f0 + u] * D, v[u]);
                }
            }
        }
        // numeric slots: e = x * w  (nn.Linear(1, D, bias=False) on x.view(-1,1))
        for (int n = 0; n < Fn; ++n) {
            if (valid) {
                const float x = __ldg(p.dense_x + b * Fn + n);
                const float4 e = f4_scale(ld_row_f4(p.dense_w + (size_t)n * D + 4 * lig), x);
                S = f4_add(S, e);
                Q = f4_sqacc(e, Q);
                if (Eb) st_stream_f4(Eb + (size_t)p.meta.num_pos[n] * D, e);
            }
        }
        if (p.S && valid) *reinterpret_cast<float4*>(p.S + (size_t)b * D + 4 * lig) = S;

        if (p.fm_out) {
            // inner_product.py:42-48: (sum^2 - sum of squares) * 0.5 per d, then sum over d
            float fm = (S.x * S.x - Q.x) * 0.5f + (S.y * S.y - Q.y) * 0.5f + (S.z * S.z - Q.z) * 0.5f +
                       (S.w * S.w - Q.w) * 0.5f;
            fm = group_sum<LPR>(fm);
            if (valid && lig == 0) p.fm_out[b] = fm;
        }
        if (p.lr_out) {
            // logistic_regression.py:30-35: D = 1 lookup of every slot, summed; lanes of the group
            // stride over the slots so one warp instruction carries 32 useful 4-byte gathers
            float lr = 0.f;
            if (valid) {
#pragma unroll 4
                for (int f = lig; f < F; f += LPR) {
                    const int32_t r = __ldg(rb + f);
                    if ((uint32_t)r < (uint64_t)p.R) lr += __ldg(p.table_lr + r);
                }
                for (int n = lig; n < Fn; n += LPR)
                    lr = fmaf(__ldg(p.dense_x + b * Fn + n), __ldg(p.dense_w_lr + n), lr);
            }
            lr = group_sum<LPR>(lr);
            if (valid && lig == 0) p.lr_out[b] = lr + bias;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward, scalar path: warp per sample, lane owns columns lane + 32k (any D <= 32*KD)
// ---------------------------------------------------------------------------------------------
template <int KD>
__global__ void __launch_bounds__(kThreads) k_embed_fm_fwd_scalar(const __grid_constant__ FwdParams p) {
    const int lane = threadIdx.x & 31;
    const int F = p.F, Fn = p.Fn, Ft = p.F + p.Fn, D = p.D;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const float bias = p.lr_bias ? __ldg(p.lr_bias) : 0.f;

    for (int64_t b = warp0; b < p.B; b += nwarps) {
        float S[KD], Q[KD];
#pragma unroll
        for (int k = 0; k < KD; ++k) S[k] = Q[k] = 0.f;
        float lr = 0.f;
        for (int f = 0; f < F; ++f) {
            const int32_t r = __ldg(p.rows + b * F + f);
            const bool ok = (uint32_t)r < (uint64_t)p.R;
            float* Eo = p.E ? p.E + ((size_t)b * Ft + p.meta.cat_pos[f]) * D : nullptr;
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                const int d = lane + 32 * k;
                if (d < D) {
                    const float e = ok ? __ldg(p.table + (size_t)r * D + d) : 0.f;
                    S[k] += e;
                    Q[k] = fmaf(e, e, Q[k]);
                    if (Eo) Eo[d] = e;
                }
            }
            if (p.lr_out && ok && lane == (f & 31)) lr += __ldg(p.table_lr + r);
        }
        for (int n = 0; n < Fn; ++n) {
            const float x = __ldg(p.dense_x + b * Fn + n);
            float* Eo = p.E ? p.E + ((size_t)b * Ft + p.meta.num_pos[n]) * D : nullptr;
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                const int d = lane + 32 * k;
                if (d < D) {
                    const float e = x * __ldg(p.dense_w + (size_t)n * D + d);
                    S[k] += e;
                    Q[k] = fmaf(e, e, Q[k]);
                    if (Eo) Eo[d] = e;
                }
            }
            if (p.lr_out && lane == (n & 31)) lr = fmaf(x, __ldg(p.dense_w_lr + n), lr);
        }
        float fm = 0.f;
#pragma unroll
        for (int k = 0; k < KD; ++k) {
            const int d = lane + 32 * k;
            if (d < D) {
                fm += (S[k] * S[k] - Q[k]) * 0.5f;
                if (p.S) p.S[(size_t)b * D + d] = S[k];
            }
        }
        if (p.fm_out) {
            fm = group_sum<32>(fm);
            if (lane == 0) p.fm_out[b] = fm;
        }
        if (p.lr_out) {
            lr = group_sum<32>(lr);
            if (lane == 0) p.lr_out[b] = lr + bias;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward, categorical slots: g_e = dE + d_fm * (S - e)  ->  red.global.add into the grad table
// ---------------------------------------------------------------------------------------------
template <int LPR, int U>
__global__ void __launch_bounds__(kThreads) k_embed_fm_bwd_vec(const __grid_constant__ BwdParams p) {
    constexpr int D = 4 * LPR;
    constexpr int SPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int lig = lane & (LPR - 1);
    const int gi = lane / LPR;
    const int F = p.F, Ft = p.F + p.Fn;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const bool has_fm = p.d_fm != nullptr;

    for (int64_t base = warp0 * SPW; base < p.B; base += nwarps * SPW) {
        const int64_t b = base + gi;
        const bool valid = b < p.B;
        const int32_t* rb = p.rows + b * F;
        const size_t eoff = (size_t)b * Ft * D + 4 * lig;
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
        float dfm = 0.f;
        if (valid && has_fm) {
            S = ld_stream_f4(p.S + (size_t)b * D + 4 * lig);
            dfm = __ldg(p.d_fm + b);
        }
        if (p.g_table) {
            for (int f0 = 0; f0 < F; f0 += U) {
                int32_t r[U];
                float4 e[U], g[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    r[u] = -1;
                    if (valid && f0 + u < F) {
                        const int32_t t = __ldg(rb + f0 + u);
                        if (t != p.pad_row[f0 + u] && (uint32_t)t < (uint64_t)p.R) r[u] = t;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    e[u] = g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r[u] >= 0) {
                        const size_t o = eoff + (size_t)p.meta.cat_pos[f0 + u] * D;
                        if (p.dE) g[u] = ld_stream_f4(p.dE + o);
                        if (has_fm)
                            e[u] = p.E ? ld_stream_f4(p.E + o) : ld_row_f4(p.table + (size_t)r[u] * D + 4 * lig);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (r[u] >= 0) {
                        if (has_fm) g[u] = f4_fma(f4_sub(S, e[u]), dfm, g[u]);
                        red_add_f4(p.g_table + (size_t)r[u] * D + 4 * lig, g[u]);
                    }
                }
            }
        }
        if (p.g_table_lr && p.d_lr && valid) {
            const float dlr = __ldg(p.d_lr + b);
#pragma unroll 4
            for (int f = lig; f < F; f += LPR) {
                const int32_t t = __ldg(rb + f);
                if (t != p.pad_row[f] && (uint32_t)t < (uint64_t)p.R) red_add_f1(p.g_table_lr + t, dlr);
            }
        }
    }
}

// backward, scalar path (any D): warp per sample
__global__ void __launch_bounds__(kThreads) k_embed_fm_bwd_scalar(const __grid_constant__ BwdParams p) {
    const int lane = threadIdx.x & 31;
    const int F = p.F, Ft = p.F + p.Fn, D = p.D;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    const bool has_fm = p.d_fm != nullptr;
    for (int64_t b = warp0; b < p.B; b += nwarps) {
        const float dfm = has_fm ? __ldg(p.d_fm + b) : 0.f;
        const float dlr = p.d_lr ? __ldg(p.d_lr + b) : 0.f;
        for (int f = 0; f < F; ++f) {
            const int32_t r = __ldg(p.rows + b * F + f);
            if (r == p.pad_row[f] || (uint32_t)r >= (uint64_t)p.R) continue;
            const size_t o = ((size_t)b * Ft + p.meta.cat_pos[f]) * D;
            if (p.g_table) {
                for (int d = lane; d < D; d += 32) {
                    float g = p.dE ? p.dE[o + d] : 0.f;
                    if (has_fm) {
                        const float e = p.E ? p.E[o + d] : __ldg(p.table + (size_t)r * D + d);
                        g = fmaf(p.S[(size_t)b * D + d] - e, dfm, g);
                    }
                    red_add_f1(p.g_table + (size_t)r * D + d, g);
                }
            }
            if (p.g_table_lr && p.d_lr && lane == 0) red_add_f1(p.g_table_lr + r, dlr);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward, numeric slots + bias: batch reductions.  Role r of a CTA column:
//   r <  Fn*D            : g_dense_w[n,d]   += sum_b x[b,n] * (dE[b,pos_n,d] + d_fm[b]*(S[b,d] - x[b,n] w[n,d]))
//   r <  Fn*D + Fn       : g_dense_w_lr[n]  += sum_b x[b,n] * d_lr[b]
//   r == Fn*D + Fn       : g_lr_bias        += sum_b d_lr[b]
// Each thread keeps one fp32 partial over its CTA's samples, then one atomic per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_dense_w_bwd(const __grid_constant__ BwdParams p) {
    const int Fn = p.Fn, D = p.D, Ft = p.F + p.Fn;
    const int role = blockIdx.y * kThreads + threadIdx.x;
    const int n_w = Fn * D;
    if (role > n_w + Fn) return;
    float acc = 0.f;
    if (role < n_w) {
        if (!p.g_dense_w) return;
        const int n = role / D, d = role - n * D;
        const float w = __ldg(p.dense_w + role);
        const size_t po = (size_t)p.meta.num_pos[n] * D + d;
        const bool has_fm = p.d_fm != nullptr;
#pragma unroll 4
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
            const float x = __ldg(p.dense_x + b * Fn + n);
            float g = p.dE ? ld_stream_f1(p.dE + (size_t)b * Ft * D + po) : 0.f;
            if (has_fm) g = fmaf(__ldg(p.S + (size_t)b * D + d) - x * w, __ldg(p.d_fm + b), g);
            acc = fmaf(x, g, acc);
        }
        atomicAdd(p.g_dense_w + role, acc);
    } else if (role < n_w + Fn) {
        if (!p.g_dense_w_lr || !p.d_lr) return;
        const int n = role - n_w;
#pragma unroll 4
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x)
            acc = fmaf(__ldg(p.dense_x + b * Fn + n), __ldg(p.d_lr + b), acc);
        atomicAdd(p.g_dense_w_lr + n, acc);
    } else {
        if (!p.g_lr_bias || !p.d_lr) return;
#pragma unroll 4
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) acc += __ldg(p.d_lr + b);
        atomicAdd(p.g_lr_bias, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------
// occupancy-capped persistent grid: min(CTAs needed, SMs x resident CTAs/SM), cached per kernel
int grid_for(const void* kernel, int64_t warps_needed) {
    struct Slot { const void* k; int occ; };
    static thread_local Slot cache[32];
    static thread_local int n_cache = 0;
    int occ = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].k == kernel) occ = cache[i].occ;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, 0) != cudaSuccess || occ <= 0) occ = 4;
        if (n_cache < 32) cache[n_cache++] = Slot{kernel, occ};
    }
    int64_t ctas = (warps_needed + kThreads / 32 - 1) / (kThreads / 32);
    const int64_t cap = (int64_t)rbx_sm_count() * occ;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

int fill_meta(SlotMeta& m, const int32_t* cat_pos, int F, const int32_t* num_pos, int Fn, const char* who) {
    const int Ft = F + Fn;
    for (int f = 0; f < F; ++f) {
        if (cat_pos[f] < 0 || cat_pos[f] >= Ft) return rbx_fail(RBX_ERR_ARG, "%s: cat_pos[%d]=%d outside [0,%d)", who, f, cat_pos[f], Ft);
        m.cat_pos[f] = (int16_t)cat_pos[f];
    }
    for (int n = 0; n < Fn; ++n) {
        if (num_pos[n] < 0 || num_pos[n] >= Ft) return rbx_fail(RBX_ERR_ARG, "%s: num_pos[%d]=%d outside [0,%d)", who, n, num_pos[n], Ft);
        m.num_pos[n] = (int16_t)num_pos[n];
    }
    return RBX_OK;
}

template <int LPR, int U>
void launch_fwd_vec(const FwdParams& p, cudaStream_t st) {
    constexpr int SPW = 32 / LPR;
    const int grid = grid_for((const void*)k_embed_fm_fwd_vec<LPR, U>, (p.B + SPW - 1) / SPW);
    k_embed_fm_fwd_vec<LPR, U><<<grid, kThreads, 0, st>>>(p);
}
template <int KD>
void launch_fwd_scalar(const FwdParams& p, cudaStream_t st) {
    const int grid = grid_for((const void*)k_embed_fm_fwd_scalar<KD>, p.B);
    k_embed_fm_fwd_scalar<KD><<<grid, kThreads, 0, st>>>(p);
}
template <int LPR, int U>
void launch_bwd_vec(const BwdParams& p, cudaStream_t st) {
    constexpr int SPW = 32 / LPR;
    const int grid = grid_for((const void*)k_embed_fm_bwd_vec<LPR, U>, (p.B + SPW - 1) / SPW);
    k_embed_fm_bwd_vec<LPR, U><<<grid, kThreads, 0, st>>>(p);
}

}  // namespace

extern "C" {

int rbx_embed_fm_fwd(const float* table, const float* table_lr, const int32_t* rows, const int32_t* cat_pos,
                     const float* dense_x, const float* dense_w, const float* dense_w_lr, const int32_t* num_pos,
                     const float* lr_bias, float* E, float* S, float* fm_out, float* lr_out, int64_t B, int64_t R,
                     int F, int Fn, int D, rbx_stream_t stream) {
    const char* who = "rbx_embed_fm_fwd";
    RBX_REQUIRE(B >= 0 && F >= 0 && Fn >= 0 && D >= 1, "%s: negative size", who);
    RBX_REQUIRE(F <= RBX_MAX_SLOTS && Fn <= RBX_MAX_SLOTS, "%s: more than %d slots", who, RBX_MAX_SLOTS);
    RBX_REQUIRE(D <= RBX_MAX_DIM, "%s: D=%d > %d", who, D, RBX_MAX_DIM);
    RBX_REQUIRE(R >= 0 && R <= INT32_MAX, "%s: R=%lld outside int32 row ids", who, (long long)R);
    if (B == 0 || F + Fn == 0) return RBX_OK;
    RBX_REQUIRE(F == 0 || (table && rows && cat_pos), "%s: table/rows/cat_pos required when F > 0", who);
    RBX_REQUIRE(Fn == 0 || (dense_x && dense_w && num_pos), "%s: dense_x/dense_w/num_pos required when Fn > 0", who);
    RBX_REQUIRE(!lr_out || ((F == 0 || table_lr) && (Fn == 0 || dense_w_lr)), "%s: lr_out needs table_lr / dense_w_lr", who);
    FwdParams p;
    p.table = table; p.table_lr = table_lr; p.rows = rows; p.dense_x = dense_x; p.dense_w = dense_w;
    p.dense_w_lr = dense_w_lr; p.lr_bias = lr_bias; p.E = E; p.S = S; p.fm_out = fm_out; p.lr_out = lr_out;
    p.B = B; p.R = R; p.F = F; p.Fn = Fn; p.D = D;
    if (int rc = fill_meta(p.meta, cat_pos, F, num_pos, Fn, who)) return rc;
    cudaStream_t st = rbx_cast_stream(stream);
    const bool aligned = ((uintptr_t)table % 16 == 0) && ((uintptr_t)dense_w % 16 == 0) && ((uintptr_t)E % 16 == 0) &&
                         ((uintptr_t)S % 16 == 0);
    if (D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && aligned) {
        switch (D / 4) {
            case 1: launch_fwd_vec<1, 8>(p, st); break;
            case 2: launch_fwd_vec<2, 8>(p, st); break;
            case 4: launch_fwd_vec<4, 8>(p, st); break;
            case 8: launch_fwd_vec<8, 8>(p, st); break;
            case 16: launch_fwd_vec<16, 8>(p, st); break;
            default: launch_fwd_vec<32, 8>(p, st); break;
        }
    } else {
        const int kd = (D + 31) / 32;
        if (kd <= 1) launch_fwd_scalar<1>(p, st);
        else if (kd <= 2) launch_fwd_scalar<2>(p, st);
        else if (kd <= 4) launch_fwd_scalar<4>(p, st);
        else if (kd <= 8) launch_fwd_scalar<8>(p, st);
        else launch_fwd_scalar<16>(p, st);
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_embed_fm_bwd(const float* table, const int32_t* rows, const int32_t* cat_pos, const int32_t* pad_row,
                     const float* dense_x, const float* dense_w, const int32_t* num_pos, const float* E,
                     const float* S, const float* dE, const float* d_fm, const float* d_lr, float* g_table,
                     float* g_table_lr, float* g_dense_w, float* g_dense_w_lr, float* g_lr_bias, int64_t B,
                     int64_t R, int F, int Fn, int D, rbx_stream_t stream) {
    const char* who = "rbx_embed_fm_bwd";
    RBX_REQUIRE(B >= 0 && F >= 0 && Fn >= 0 && D >= 1, "%s: negative size", who);
    RBX_REQUIRE(F <= RBX_MAX_SLOTS && Fn <= RBX_MAX_SLOTS, "%s: more than %d slots", who, RBX_MAX_SLOTS);
    RBX_REQUIRE(D <= RBX_MAX_DIM, "%s: D=%d > %d", who, D, RBX_MAX_DIM);
    RBX_REQUIRE(R >= 0 && R <= INT32_MAX, "%s: R=%lld outside int32 row ids", who, (long long)R);
    if (B == 0 || F + Fn == 0) return RBX_OK;
    RBX_REQUIRE(F == 0 || (rows && cat_pos), "%s: rows/cat_pos required when F > 0", who);
    RBX_REQUIRE(Fn == 0 || (dense_x && dense_w && num_pos), "%s: dense_x/dense_w/num_pos required when Fn > 0", who);
    RBX_REQUIRE(!d_fm || S, "%s: S (saved by the forward) required with d_fm", who);
    RBX_REQUIRE(!d_fm || E || table || F == 0, "%s: E or table required with d_fm", who);
    BwdParams p;
    p.table = table; p.rows = rows; p.dense_x = dense_x; p.dense_w = dense_w; p.E = E; p.S = S; p.dE = dE;
    p.d_fm = d_fm; p.d_lr = d_lr; p.g_table = g_table; p.g_table_lr = g_table_lr; p.g_dense_w = g_dense_w;
    p.g_dense_w_lr = g_dense_w_lr; p.g_lr_bias = g_lr_bias; p.B = B; p.R = R; p.F = F; p.Fn = Fn; p.D = D;
    if (int rc = fill_meta(p.meta, cat_pos, F, num_pos, Fn, who)) return rc;
    for (int f = 0; f < F; ++f) p.pad_row[f] = pad_row ? pad_row[f] : -1;
    cudaStream_t st = rbx_cast_stream(stream);

    if (F > 0 && (g_table || (g_table_lr && d_lr)) && (dE || d_fm || d_lr)) {
        const bool aligned = ((uintptr_t)table % 16 == 0) && ((uintptr_t)E % 16 == 0) && ((uintptr_t)S % 16 == 0) &&
                             ((uintptr_t)dE % 16 == 0) && ((uintptr_t)g_table % 16 == 0);
        if (D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && aligned) {
            switch (D / 4) {
                case 1: launch_bwd_vec<1, 8>(p, st); break;
                case 2: launch_bwd_vec<2, 8>(p, st); break;
                case 4: launch_bwd_vec<4, 8>(p, st); break;
                case 8: launch_bwd_vec<8, 8>(p, st); break;
                case 16: launch_bwd_vec<16, 4>(p, st); break;
                default: launch_bwd_vec<32, 4>(p, st); break;
            }
        } else {
            const int grid = grid_for((const void*)k_embed_fm_bwd_scalar, B);
            k_embed_fm_bwd_scalar<<<grid, kThreads, 0, st>>>(p);
        }
        RBX_LAUNCH_CHECK(who);
    }
    const bool want_dense = (Fn > 0 && ((g_dense_w && (dE || d_fm)) || (g_dense_w_lr && d_lr))) || (g_lr_bias && d_lr);
    if (want_dense) {
        const int roles = Fn * D + Fn + 1;
        int gx = rbx_sm_count() * 4;
        if (gx > B) gx = (int)B;
        dim3 grid(gx, (roles + kThreads - 1) / kThreads);
        k_dense_w_bwd<<<grid, kThreads, 0, st>>>(p);
        RBX_LAUNCH_CHECK(who);
    }
    return RBX_OK;
}

}  // extern "C"
