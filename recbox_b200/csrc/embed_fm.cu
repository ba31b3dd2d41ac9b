// K1/K2/K3: fused multi-slot embedding gather + FM product_sum + LR (forward) and the gradient
// scatter-add (backward).  See include/recbox_b200.h for the contract and the reference lines
// each entry point replaces; DESIGN.md "Kernels" for the roofline of each.
//
// Thread mapping (vector path, D % 4 == 0, D <= 128): a row of D floats is D/4 float4; LPR = D/4
// consecutive lanes own one SAMPLE and walk its slots, so a warp covers SPW = 32/LPR samples and
// one warp-level LDG.128 fetches SPW complete rows.  The FM sums S = sum_f e_f, Q = sum_f e_f^2 stay
// in registers; the sum over d is a log2(LPR)-step shuffle.
//
// Work distribution: persistent grid of (SMs x resident CTAs/SM) CTAs; warp w of the grid walks the
// warp groups (SPW consecutive samples) w, w + #warps, ... -- no CTA-wide barrier anywhere.
//
// Index staging (kStaged): each warp owns a private, double-buffered slot of shared memory.  While
// it works on group g, lane 0 has already launched ONE cp.async.bulk (TMA, SASS UBLKCP) for the id
// span of group g + #warps (SPW*F*4 contiguous bytes of `rows`) and one for its dense_x span,
// completing on the warp's own mbarrier.  Row addresses then come from shared memory (~30 cycles)
// instead of a dependent global load (~1 us under load), so a warp's row gathers issue back to
// back and the DRAM pipe stays full.  Spans that do not start on a 16-byte boundary are copied
// from the boundary below (skew of <= 3 words).  The direct variant (ids read with LDG) remains for
// unaligned base pointers.
//
// Scalar path (any D <= 512): one warp per sample, lane d owns columns d, d+32, ...
#include "rbx_common.cuh"

namespace {

struct SlotMeta {
    int16_t cat_pos[RBX_MAX_SLOTS];    // output position of categorical slot f
    int16_t num_pos[RBX_MAX_SLOTS];    // output position of numeric slot n
    int16_t num_widx[RBX_MAX_SLOTS];   // row of numeric slot n in dense_w / dense_w_lr
    int32_t lr_delta[RBX_MAX_SLOTS];   // row of slot f in table_lr = rows[b,f] + lr_delta[f]
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
    int F, Fn, D, Ft;
    // row-sharded table over NVLink peer memory (rbx_embed_fm_fwd_sharded): global row r lives on
    // shard r & (2^wlog2 - 1) at local row r >> wlog2; shard[w] / shard_lr[w] are device pointers
    // valid on THIS device (the local allocation, or a peer's mapped through CUDA IPC)
    int wlog2;
    const float* shard[RBX_MAX_WORLD];
    const float* shard_lr[RBX_MAX_WORLD];
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
    int F, Fn, D, Ft;
    int wlog2;
    int self_shard;                          // shard owned by this process (rbx_shard_set_rank), -1 = unknown
    const float* shard[RBX_MAX_WORLD];       // tables (re-gather of e when E is not given)
    float* g_shard[RBX_MAX_WORLD];           // gradient tables of every shard (local + peer-mapped)
    float* g_shard_lr[RBX_MAX_WORLD];
    SlotMeta meta;
    int32_t pad_row[RBX_MAX_SLOTS];
    // numeric-slot / bias batch reductions folded into the vector kernel: the CTA accumulates
    // g_dense_w [Fn,D] | g_dense_w_lr [Fn] | g_lr_bias [1] in the first num_smem bytes of its dynamic
    // shared memory and flushes them with one red per element at the end (0 = separate kernels)
    int num_smem;
};

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// tuning knobs (compile-time; profiles/ records the sweep that chose the defaults)
#ifndef RBX_FWD_MINB
#define RBX_FWD_MINB 3      // resident CTAs/SM the forward kernel is register-budgeted for
#endif
#ifndef RBX_BWD_MINB
#define RBX_BWD_MINB 4
#endif
#ifndef RBX_FWD_U
#define RBX_FWD_U 8         // row gathers in flight per lane, forward
#endif
#ifndef RBX_BWD_U
#define RBX_BWD_U 4         // slots per chunk, backward (2 streaming loads + 1 red each)
#endif
#ifndef RBX_BWD_WARP_AGG
#define RBX_BWD_WARP_AGG 1  // aggregate same-row reductions across the samples of a warp before the atomic
#endif
#ifndef RBX_BWD_FUSE_NUM
#define RBX_BWD_FUSE_NUM 1  // numeric-slot / bias batch reductions inside the vector backward kernel (0: k_dense_w_bwd*)
#endif
#ifndef RBX_BWD_REVERSE
#define RBX_BWD_REVERSE 0   // 1: the backward walks the batch back to front (L2 reuse of the forward's tail)
#endif
#ifndef RBX_L2_HINTS
#define RBX_L2_HINTS 3      // (sweeps: profiles/r1_variant_sweeps.md) bit 0: table row loads evict_last; bit 1: E/dE streams evict_first; bit 2: grad reds evict_last
#endif
#if RBX_L2_HINTS & 1
#define LD_ROW(p) ld_row_f4_hint(p, pol_keep)
#else
#define LD_ROW(p) ld_row_f4(p)
#endif
#if RBX_L2_HINTS & 4
#define RED_ROW(p, v) red_add_f4_hint(p, v, pol_keep)
#else
#define RED_ROW(p, v) red_add_f4(p, v)
#endif
#ifndef RBX_BWD_BULK
#define RBX_BWD_BULK 0      // gradient rows leave through ONE TMA bulk reduction per row (cp.reduce.async.bulk ... add.f32 from a
                            // per-warp shared-memory staging slot) instead of D/4 red.global.add.v4.f32 -- one NVLink packet per
                            // row.  1: every row of every staged kernel; 2: only rows owned by a PEER in the sharded kernels
                            // (rbx_shard_set_rank tells the library which shard is local); 0: never
#endif
// reductions into a peer's gradient table travel over NVLink: vector (v4) or four scalar reds
#ifndef RBX_PEER_RED_V4
#define RBX_PEER_RED_V4 1
#endif
#define RED_GRAD(p, v)                                                   \
    do {                                                                 \
        if (!kSharded || RBX_PEER_RED_V4) { RED_ROW(p, v); }             \
        else { red_add_f1((p), (v).x); red_add_f1((p) + 1, (v).y); red_add_f1((p) + 2, (v).z); red_add_f1((p) + 3, (v).w); } \
    } while (0)
#if RBX_L2_HINTS & 2
#define LD_STREAM(p) ld_stream_f4_hint(p, pol_stream)
#define ST_STREAM(p, v) st_stream_f4_hint(p, v, pol_stream)
#else
#define LD_STREAM(p) ld_stream_f4(p)
#define ST_STREAM(p, v) st_stream_f4(p, v)
#endif

// ---------------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// TMA bulk reduction shared -> global (element-wise fp32 add of `bytes` contiguous bytes, multiple of 16, both 16-B aligned)
__device__ __forceinline__ void bulk_red_add_f32(float* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Stage the `n` 4-byte words starting at word offset `wo` of `base` (base 16-B aligned) into the
// shared buffer `dst` (16-B aligned): the copy starts at the 16-B boundary below (skew = wo & 3
// words, so word i lands at dst[skew + i]); the 16-B multiple goes through one TMA bulk copy that
// completes on `bar`, the <= 3 trailing words through plain stores (so nothing past `wo + n` is
// ever read).  Called by one lane, which arms `bar` with the returned transaction bytes first.
__device__ __forceinline__ uint32_t stage_tx_bytes(int64_t wo, int64_t n) {
    return (uint32_t)((((wo & 3) + n) * 4) & ~(int64_t)15);
}
__device__ __forceinline__ void stage_words(void* dst, const void* base, int64_t wo, int64_t n, uint64_t* bar) {
    const int64_t skew = wo & 3;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(base) + (wo - skew);
    const uint32_t bulk = stage_tx_bytes(wo, n);
    if (bulk) tma_load_1d(dst, src, bulk, bar);
    for (int64_t w = bulk / 4; w < skew + n; ++w) reinterpret_cast<uint32_t*>(dst)[w] = __ldg(src + w);
}

// ---------------------------------------------------------------------------------------------
// row addressing: one local fused table, or 2^wlog2 shards reached through NVLink peer mappings
// ---------------------------------------------------------------------------------------------
// kSharded: 0 = one local table; 1 = row shards, first-order weights in their own shard vectors;
// 2 = row shards in the ROW+LR layout: a physical row is RS = 2 D floats, [e_0 .. e_{D-1} | w_lr | 0 ...],
// so the embedding row and its first-order weight travel in ONE NVLink request (remote
// traffic of 64-byte rows is bound by request count, not bytes)
template <int kSharded, int RS, typename P>
__device__ __forceinline__ const float* row_src(const P& p, int32_t r) {
    if constexpr (kSharded == 0) return p.table + (size_t)r * RS;
    else return p.shard[r & ((1 << p.wlog2) - 1)] + (size_t)(r >> p.wlog2) * RS;
}
template <int kSharded>
__device__ __forceinline__ const float* lr_src(const FwdParams& p, int32_t r, int f) {
    if constexpr (kSharded == 0) return p.table_lr + (r + p.meta.lr_delta[f]);
    else return p.shard_lr[r & ((1 << p.wlog2) - 1)] + (r >> p.wlog2);
}
template <int kSharded, int RS>
__device__ __forceinline__ float* grad_dst(const BwdParams& p, int32_t r) {
    if constexpr (kSharded == 0) return p.g_table + (size_t)r * RS;
    else return p.g_shard[r & ((1 << p.wlog2) - 1)] + (size_t)(r >> p.wlog2) * RS;
}
template <int kSharded>
__device__ __forceinline__ float* grad_lr_dst(const BwdParams& p, int32_t r, int f) {
    if constexpr (kSharded == 0) return p.g_table_lr + (r + p.meta.lr_delta[f]);
    else return p.g_shard_lr[r & ((1 << p.wlog2) - 1)] + (r >> p.wlog2);
}

// ---------------------------------------------------------------------------------------------
// forward, vector path
// ---------------------------------------------------------------------------------------------
template <int LPR, int U, bool kStaged, int kSharded>
__global__ void __launch_bounds__(kThreads, RBX_FWD_MINB) k_embed_fm_fwd(const __grid_constant__ FwdParams p) {
    constexpr bool kRowLr = kSharded == 2;        // LPR lanes span a physical row of RS floats; the first LPE carry e
    constexpr int RS = 4 * LPR;
    constexpr int D = kRowLr ? RS / 2 : RS;
    constexpr int LPE = D / 4;
    constexpr int SPW = 32 / LPR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lig = lane & (LPR - 1), gi = lane / LPR;
    const bool elane = !kRowLr || lig < LPE;       // this lane owns 4 embedding columns
    const int F = p.F, Fn = p.Fn, Ft = p.Ft;
    const float bias = p.lr_bias ? __ldg(p.lr_bias) : 0.f;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    (void)pol_keep; (void)pol_stream;

    // per-warp staging slots: idx[2][wi] per warp | dx[2][wx] per warp | 2 mbarriers per warp
    const int wi = (SPW * F + 6) & ~3, wx = Fn ? ((SPW * Fn + 6) & ~3) : 0;
    int32_t* my_idx = reinterpret_cast<int32_t*>(smem_raw) + (size_t)warp * 2 * wi;
    float* my_dx = reinterpret_cast<float*>(smem_raw) + (size_t)kWarps * 2 * wi + (size_t)warp * 2 * wx;
    uint64_t* my_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarps * 2 * (wi + wx) * 4) + warp * 2;

    const int64_t G = (p.B + SPW - 1) / SPW;
    const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
    auto issue = [&](int64_t g, int buf) {   // lane 0: arm the barrier, launch the bulk copies of group g
        const int64_t b0 = g * SPW;
        const int64_t ns = (p.B - b0 < SPW) ? p.B - b0 : SPW;
        mbar_expect_tx(&my_bar[buf], (F ? stage_tx_bytes(b0 * F, ns * F) : 0u) + (Fn ? stage_tx_bytes(b0 * Fn, ns * Fn) : 0u));
        if (F) stage_words(my_idx + (size_t)buf * wi, p.rows, b0 * F, ns * F, &my_bar[buf]);
        if (Fn) stage_words(my_dx + (size_t)buf * wx, p.dense_x, b0 * Fn, ns * Fn, &my_bar[buf]);
    };
    if (kStaged) {
        if (lane == 0) {
            mbar_init(&my_bar[0], 1);
            mbar_init(&my_bar[1], 1);
            fence_mbar_init();
            if (gw < G) issue(gw, 0);
        }
        __syncwarp();
    }

    int it = 0;
    for (int64_t g = gw; g < G; g += nw, ++it) {
        const int buf = it & 1;
        const int64_t b0 = g * SPW;
        if (kStaged) {
            if (lane == 0 && g + nw < G) issue(g + nw, buf ^ 1);   // slot buf^1 was released by the __syncwarp below
            mbar_wait(&my_bar[buf], (it >> 1) & 1);
            __syncwarp();                                           // trailing words written by lane 0
        }
        const int64_t b = b0 + gi;
        const bool valid = b < p.B;
        const int32_t* rb = kStaged ? my_idx + (size_t)buf * wi + ((b0 * F) & 3) + gi * F : p.rows + b * F;
        const float* xb = kStaged ? my_dx + (size_t)buf * wx + ((b0 * Fn) & 3) + gi * Fn : p.dense_x + b * Fn;
        float* Eb = (p.E && elane) ? p.E + (size_t)b * Ft * D + 4 * lig : nullptr;
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;

        // first-order gathers first: they are independent of everything below and overlap it
        float lr = 0.f;
        if (!kRowLr && p.lr_out && valid) {
#pragma unroll 4
            for (int f = lig; f < F; f += LPR) {
                const int32_t r = kStaged ? rb[f] : __ldg(rb + f);
                if ((uint32_t)r < (uint64_t)p.R) lr += __ldg(lr_src<kSharded>(p, r, f));
            }
        }
        for (int f0 = 0; (kSharded || p.table) && f0 < F; f0 += U) {
            int32_t r[U];
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                r[u] = -1;
                if (valid && f0 + u < F) r[u] = kStaged ? rb[f0 + u] : __ldg(rb + f0 + u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if ((uint32_t)r[u] < (uint64_t)p.R && (!kRowLr || lig <= LPE))
                    v[u] = LD_ROW((row_src<kSharded, RS>(p, r[u]) + 4 * lig));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (valid && f0 + u < F) {
                    if (elane) {
                        S = f4_add(S, v[u]);
                        Q = f4_sqacc(v[u], Q);
                        if (Eb) ST_STREAM(Eb + (size_t)p.meta.cat_pos[f0 + u] * D, v[u]);
                    } else if (lig == LPE) {
                        lr += v[u].x;             // the first-order weight rode along in column D of the row
                    }
                }
            }
        }
        // numeric slots: e = x * w  (nn.Linear(1, D, bias=False) on x.view(-1,1))
        for (int n = 0; p.dense_w && n < Fn; ++n) {
            if (valid && elane) {
                const float x = kStaged ? xb[n] : __ldg(xb + n);
                const float4 e = f4_scale(ld_row_f4(p.dense_w + (size_t)p.meta.num_widx[n] * D + 4 * lig), x);
                S = f4_add(S, e);
                Q = f4_sqacc(e, Q);
                if (Eb) ST_STREAM(Eb + (size_t)p.meta.num_pos[n] * D, e);
            }
        }
        if (p.S && valid && elane) *reinterpret_cast<float4*>(p.S + (size_t)b * D + 4 * lig) = S;

        if (p.fm_out) {
            // inner_product.py:42-48: (sum^2 - sum of squares) * 0.5 per d, then sum over d
            float fm = (S.x * S.x - Q.x) * 0.5f + (S.y * S.y - Q.y) * 0.5f + (S.z * S.z - Q.z) * 0.5f +
                       (S.w * S.w - Q.w) * 0.5f;
            fm = group_sum<LPR>(fm);
            if (valid && lig == 0) p.fm_out[b] = fm;
        }
        if (p.lr_out) {
            // logistic_regression.py:30-35: D = 1 lookup of every slot, summed, + bias
            if (valid)
                for (int n = lig; n < Fn; n += LPR) {
                    const float x = kStaged ? xb[n] : __ldg(xb + n);
                    lr = fmaf(x, __ldg(p.dense_w_lr + p.meta.num_widx[n]), lr);
                }
            lr = group_sum<LPR>(lr);
            if (valid && lig == 0) p.lr_out[b] = lr + bias;
        }
        if (kStaged) __syncwarp();   // every lane is done with slot `buf` before it is refilled
    }
}

// ---------------------------------------------------------------------------------------------
// forward, scalar path: warp per sample, lane owns columns lane + 32k (any D <= 32*KD)
// ---------------------------------------------------------------------------------------------
template <int KD>
__global__ void __launch_bounds__(kThreads) k_embed_fm_fwd_scalar(const __grid_constant__ FwdParams p) {
    const int lane = threadIdx.x & 31;
    const int F = p.F, Fn = p.Fn, Ft = p.Ft, D = p.D;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
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
                    const float e = (ok && p.table) ? __ldg(p.table + (size_t)r * D + d) : 0.f;
                    S[k] += e;
                    Q[k] = fmaf(e, e, Q[k]);
                    if (Eo) Eo[d] = e;
                }
            }
            if (p.lr_out && ok && lane == (f & 31)) lr += __ldg(p.table_lr + (r + p.meta.lr_delta[f]));
        }
        for (int n = 0; n < Fn; ++n) {
            const float x = __ldg(p.dense_x + b * Fn + n);
            float* Eo = p.E ? p.E + ((size_t)b * Ft + p.meta.num_pos[n]) * D : nullptr;
#pragma unroll
            for (int k = 0; k < KD; ++k) {
                const int d = lane + 32 * k;
                if (d < D) {
                    const float e = p.dense_w ? x * __ldg(p.dense_w + (size_t)p.meta.num_widx[n] * D + d) : 0.f;
                    S[k] += e;
                    Q[k] = fmaf(e, e, Q[k]);
                    if (Eo) Eo[d] = e;
                }
            }
            if (p.lr_out && lane == (n & 31)) lr = fmaf(x, __ldg(p.dense_w_lr + p.meta.num_widx[n]), lr);
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
// backward, categorical slots: g_e = dE + d_fm * (S - e)  ->  red.global.add.v4 into the grad table
// The upstream-gradient and saved-activation loads do not depend on the ids (only the reduction
// address does), so they are issued for the whole chunk before anything is consumed.
// ---------------------------------------------------------------------------------------------
template <int LPR, int U, bool kStaged, int kSharded>
__global__ void __launch_bounds__(kThreads, RBX_BWD_MINB) k_embed_fm_bwd(const __grid_constant__ BwdParams p) {
    constexpr bool kRowLr = kSharded == 2;
    constexpr int RS = 4 * LPR;
    constexpr int D = kRowLr ? RS / 2 : RS;
    constexpr int LPE = D / 4;
    constexpr int SPW = 32 / LPR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lig = lane & (LPR - 1), gi = lane / LPR;
    const bool elane = !kRowLr || lig < LPE;
    const int F = p.F, Ft = p.Ft;
    const bool has_fm = p.d_fm != nullptr;
    constexpr bool kAgg = RBX_BWD_WARP_AGG && SPW > 1;
    constexpr unsigned kGroupMask = LPR >= 32 ? 0xffffffffu : ((1u << LPR) - 1u);
    (void)kGroupMask;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    (void)pol_keep; (void)pol_stream;

    const int wi = (SPW * F + 6) & ~3;
    float* s_gw = reinterpret_cast<float*>(smem_raw);                       // [Fn*D] | [Fn] | [1], see BwdParams::num_smem
    unsigned char* stage_raw = smem_raw + p.num_smem;
    int32_t* my_idx = reinterpret_cast<int32_t*>(stage_raw) + (size_t)warp * 2 * wi;
    uint64_t* my_bar = reinterpret_cast<uint64_t*>(stage_raw + (size_t)kWarps * 2 * wi * 4) + warp * 2;
    const bool fuse_num = p.num_smem > 0;
    if (fuse_num) {
        for (int i = threadIdx.x; i < p.Fn * D + p.Fn + 1; i += kThreads) s_gw[i] = 0.f;
        __syncthreads();
    }
#if RBX_BWD_BULK
    // per warp: 2 x (U*SPW rows of RS floats) staging + 2 x (U*SPW) destination pointers, behind the id staging area
    constexpr int kBulkRows = U * SPW;
    constexpr bool kBulk = kStaged && kBulkRows <= 32 && (RBX_BWD_BULK == 1 || kSharded != 0);
    unsigned char* bulk_raw = stage_raw + (size_t)kWarps * 2 * wi * 4 + (size_t)kWarps * 2 * 8;
    float* my_gstage = reinterpret_cast<float*>(bulk_raw) + (size_t)warp * 2 * kBulkRows * RS;
    unsigned long long* my_gdst = reinterpret_cast<unsigned long long*>(bulk_raw + (size_t)kWarps * 2 * kBulkRows * RS * 4) + warp * 2 * kBulkRows;
    int chunk_no = 0;
#else
    constexpr bool kBulk = false;
#endif
    const int64_t G = (p.B + SPW - 1) / SPW;
    const int64_t gw = (int64_t)blockIdx.x * kWarps + warp, nw = (int64_t)gridDim.x * kWarps;
    auto issue = [&](int64_t g, int buf) {
        const int64_t b0 = g * SPW;
        const int64_t ns = (p.B - b0 < SPW) ? p.B - b0 : SPW;
        mbar_expect_tx(&my_bar[buf], stage_tx_bytes(b0 * F, ns * F));
        stage_words(my_idx + (size_t)buf * wi, p.rows, b0 * F, ns * F, &my_bar[buf]);
    };
    // RBX_BWD_REVERSE: walk the batch from its END.  The forward wrote E front to back, so the tail of
    // E (and of the ids) is what is still resident in L2 when the backward starts.
    auto group_of = [&](int64_t t) { return RBX_BWD_REVERSE ? G - 1 - t : t; };
    if (kStaged) {
        if (lane == 0) {
            mbar_init(&my_bar[0], 1);
            mbar_init(&my_bar[1], 1);
            fence_mbar_init();
            if (gw < G) issue(group_of(gw), 0);
        }
        __syncwarp();
    }

    int it = 0;
    for (int64_t t = gw; t < G; t += nw, ++it) {
        const int64_t g = group_of(t);
        const int buf = it & 1;
        const int64_t b0 = g * SPW;
        if (kStaged) {
            if (lane == 0 && t + nw < G) issue(group_of(t + nw), buf ^ 1);
            mbar_wait(&my_bar[buf], (it >> 1) & 1);
            __syncwarp();
        }
        const int64_t b = b0 + gi;
        const bool valid = b < p.B;
        const int32_t* rb = kStaged ? my_idx + (size_t)buf * wi + ((b0 * F) & 3) + gi * F : p.rows + b * F;
        const size_t eoff = (size_t)b * Ft * D + 4 * lig;
        float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
        float dfm = 0.f;
        if (valid && has_fm && elane) {
            S = ld_stream_f4(p.S + (size_t)b * D + 4 * lig);
            dfm = __ldg(p.d_fm + b);
        }
        float dlr_row = 0.f;                          // ROW+LR layout: the first-order gradient rides in column D
        if (kRowLr && p.d_lr && valid && lig == LPE) dlr_row = __ldg(p.d_lr + b);
        if (!kRowLr && (kSharded ? p.g_shard_lr[0] != nullptr : p.g_table_lr != nullptr) && p.d_lr) {
            const float dlr = valid ? __ldg(p.d_lr + b) : 0.f;
#pragma unroll 2
            for (int f0 = 0; f0 < F; f0 += LPR) {            // uniform trip count: the warp votes inside
                const int f = f0 + lig;
                int32_t r = -1, key = -1;
                if (valid && f < F) {
                    const int32_t tr = kStaged ? rb[f] : __ldg(rb + f);
                    if (tr != p.pad_row[f] && (uint32_t)tr < (uint64_t)p.R) {
                        r = tr;
                        key = kSharded ? tr : tr + p.meta.lr_delta[f];
                    }
                }
                float v = dlr;
                bool leader = true;
                if (kAgg) {
                    // warp-aggregated atomics: lanes (= samples) hitting the same first-order row send ONE red
                    const unsigned peers = __match_any_sync(0xffffffffu, key);
                    if (__any_sync(0xffffffffu, key >= 0 && peers != (1u << lane))) {
                        float acc = 0.f;
#pragma unroll
                        for (int sg = 0; sg < SPW; ++sg) {
                            const float d = __shfl_sync(0xffffffffu, dlr, sg * LPR);
                            acc = fmaf(d, (float)__popc(peers & (kGroupMask << (sg * LPR))), acc);
                        }
                        v = acc;
                        leader = (__ffs(peers) - 1) == lane;
                    }
                }
                if (r >= 0 && leader) red_add_f1(grad_lr_dst<kSharded>(p, r, f), v);
            }
        }
        if (kSharded ? p.g_shard[0] != nullptr : p.g_table != nullptr) {
            const bool from_table = has_fm && !p.E;
            for (int f0 = 0; f0 < F; f0 += U) {
                int32_t r[U];
                float4 e[U], gr[U];
#if RBX_BWD_BULK
                const int sb = chunk_no & 1;
                if (kBulk) {
                    ++chunk_no;
                    bulk_wait_read<1>();          // this lane's bulk ops of two chunks ago have read their staging rows
                    __syncwarp();
                }
#endif
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    r[u] = -1;
                    e[u] = gr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid && f0 + u < F) {
                        const size_t o = eoff + (size_t)p.meta.cat_pos[f0 + u] * D;
                        if (p.dE && elane) gr[u] = LD_STREAM(p.dE + o);
                        if (has_fm && p.E && elane) e[u] = LD_STREAM(p.E + o);
                        if (kRowLr && lig == LPE) gr[u].x = dlr_row;
                        const int32_t tr = kStaged ? rb[f0 + u] : __ldg(rb + f0 + u);
                        if (tr != p.pad_row[f0 + u] && (uint32_t)tr < (uint64_t)p.R) r[u] = tr;
                    }
                }
                if (from_table) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (r[u] >= 0 && elane) e[u] = LD_ROW((row_src<kSharded, RS>(p, r[u]) + 4 * lig));
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (r[u] >= 0 && has_fm && elane) gr[u] = f4_fma(f4_sub(S, e[u]), dfm, gr[u]);
                    bool leader = true;
                    if (kAgg) {
                        // warp-aggregated atomics: the SPW samples of this warp that hit the same row in this slot
                        // (hot ids under Zipf) are summed with shuffles and reduced into memory once
                        const unsigned peers = __match_any_sync(0xffffffffu, r[u]);
                        if (__any_sync(0xffffffffu, r[u] >= 0 && peers != (kGroupMask << (gi * LPR)))) {
                            float4 acc = gr[u];
#pragma unroll
                            for (int sg = 0; sg < SPW; ++sg) {
                                float4 v;
                                v.x = __shfl_sync(0xffffffffu, gr[u].x, sg * LPR + lig);
                                v.y = __shfl_sync(0xffffffffu, gr[u].y, sg * LPR + lig);
                                v.z = __shfl_sync(0xffffffffu, gr[u].z, sg * LPR + lig);
                                v.w = __shfl_sync(0xffffffffu, gr[u].w, sg * LPR + lig);
                                if (sg != gi && ((peers >> (sg * LPR)) & 1u)) acc = f4_add(acc, v);
                            }
                            gr[u] = acc;
                            leader = (__ffs(peers) - 1) / LPR == gi;
                        }
                    }
#if RBX_BWD_BULK
                    if (kBulk) {
                        const int slot = sb * kBulkRows + u * SPW + gi;
                        // mode 2: rows this rank owns go the direct way (into local HBM red.v4 is faster, profiles/r1_variant_sweeps.md)
                        const bool remote = RBX_BWD_BULK == 1 || !kSharded || (r[u] & ((1 << p.wlog2) - 1)) != p.self_shard;
                        if (remote) {
                            if (!kRowLr || lig <= LPE) *reinterpret_cast<float4*>(my_gstage + (size_t)slot * RS + 4 * lig) = gr[u];
                        } else if (r[u] >= 0 && leader && (!kRowLr || lig <= LPE)) {
                            RED_GRAD((grad_dst<kSharded, RS>(p, r[u]) + 4 * lig), gr[u]);
                        }
                        if (lig == 0)
                            my_gdst[slot] = (remote && r[u] >= 0 && leader) ? (unsigned long long)(uintptr_t)grad_dst<kSharded, RS>(p, r[u]) : 0ull;
                        continue;
                    }
#endif
                    if (r[u] >= 0 && leader && (!kRowLr || lig <= LPE)) RED_GRAD((grad_dst<kSharded, RS>(p, r[u]) + 4 * lig), gr[u]);
                }
#if RBX_BWD_BULK
                if (kBulk) {
                    fence_proxy_async_smem();     // staging rows written through the generic proxy -> visible to the TMA engine
                    __syncwarp();
                    if (lane < kBulkRows) {
                        const unsigned long long d = my_gdst[sb * kBulkRows + lane];
                        if (d) bulk_red_add_f32(reinterpret_cast<float*>((uintptr_t)d), my_gstage + (size_t)(sb * kBulkRows + lane) * RS,
                                                kRowLr ? (D + 4) * 4 : D * 4);
                    }
                    bulk_commit();
                }
#endif
            }
        }
        if (fuse_num) {
            // numeric slots (nn.Linear(1, D) on x.view(-1,1)) and the first-order / bias terms are batch
            // reductions: sum over the warp's samples with shuffles, then one shared-memory add per element
            const int Fn = p.Fn;
            const float* xb = p.dense_x + b * Fn;
            if (p.g_dense_w && (p.dE || has_fm)) {
                for (int n0 = 0; n0 < Fn; n0 += U) {
                    float x[U];
                    float4 g[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        x[u] = 0.f;
                        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid && elane && n0 + u < Fn) {
                            x[u] = __ldg(xb + n0 + u);
                            if (p.dE) g[u] = LD_STREAM(p.dE + eoff + (size_t)p.meta.num_pos[n0 + u] * D);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (n0 + u < Fn) {
                            const int n = n0 + u;
                            if (has_fm && elane) {
                                const float4 w = ld_row_f4(p.dense_w + (size_t)p.meta.num_widx[n] * D + 4 * lig);
                                g[u] = f4_fma(f4_sub(S, f4_scale(w, x[u])), dfm, g[u]);
                            }
                            float4 c = f4_scale(g[u], x[u]);
#pragma unroll
                            for (int o = LPR; o < 32; o <<= 1) {
                                c.x += __shfl_xor_sync(0xffffffffu, c.x, o);
                                c.y += __shfl_xor_sync(0xffffffffu, c.y, o);
                                c.z += __shfl_xor_sync(0xffffffffu, c.z, o);
                                c.w += __shfl_xor_sync(0xffffffffu, c.w, o);
                            }
                            if (gi == 0 && elane) {
                                float* a = s_gw + n * D + 4 * lig;
                                atomicAdd(a, c.x);
                                atomicAdd(a + 1, c.y);
                                atomicAdd(a + 2, c.z);
                                atomicAdd(a + 3, c.w);
                            }
                        }
                    }
                }
            }
            if (p.d_lr && (p.g_dense_w_lr || p.g_lr_bias)) {
                const float dlr = valid ? __ldg(p.d_lr + b) : 0.f;
                for (int n0 = 0; n0 <= Fn; n0 += LPR) {          // role n < Fn: g_dense_w_lr[n]; n == Fn: bias
                    const int n = n0 + lig;
                    float v = 0.f;
                    if (valid && n < Fn) v = __ldg(xb + n) * dlr;
                    else if (n == Fn) v = dlr;
#pragma unroll
                    for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (gi == 0 && n <= Fn) atomicAdd(s_gw + Fn * D + n, v);
                }
            }
        }
        if (kStaged) __syncwarp();
    }
#if RBX_BWD_BULK
    if (kBulk) bulk_wait<0>();
#endif
    if (fuse_num) {
        __syncthreads();
        const int Fn = p.Fn;
        if (p.g_dense_w)
            for (int i = threadIdx.x; i < Fn * LPE; i += kThreads) {
                const int n = i / LPE, c = i - n * LPE;
                red_add_f4(p.g_dense_w + (size_t)p.meta.num_widx[n] * D + 4 * c, *reinterpret_cast<const float4*>(s_gw + n * D + 4 * c));
            }
        if (p.d_lr)
            for (int i = threadIdx.x; i <= Fn; i += kThreads) {
                if (i < Fn && p.g_dense_w_lr) red_add_f1(p.g_dense_w_lr + p.meta.num_widx[i], s_gw[Fn * D + i]);
                if (i == Fn && p.g_lr_bias) red_add_f1(p.g_lr_bias, s_gw[Fn * D + i]);
            }
    }
}

// backward, scalar path (any D): warp per sample
__global__ void __launch_bounds__(kThreads) k_embed_fm_bwd_scalar(const __grid_constant__ BwdParams p) {
    const int lane = threadIdx.x & 31;
    const int F = p.F, Ft = p.Ft, D = p.D;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
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
            if (p.g_table_lr && p.d_lr && lane == 0) red_add_f1(p.g_table_lr + (r + p.meta.lr_delta[f]), dlr);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward, numeric slots + bias: batch reductions (no gather, pure streaming)
//   g_dense_w[n,:]  += sum_b x[b,n] * (dE[b,pos_n,:] + d_fm[b]*(S[b,:] - x[b,n] w[n,:]))
//   g_dense_w_lr[n] += sum_b x[b,n] * d_lr[b] ;  g_lr_bias += sum_b d_lr[b]
// Vector kernel (D % 4 == 0): thread <-> (sample lane `sub`, slot n, float4 column c); SUB sample
// lanes per CTA run in parallel, partials meet in shared memory, one red per (n, c) per CTA.
// The first-order roles follow in the same launch.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_dense_w_bwd_vec(const __grid_constant__ BwdParams p) {
    __shared__ float4 s_acc[kThreads];
    __shared__ float s_lr[RBX_MAX_SLOTS + 1];
    for (int i = threadIdx.x; i <= p.Fn; i += kThreads) s_lr[i] = 0.f;
    const int Fn = p.Fn, D = p.D, Ft = p.Ft, V = D / 4;
    const int R4 = Fn * V;                       // vector roles (<= kThreads, host-checked)
    const int SUB = kThreads / R4;               // sample lanes
    const int t = threadIdx.x;
    const bool has_fm = p.d_fm != nullptr;
    if (t < SUB * R4) {
        const int sub = t / R4, role = t - sub * R4;
        const int n = role / V, c = role - n * V;
        const size_t wo = (size_t)p.meta.num_widx[n] * D + 4 * c;
        const float4 w = p.dense_w ? ld_row_f4(p.dense_w + wo) : make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t po = (size_t)p.meta.num_pos[n] * D + 4 * c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.g_dense_w) {
#pragma unroll 4
            for (int64_t b = (int64_t)blockIdx.x * SUB + sub; b < p.B; b += (int64_t)gridDim.x * SUB) {
                const float x = __ldg(p.dense_x + b * Fn + n);
                float4 g = p.dE ? ld_stream_f4(p.dE + (size_t)b * Ft * D + po) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_fm) {
                    const float4 s4 = ld_row_f4(p.S + (size_t)b * D + 4 * c);
                    g = f4_fma(f4_sub(s4, f4_scale(w, x)), __ldg(p.d_fm + b), g);
                }
                acc = f4_fma(g, x, acc);
            }
        }
        s_acc[t] = acc;
    }
    __syncthreads();
    if (t < R4 && p.g_dense_w) {
        float4 a = s_acc[t];
        for (int sub = 1; sub < SUB; ++sub) a = f4_add(a, s_acc[sub * R4 + t]);
        const int n = t / V, c = t - n * V;
        red_add_f4(p.g_dense_w + (size_t)p.meta.num_widx[n] * D + 4 * c, a);
    }
    // first-order roles, all threads: role j < Fn -> g_dense_w_lr[j]; j == Fn -> bias; `lanes` sample
    // lanes per role meet in shared memory, one global atomic per role per CTA
    if (p.d_lr && (p.g_dense_w_lr || p.g_lr_bias)) {
        const int roles = Fn + 1, lanes = kThreads / roles;
        if (t < lanes * roles) {
            const int ln = t / roles, j = t - ln * roles;
            float acc = 0.f;
#pragma unroll 4
            for (int64_t b = (int64_t)blockIdx.x * lanes + ln; b < p.B; b += (int64_t)gridDim.x * lanes)
                acc = fmaf(j < Fn ? __ldg(p.dense_x + b * Fn + j) : 1.f, __ldg(p.d_lr + b), acc);
            atomicAdd(&s_lr[j], acc);
        }
        __syncthreads();
        if (t < Fn && p.g_dense_w_lr) atomicAdd(p.g_dense_w_lr + p.meta.num_widx[t], s_lr[t]);
        if (t == Fn && p.g_lr_bias) atomicAdd(p.g_lr_bias, s_lr[t]);
    }
}

// generic scalar kernel: role r of a CTA column (any D, any Fn); also covers the first-order roles
// when the vector kernel has no spare threads (first_role selects where to start)
__global__ void __launch_bounds__(kThreads) k_dense_w_bwd(const __grid_constant__ BwdParams p, int first_role) {
    const int Fn = p.Fn, D = p.D, Ft = p.Ft;
    const int role = first_role + blockIdx.y * kThreads + threadIdx.x;
    const int n_w = Fn * D;
    if (role > n_w + Fn) return;
    float acc = 0.f;
    if (role < n_w) {
        if (!p.g_dense_w) return;
        const int n = role / D, d = role - n * D;
        const size_t wo = (size_t)p.meta.num_widx[n] * D + d;
        const float w = p.dense_w ? __ldg(p.dense_w + wo) : 0.f;
        const size_t po = (size_t)p.meta.num_pos[n] * D + d;
        const bool has_fm = p.d_fm != nullptr;
#pragma unroll 4
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
            const float x = __ldg(p.dense_x + b * Fn + n);
            float g = p.dE ? ld_stream_f1(p.dE + (size_t)b * Ft * D + po) : 0.f;
            if (has_fm) g = fmaf(__ldg(p.S + (size_t)b * D + d) - x * w, __ldg(p.d_fm + b), g);
            acc = fmaf(x, g, acc);
        }
        atomicAdd(p.g_dense_w + wo, acc);
    } else if (role < n_w + Fn) {
        if (!p.g_dense_w_lr || !p.d_lr) return;
        const int n = role - n_w;
#pragma unroll 4
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x)
            acc = fmaf(__ldg(p.dense_x + b * Fn + n), __ldg(p.d_lr + b), acc);
        atomicAdd(p.g_dense_w_lr + p.meta.num_widx[n], acc);
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
// resident CTAs per SM for (kernel, dynamic smem), cached per (kernel, smem)
int occupancy(const void* kernel, size_t smem) {
    struct Slot { const void* k; size_t smem; int occ; };
    static thread_local Slot cache[64];
    static thread_local int n_cache = 0;
    for (int i = 0; i < n_cache; ++i)
        if (cache[i].k == kernel && cache[i].smem == smem) return cache[i].occ;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || occ <= 0) {
        cudaGetLastError();
        occ = 2;
    }
    if (n_cache < 64) cache[n_cache++] = Slot{kernel, smem, occ};
    return occ;
}

int grid_for(const void* kernel, size_t smem, int64_t warps_needed) {
    int64_t ctas = (warps_needed + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)rbx_sm_count() * occupancy(kernel, smem);
    if (ctas > cap) ctas = cap;
    return ctas < 1 ? 1 : (int)ctas;
}

int fill_meta(SlotMeta& m, const int32_t* cat_pos, int F, const int32_t* num_pos, int Fn, int Ft, const int32_t* num_widx,
              const int32_t* lr_delta, const char* who) {
    for (int f = 0; f < F; ++f) m.lr_delta[f] = lr_delta ? lr_delta[f] : 0;
    for (int n = 0; n < Fn; ++n) {
        const int w = num_widx ? num_widx[n] : n;
        if (w < 0 || w > INT16_MAX) return rbx_fail(RBX_ERR_ARG, "%s: num_widx[%d]=%d", who, n, w);
        m.num_widx[n] = (int16_t)w;
    }
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

// dynamic shared memory of the staged kernels: kWarps x 2 buffers of (SPW*words + skew slack) + barriers
size_t staged_smem(int spw, int F, int Fn) {
    const size_t wi = ((size_t)spw * F + 6) & ~(size_t)3, wx = Fn ? (((size_t)spw * Fn + 6) & ~(size_t)3) : 0;
    return (size_t)kWarps * 2 * (wi + wx) * 4 + (size_t)kWarps * 2 * 8;
}

template <typename K>
int set_smem_limit(K kernel, size_t smem) {
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    }
    return 0;
}

// sharded: 0 local table, 1 row shards, 2 row shards in the ROW+LR layout (LPR lanes span the 2 D-float physical row)
template <int LPR, int U>
int launch_fwd(FwdParams& p, bool staged, int sharded, cudaStream_t st) {
    constexpr int SPW = 32 / LPR;
    const int64_t groups = (p.B + SPW - 1) / SPW;
    const size_t smem = staged_smem(SPW, p.F, p.Fn);
    if (sharded == 2) {
        if constexpr (LPR >= 2 && LPR <= 8) {
            if (!staged || smem > 200 * 1024 || set_smem_limit(k_embed_fm_fwd<LPR, U, true, 2>, smem) != 0) return -1;
            const int grid = grid_for((const void*)k_embed_fm_fwd<LPR, U, true, 2>, smem, groups);
            k_embed_fm_fwd<LPR, U, true, 2><<<grid, kThreads, smem, st>>>(p);
            return 0;
        } else {
            return -1;
        }
    }
    if (sharded) {   // the peer-memory variant exists in its staged form only
        if (!staged || smem > 200 * 1024 || set_smem_limit(k_embed_fm_fwd<LPR, U, true, 1>, smem) != 0) return -1;
        const int grid = grid_for((const void*)k_embed_fm_fwd<LPR, U, true, 1>, smem, groups);
        k_embed_fm_fwd<LPR, U, true, 1><<<grid, kThreads, smem, st>>>(p);
        return 0;
    }
    if (staged) {
        if (smem <= 200 * 1024 && set_smem_limit(k_embed_fm_fwd<LPR, U, true, 0>, smem) == 0) {
            const int grid = grid_for((const void*)k_embed_fm_fwd<LPR, U, true, 0>, smem, groups);
            k_embed_fm_fwd<LPR, U, true, 0><<<grid, kThreads, smem, st>>>(p);
            return 0;
        }
    }
    const int grid = grid_for((const void*)k_embed_fm_fwd<LPR, U, false, 0>, 0, groups);
    k_embed_fm_fwd<LPR, U, false, 0><<<grid, kThreads, 0, st>>>(p);
    return 0;
}

template <int KD>
void launch_fwd_scalar(const FwdParams& p, cudaStream_t st) {
    const int grid = grid_for((const void*)k_embed_fm_fwd_scalar<KD>, 0, p.B);
    k_embed_fm_fwd_scalar<KD><<<grid, kThreads, 0, st>>>(p);
}

template <int LPR, int U>
int launch_bwd(BwdParams& p, bool staged, int sharded, cudaStream_t st) {
    constexpr int SPW = 32 / LPR;
    const int64_t groups = (p.B + SPW - 1) / SPW;
    size_t smem = staged_smem(SPW, p.F, 0) + p.num_smem;
#if RBX_BWD_BULK
    if (staged && U * SPW <= 32 && (RBX_BWD_BULK == 1 || sharded != 0))
        smem += (size_t)kWarps * 2 * (U * SPW) * (4 * LPR * 4 + 8);   // gradient-row staging + destinations
#endif
    if (sharded == 2) {
        if constexpr (LPR >= 2 && LPR <= 8) {
            if (!staged || smem > 200 * 1024 || set_smem_limit(k_embed_fm_bwd<LPR, U, true, 2>, smem) != 0) return -1;
            const int grid = grid_for((const void*)k_embed_fm_bwd<LPR, U, true, 2>, smem, groups);
            k_embed_fm_bwd<LPR, U, true, 2><<<grid, kThreads, smem, st>>>(p);
            return 0;
        } else {
            return -1;
        }
    }
    if (sharded) {
        if (!staged || smem > 200 * 1024 || set_smem_limit(k_embed_fm_bwd<LPR, U, true, 1>, smem) != 0) return -1;
        const int grid = grid_for((const void*)k_embed_fm_bwd<LPR, U, true, 1>, smem, groups);
        k_embed_fm_bwd<LPR, U, true, 1><<<grid, kThreads, smem, st>>>(p);
        return 0;
    }
    if (staged) {
        if (smem <= 200 * 1024 && set_smem_limit(k_embed_fm_bwd<LPR, U, true, 0>, smem) == 0) {
            const int grid = grid_for((const void*)k_embed_fm_bwd<LPR, U, true, 0>, smem, groups);
            k_embed_fm_bwd<LPR, U, true, 0><<<grid, kThreads, smem, st>>>(p);
            return 0;
        }
    }
    const int grid = grid_for((const void*)k_embed_fm_bwd<LPR, U, false, 0>, p.num_smem, groups);
    k_embed_fm_bwd<LPR, U, false, 0><<<grid, kThreads, p.num_smem, st>>>(p);
    return 0;
}

inline bool al16(const void* p) { return (uintptr_t)p % 16 == 0; }

// which shard of a row-sharded table this process owns (one process per GPU); -1 = not told
int g_shard_rank = -1;

// ---------------------------------------------------------------------------------------------
// shared bodies of the plain and the sharded entry points
// ---------------------------------------------------------------------------------------------
struct ShardArgs {                    // world == 0: single local table
    int world = 0;
    const float* const* tables = nullptr;
    const float* const* tables_lr = nullptr;
    float* const* g_tables = nullptr;
    float* const* g_tables_lr = nullptr;
    bool row_lr = false;              // ROW+LR layout: shard rows are 2 D floats, first-order weight at column D
};

int shard_log2(int world, const char* who) {
    int l = 0;
    while ((1 << l) < world) ++l;
    if (world < 1 || world > RBX_MAX_WORLD || (1 << l) != world)
        return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: world=%d must be a power of two <= %d", who, world, RBX_MAX_WORLD);
    return l;
}

int embed_fm_fwd_impl(const char* who, const float* table, const float* table_lr, const ShardArgs& sh, const int32_t* rows,
                      const int32_t* cat_pos, const int32_t* lr_delta, const float* dense_x, const float* dense_w,
                      const float* dense_w_lr, const int32_t* num_pos, const int32_t* num_widx, const float* lr_bias,
                      float* E, float* S, float* fm_out, float* lr_out, int64_t B, int64_t R, int F, int Fn, int D,
                      int n_slots, rbx_stream_t stream) {
    RBX_RANGE(who);
    const bool sharded = sh.world > 0;
    RBX_REQUIRE(B >= 0 && F >= 0 && Fn >= 0 && D >= 1, "%s: negative size", who);
    RBX_REQUIRE(F <= RBX_MAX_SLOTS && Fn <= RBX_MAX_SLOTS, "%s: more than %d slots", who, RBX_MAX_SLOTS);
    RBX_REQUIRE(D <= RBX_MAX_DIM, "%s: D=%d > %d", who, D, RBX_MAX_DIM);
    RBX_REQUIRE(R >= 0 && R <= INT32_MAX, "%s: R=%lld outside int32 row ids", who, (long long)R);
    if (B == 0 || F + Fn == 0) return RBX_OK;
    if (n_slots <= 0) n_slots = F + Fn;
    RBX_REQUIRE(n_slots >= F + Fn, "%s: n_slots=%d < F + Fn", who, n_slots);
    const bool lr_only = !E && !S && !fm_out;   // LogisticRegression alone: no D-dim tables needed
    RBX_REQUIRE(F == 0 || (rows && cat_pos && (table || sharded || lr_only)), "%s: table/rows/cat_pos required when F > 0", who);
    RBX_REQUIRE(Fn == 0 || (dense_x && num_pos && (dense_w || lr_only)), "%s: dense_x/dense_w/num_pos required when Fn > 0", who);
    RBX_REQUIRE(!lr_out || ((F == 0 || table_lr || (sharded && (sh.tables_lr || sh.row_lr))) && (Fn == 0 || dense_w_lr)),
                "%s: lr_out needs table_lr / dense_w_lr", who);
    RBX_REQUIRE(!sh.row_lr || D == 4 || D == 8 || D == 16, "%s: the ROW+LR layout covers D in {4, 8, 16} (D=%d)", who, D);
    if (lr_only && !sharded) { table = nullptr; dense_w = nullptr; D = 16; }
    FwdParams p;
    p.table = table; p.table_lr = table_lr; p.rows = rows; p.dense_x = dense_x; p.dense_w = dense_w;
    p.dense_w_lr = dense_w_lr; p.lr_bias = lr_bias; p.E = E; p.S = S; p.fm_out = fm_out; p.lr_out = lr_out;
    p.B = B; p.R = R; p.F = F; p.Fn = Fn; p.D = D; p.Ft = n_slots; p.wlog2 = 0;
    bool shard_aligned = true;
    if (sharded) {
        const int l = shard_log2(sh.world, who);
        if (l < 0) return l;
        p.wlog2 = l;
        for (int w = 0; w < sh.world; ++w) {
            RBX_REQUIRE(sh.tables && sh.tables[w], "%s: shard table %d is null", who, w);
            p.shard[w] = sh.tables[w];
            p.shard_lr[w] = sh.tables_lr ? sh.tables_lr[w] : nullptr;
            RBX_REQUIRE(!lr_out || F == 0 || sh.row_lr || p.shard_lr[w], "%s: shard lr table %d is null", who, w);
            shard_aligned = shard_aligned && (uintptr_t)p.shard[w] % 16 == 0;
        }
        RBX_REQUIRE(!lr_delta, "%s: sharded tables share one row numbering (lr_delta must be NULL)", who);
    }
    if (int rc = fill_meta(p.meta, cat_pos, F, num_pos, Fn, n_slots, num_widx, lr_delta, who)) return rc;
    cudaStream_t st = rbx_cast_stream(stream);
    const bool aligned = al16(table) && al16(dense_w) && al16(E) && al16(S) && shard_aligned;
    if (D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && aligned) {
        const bool staged = al16(rows) && al16(dense_x);
        const int smode = sh.row_lr ? 2 : (sharded ? 1 : 0);
        int rc = 0;
        switch (sh.row_lr ? D / 2 : D / 4) {      // lanes per physical row
            case 1: rc = launch_fwd<1, RBX_FWD_U>(p, staged, smode, st); break;
            case 2: rc = launch_fwd<2, RBX_FWD_U>(p, staged, smode, st); break;
            case 4: rc = launch_fwd<4, RBX_FWD_U>(p, staged, smode, st); break;
            case 8: rc = launch_fwd<8, RBX_FWD_U>(p, staged, smode, st); break;
            case 16: rc = launch_fwd<16, RBX_FWD_U>(p, staged, smode, st); break;
            default: rc = launch_fwd<32, RBX_FWD_U>(p, staged, smode, st); break;
        }
        if (rc) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: sharded path needs 16-byte aligned rows / dense_x", who);
    } else {
        if (sharded)
            return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: sharded path covers D in {4,8,16,32,64,128} with 16-byte aligned buffers (D=%d)", who, D);
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

int embed_fm_bwd_impl(const char* who, const float* table, const ShardArgs& sh, const int32_t* rows, const int32_t* cat_pos,
                      const int32_t* pad_row, const int32_t* lr_delta, const float* dense_x, const float* dense_w,
                      const int32_t* num_pos, const int32_t* num_widx, const float* E, const float* S, const float* dE,
                      const float* d_fm, const float* d_lr, float* g_table, float* g_table_lr, float* g_dense_w,
                      float* g_dense_w_lr, float* g_lr_bias, int64_t B, int64_t R, int F, int Fn, int D, int n_slots,
                      rbx_stream_t stream) {
    RBX_RANGE(who);
    const bool sharded = sh.world > 0;
    RBX_REQUIRE(B >= 0 && F >= 0 && Fn >= 0 && D >= 1, "%s: negative size", who);
    RBX_REQUIRE(F <= RBX_MAX_SLOTS && Fn <= RBX_MAX_SLOTS, "%s: more than %d slots", who, RBX_MAX_SLOTS);
    RBX_REQUIRE(D <= RBX_MAX_DIM, "%s: D=%d > %d", who, D, RBX_MAX_DIM);
    RBX_REQUIRE(R >= 0 && R <= INT32_MAX, "%s: R=%lld outside int32 row ids", who, (long long)R);
    if (B == 0 || F + Fn == 0) return RBX_OK;
    if (n_slots <= 0) n_slots = F + Fn;
    RBX_REQUIRE(n_slots >= F + Fn, "%s: n_slots=%d < F + Fn", who, n_slots);
    RBX_REQUIRE(F == 0 || (rows && cat_pos), "%s: rows/cat_pos required when F > 0", who);
    const bool lr_only = !dE && !d_fm;
    RBX_REQUIRE(Fn == 0 || (dense_x && num_pos && (dense_w || lr_only)), "%s: dense_x/dense_w/num_pos required when Fn > 0", who);
    RBX_REQUIRE(!d_fm || S, "%s: S (saved by the forward) required with d_fm", who);
    if (lr_only && !sharded) { g_table = nullptr; g_dense_w = nullptr; if (!table && !E) D = 16; }
    RBX_REQUIRE(!d_fm || E || table || sharded || F == 0, "%s: E or table required with d_fm", who);
    BwdParams p;
    p.table = table; p.rows = rows; p.dense_x = dense_x; p.dense_w = dense_w; p.E = E; p.S = S; p.dE = dE;
    p.d_fm = d_fm; p.d_lr = d_lr; p.g_table = g_table; p.g_table_lr = g_table_lr; p.g_dense_w = g_dense_w;
    p.g_dense_w_lr = g_dense_w_lr; p.g_lr_bias = g_lr_bias; p.B = B; p.R = R; p.F = F; p.Fn = Fn; p.D = D; p.Ft = n_slots;
    p.wlog2 = 0;
    p.self_shard = g_shard_rank;
    bool shard_aligned = true, any_g = g_table != nullptr, any_g_lr = g_table_lr != nullptr;
    if (sharded) {
        const int l = shard_log2(sh.world, who);
        if (l < 0) return l;
        p.wlog2 = l;
        any_g = sh.g_tables && sh.g_tables[0] && (!lr_only || sh.row_lr);
        any_g_lr = sh.g_tables_lr && sh.g_tables_lr[0];
        RBX_REQUIRE(!sh.row_lr || D == 4 || D == 8 || D == 16, "%s: the ROW+LR layout covers D in {4, 8, 16} (D=%d)", who, D);
        for (int w = 0; w < sh.world; ++w) {
            p.shard[w] = sh.tables ? sh.tables[w] : nullptr;
            p.g_shard[w] = any_g ? sh.g_tables[w] : nullptr;
            p.g_shard_lr[w] = any_g_lr ? sh.g_tables_lr[w] : nullptr;
            RBX_REQUIRE(!any_g || p.g_shard[w], "%s: shard grad table %d is null", who, w);
            RBX_REQUIRE(!any_g_lr || p.g_shard_lr[w], "%s: shard lr grad table %d is null", who, w);
            RBX_REQUIRE(!(d_fm && !E) || p.shard[w], "%s: shard table %d needed to re-gather e", who, w);
            shard_aligned = shard_aligned && (uintptr_t)p.shard[w] % 16 == 0 && (uintptr_t)p.g_shard[w] % 16 == 0;
        }
        RBX_REQUIRE(!lr_delta, "%s: sharded tables share one row numbering (lr_delta must be NULL)", who);
    }
    if (int rc = fill_meta(p.meta, cat_pos, F, num_pos, Fn, n_slots, num_widx, lr_delta, who)) return rc;
    for (int f = 0; f < F; ++f) p.pad_row[f] = pad_row ? pad_row[f] : -1;
    cudaStream_t st = rbx_cast_stream(stream);

    const bool want_w = Fn > 0 && g_dense_w && (dE || d_fm);
    const bool want_lr = d_lr && ((Fn > 0 && g_dense_w_lr) || g_lr_bias);
    bool num_fused = false;
    p.num_smem = 0;
    if (F > 0 && (any_g || (any_g_lr && d_lr)) && (dE || d_fm || d_lr)) {
        const bool aligned = al16(table) && al16(E) && al16(S) && al16(dE) && al16(g_table) && shard_aligned;
        if (D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0 && aligned) {
            const bool staged = al16(rows);
            const int smode = sh.row_lr ? 2 : (sharded ? 1 : 0);
            // fold the numeric-slot / bias reductions into this launch (dE is then read exactly once)
            if (RBX_BWD_FUSE_NUM && (want_w || want_lr) && (Fn == 0 || dense_x) && Fn * D <= 4096 && al16(dense_w) &&
                al16(g_dense_w)) {
                p.num_smem = (int)((((size_t)Fn * D + Fn + 1) * 4 + 15) & ~(size_t)15);
                num_fused = true;
            }
            int rc = 0;
            switch (sh.row_lr ? D / 2 : D / 4) {
                case 1: rc = launch_bwd<1, RBX_BWD_U>(p, staged, smode, st); break;
                case 2: rc = launch_bwd<2, RBX_BWD_U>(p, staged, smode, st); break;
                case 4: rc = launch_bwd<4, RBX_BWD_U>(p, staged, smode, st); break;
                case 8: rc = launch_bwd<8, RBX_BWD_U>(p, staged, smode, st); break;
                case 16: rc = launch_bwd<16, RBX_BWD_U>(p, staged, smode, st); break;
                default: rc = launch_bwd<32, RBX_BWD_U>(p, staged, smode, st); break;
            }
            if (rc) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: sharded path needs 16-byte aligned rows", who);
        } else {
            if (sharded)
                return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: sharded path covers D in {4,8,16,32,64,128} with 16-byte aligned buffers (D=%d)", who, D);
            const int grid = grid_for((const void*)k_embed_fm_bwd_scalar, 0, B);
            k_embed_fm_bwd_scalar<<<grid, kThreads, 0, st>>>(p);
        }
        RBX_LAUNCH_CHECK(who);
    }
    if ((want_w || want_lr) && !num_fused) {
        int first_scalar_role = 0;        // roles still to be covered by the scalar kernel
        bool scalar_needed = true;
        const int R4 = Fn * (D / 4);
        const bool vec = want_w && D % 4 == 0 && R4 >= 1 && R4 <= kThreads && al16(dense_w) && al16(dE) && al16(S) && al16(g_dense_w);
        if (vec) {
            const int SUB = kThreads / R4;
            int64_t gx = (B + SUB - 1) / SUB;
            const int64_t cap = (int64_t)rbx_sm_count() * 8;
            if (gx > cap) gx = cap;
            k_dense_w_bwd_vec<<<(int)gx, kThreads, 0, st>>>(p);
            RBX_LAUNCH_CHECK(who);
            scalar_needed = false;                      // the vector kernel also covers the first-order roles
        } else if (!want_w) {
            first_scalar_role = Fn * D;
        }
        if (scalar_needed) {
            const int roles = Fn * D + Fn + 1 - first_scalar_role;
            int gx = rbx_sm_count() * 8;
            if (gx > B) gx = (int)B;
            dim3 grid(gx, (roles + kThreads - 1) / kThreads);
            k_dense_w_bwd<<<grid, kThreads, 0, st>>>(p, first_scalar_role);
            RBX_LAUNCH_CHECK(who);
        }
    }
    return RBX_OK;
}

}  // namespace

extern "C" {

int rbx_shard_set_rank(int rank) {
    if (rank < -1 || rank >= RBX_MAX_WORLD) return rbx_fail(RBX_ERR_ARG, "rbx_shard_set_rank: rank %d outside [-1, %d)", rank, RBX_MAX_WORLD);
    g_shard_rank = rank;
    return RBX_OK;
}

int rbx_embed_fm_fwd(const float* table, const float* table_lr, const int32_t* rows, const int32_t* cat_pos,
                     const int32_t* lr_delta, const float* dense_x, const float* dense_w, const float* dense_w_lr,
                     const int32_t* num_pos, const int32_t* num_widx, const float* lr_bias, float* E, float* S,
                     float* fm_out, float* lr_out, int64_t B, int64_t R, int F, int Fn, int D, int n_slots,
                     rbx_stream_t stream) {
    return embed_fm_fwd_impl("rbx_embed_fm_fwd", table, table_lr, ShardArgs(), rows, cat_pos, lr_delta, dense_x, dense_w,
                             dense_w_lr, num_pos, num_widx, lr_bias, E, S, fm_out, lr_out, B, R, F, Fn, D, n_slots, stream);
}

int rbx_embed_fm_bwd(const float* table, const int32_t* rows, const int32_t* cat_pos, const int32_t* pad_row,
                     const int32_t* lr_delta, const float* dense_x, const float* dense_w, const int32_t* num_pos,
                     const int32_t* num_widx, const float* E, const float* S, const float* dE, const float* d_fm,
                     const float* d_lr, float* g_table, float* g_table_lr, float* g_dense_w, float* g_dense_w_lr,
                     float* g_lr_bias, int64_t B, int64_t R, int F, int Fn, int D, int n_slots, rbx_stream_t stream) {
    return embed_fm_bwd_impl("rbx_embed_fm_bwd", table, ShardArgs(), rows, cat_pos, pad_row, lr_delta, dense_x, dense_w,
                             num_pos, num_widx, E, S, dE, d_fm, d_lr, g_table, g_table_lr, g_dense_w, g_dense_w_lr,
                             g_lr_bias, B, R, F, Fn, D, n_slots, stream);
}

int rbx_embed_fm_fwd_sharded(const float* const* shard_tables, const float* const* shard_tables_lr, int world,
                             const int32_t* rows, const int32_t* cat_pos, const float* dense_x, const float* dense_w,
                             const float* dense_w_lr, const int32_t* num_pos, const int32_t* num_widx,
                             const float* lr_bias, float* E, float* S, float* fm_out, float* lr_out, int64_t B, int64_t R,
                             int F, int Fn, int D, int n_slots, rbx_stream_t stream) {
    const char* who = "rbx_embed_fm_fwd_sharded";
    RBX_REQUIRE(world >= 1 && shard_tables, "%s: shard tables required", who);
    ShardArgs sh;
    sh.world = world; sh.tables = shard_tables; sh.tables_lr = shard_tables_lr;
    return embed_fm_fwd_impl(who, nullptr, nullptr, sh, rows, cat_pos, nullptr, dense_x, dense_w, dense_w_lr, num_pos,
                             num_widx, lr_bias, E, S, fm_out, lr_out, B, R, F, Fn, D, n_slots, stream);
}

int rbx_embed_fm_bwd_sharded(const float* const* shard_tables, float* const* shard_g_tables,
                             float* const* shard_g_tables_lr, int world, const int32_t* rows, const int32_t* cat_pos,
                             const int32_t* pad_row, const float* dense_x, const float* dense_w, const int32_t* num_pos,
                             const int32_t* num_widx, const float* E, const float* S, const float* dE, const float* d_fm,
                             const float* d_lr, float* g_dense_w, float* g_dense_w_lr, float* g_lr_bias, int64_t B,
                             int64_t R, int F, int Fn, int D, int n_slots, rbx_stream_t stream) {
    const char* who = "rbx_embed_fm_bwd_sharded";
    RBX_REQUIRE(world >= 1, "%s: world", who);
    ShardArgs sh;
    sh.world = world; sh.tables = shard_tables; sh.g_tables = shard_g_tables; sh.g_tables_lr = shard_g_tables_lr;
    return embed_fm_bwd_impl(who, nullptr, sh, rows, cat_pos, pad_row, nullptr, dense_x, dense_w, num_pos, num_widx, E, S,
                             dE, d_fm, d_lr, nullptr, nullptr, g_dense_w, g_dense_w_lr, g_lr_bias, B, R, F, Fn, D, n_slots,
                             stream);
}


// ROW+LR layout: every shard is [cap, 2 D] floats, row = [e_0 .. e_{D-1} | w_lr | 0 ...]; the
// gradient shards mirror it.  One request per slot over NVLink instead of two.
int rbx_embed_fm_fwd_sharded_rowlr(const float* const* shard_tables, int world, const int32_t* rows, const int32_t* cat_pos,
                                   const float* dense_x, const float* dense_w, const float* dense_w_lr,
                                   const int32_t* num_pos, const int32_t* num_widx, const float* lr_bias, float* E,
                                   float* S, float* fm_out, float* lr_out, int64_t B, int64_t R, int F, int Fn, int D,
                                   int n_slots, rbx_stream_t stream) {
    const char* who = "rbx_embed_fm_fwd_sharded_rowlr";
    RBX_REQUIRE(world >= 1 && shard_tables, "%s: shard tables required", who);
    ShardArgs sh;
    sh.world = world; sh.tables = shard_tables; sh.row_lr = true;
    return embed_fm_fwd_impl(who, nullptr, nullptr, sh, rows, cat_pos, nullptr, dense_x, dense_w, dense_w_lr, num_pos,
                             num_widx, lr_bias, E, S, fm_out, lr_out, B, R, F, Fn, D, n_slots, stream);
}

int rbx_embed_fm_bwd_sharded_rowlr(const float* const* shard_tables, float* const* shard_g_tables, int world,
                                   const int32_t* rows, const int32_t* cat_pos, const int32_t* pad_row,
                                   const float* dense_x, const float* dense_w, const int32_t* num_pos,
                                   const int32_t* num_widx, const float* E, const float* S, const float* dE,
                                   const float* d_fm, const float* d_lr, float* g_dense_w, float* g_dense_w_lr,
                                   float* g_lr_bias, int64_t B, int64_t R, int F, int Fn, int D, int n_slots,
                                   rbx_stream_t stream) {
    const char* who = "rbx_embed_fm_bwd_sharded_rowlr";
    RBX_REQUIRE(world >= 1 && shard_g_tables, "%s: gradient shards required", who);
    ShardArgs sh;
    sh.world = world; sh.tables = shard_tables; sh.g_tables = shard_g_tables; sh.row_lr = true;
    return embed_fm_bwd_impl(who, nullptr, sh, rows, cat_pos, pad_row, nullptr, dense_x, dense_w, num_pos, num_widx, E, S,
                             dE, d_fm, d_lr, nullptr, nullptr, g_dense_w, g_dense_w_lr, g_lr_bias, B, R, F, Fn, D, n_slots,
                             stream);
}

}  // extern "C"
