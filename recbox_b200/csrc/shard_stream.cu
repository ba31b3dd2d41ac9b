// (e) streamed exchange for the row-sharded table (BASELINE configs[3]): the all-to-all of SURVEY.md
// section 8e done by the kernels themselves, with every random access local to the row's owner and
// only CONTIGUOUS streams crossing NVLink / NVSwitch (posted stores, full 128-byte lines):
//
//   requester q                                            owner w
//   k_xs_route : ids of a tile of T samples bucketed by
//                owner in shared memory, each bucket
//                stored as ONE contiguous run ------------> inbox_ids[parity][q][slot .. slot+n)
//   k_xs_barrier (counts ride along) <------ flags ------> k_xs_barrier
//                                                          k_xs_serve : local gather of row + first-order
//   rowbuf[w][slot ..) <------------------------------------ weight, stored in slot order (contiguous)
//   k_xs_barrier <-------------------- flags ------------> k_xs_barrier
//   k_xs_consume : tile by tile, the W contiguous runs are
//                scattered into a shared-memory E tile, the
//                FM / LR sums are taken per sample, E leaves
//                with coalesced 512-byte stores
//   ........................................ backward ............................................
//   k_xs_grad_push : g = dE + d_fm (S - e) in slot order --> ginbox[q][slot ..)
//   k_xs_barrier <-------------------- flags ------------> k_xs_barrier
//                                                          k_xs_apply : red.add into the local gradient shard
//
// The reference has no counterpart (its only multi-device modes are replicas); the contract is that the
// results equal the single-table fused kernels' (tests/test_sharded_gpu.py).  Slots: a tile's ids for owner
// w occupy [tile_base[tile][w], +tile_cnt[tile][w]) of the (owner w, requester q) lane, reserved with one
// atomicAdd per tile and owner; the order inside a run is the (sample, slot) order, recorded in pair_sorted.
// Inboxes are double-buffered by step parity, so no barrier is needed after k_xs_apply.
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPairs = 2048;            // (samples per tile) x (categorical slots) <= kMaxPairs
constexpr int kMaxChunks = kMaxPairs / 32;
constexpr int kMaxF = 64;
constexpr int kW = RBX_MAX_WORLD;

struct Peers {
    void* p[kW];
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------
// route
// ---------------------------------------------------------------------------------------------
struct RouteParams {
    const int32_t* rows;       // [B, F] global rows
    int64_t B, R, cap;
    int F, T, wlog2, rank;
    int32_t* cursor;           // [kW] slots handed out per owner (zeroed by k_xs_barrier after it is published)
    int32_t* tile_base;        // [n_tiles, kW]
    int32_t* tile_cnt;         // [n_tiles, kW]
    uint16_t* pair_sorted;     // [n_tiles, T*F]  b_local*F + f of the id in each slot of the tile's runs
    int32_t* overflow;         // [1]
    Peers inbox_ids;           // owner w's id inbox of this parity: int32 [kW(requester), cap]
};

__global__ void __launch_bounds__(kThreads) k_xs_route(const __grid_constant__ RouteParams p) {
    __shared__ int32_t s_pre[kMaxChunks][kW];     // per 32-id chunk: count, then exclusive prefix, per owner
    __shared__ int32_t s_base[kW], s_off[kW + 1];
    __shared__ int32_t s_ids[kMaxPairs];
    __shared__ uint16_t s_pair[kMaxPairs];
    const int W = 1 << p.wlog2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = p.F, T = p.T;
    const int64_t n_tiles = (p.B + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * T;
        const int th = (int)(p.B - b0 < T ? p.B - b0 : T);
        const int np = th * F, nch = (np + 31) >> 5;
        const int32_t* rt = p.rows + b0 * F;
        for (int c = warp; c < nch; c += kWarps) {
            const int k = c * 32 + lane;
            const int32_t r = k < np ? __ldg(rt + k) : -1;
            const int o = ((uint32_t)r < (uint64_t)p.R) ? (r & (W - 1)) : -1;
            for (int oo = 0; oo < W; ++oo) {
                const unsigned m = __ballot_sync(0xffffffffu, o == oo);
                if (lane == oo) s_pre[c][oo] = __popc(m);
            }
        }
        __syncthreads();
        if (threadIdx.x < W) {
            const int o = threadIdx.x;
            int run = 0;
            for (int c = 0; c < nch; ++c) {
                const int t = s_pre[c][o];
                s_pre[c][o] = run;
                run += t;
            }
            const int base = atomicAdd(p.cursor + o, run);
            s_base[o] = base;
            s_off[o + 1] = run;
            p.tile_base[tile * kW + o] = base;
            p.tile_cnt[tile * kW + o] = run;
            if ((int64_t)base + run > p.cap) *p.overflow = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_off[0] = 0;
            for (int o = 0; o < W; ++o) s_off[o + 1] += s_off[o];
        }
        __syncthreads();
        for (int c = warp; c < nch; c += kWarps) {
            const int k = c * 32 + lane;
            const int32_t r = k < np ? __ldg(rt + k) : -1;
            const int o = ((uint32_t)r < (uint64_t)p.R) ? (r & (W - 1)) : -1;
            int rank_in = 0;
            for (int oo = 0; oo < W; ++oo) {
                const unsigned m = __ballot_sync(0xffffffffu, o == oo);
                if (o == oo) rank_in = __popc(m & ((1u << lane) - 1u));
            }
            if (o >= 0) {
                const int pos = s_off[o] + s_pre[c][o] + rank_in;
                s_ids[pos] = r >> p.wlog2;
                s_pair[pos] = (uint16_t)k;
            }
        }
        __syncthreads();
        const int nv = s_off[W];
        uint16_t* ps = p.pair_sorted + tile * (int64_t)(T * F);
        for (int i = threadIdx.x; i < nv; i += kThreads) {
            int o = 0;
            while (i >= s_off[o + 1]) ++o;
            const int64_t slot = (int64_t)s_base[o] + (i - s_off[o]);
            if (slot < p.cap) reinterpret_cast<int32_t*>(p.inbox_ids.p[o])[(int64_t)p.rank * p.cap + slot] = s_ids[i];
            ps[i] = s_pair[i];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// cross-rank barrier over peer-mapped flags (one CTA); optionally publishes the route counts
// ---------------------------------------------------------------------------------------------
struct BarrierParams {
    Peers flags;               // rank w's flag words: uint32 [kW]; word [me] is written by me
    Peers meta;                // rank w's count inbox of this parity: int32 [kW]; word [me] = ids I sent to w (nullable)
    int32_t* cursor;           // [kW] | NULL
    int rank, world;
    uint32_t epoch;
};

__global__ void __launch_bounds__(32) k_xs_barrier(const __grid_constant__ BarrierParams p) {
    const int t = threadIdx.x;
    if (t < p.world) {
        if (p.cursor) {
            reinterpret_cast<int32_t*>(p.meta.p[t])[p.rank] = p.cursor[t];
            p.cursor[t] = 0;
        }
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(p.flags.p[t]) + p.rank, p.epoch);
    }
    if (t < p.world) {
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(p.flags.p[p.rank]) + t;
        while ((int32_t)(ld_acquire_sys(mine) - p.epoch) < 0) __nanosleep(64);
    }
}

// ---------------------------------------------------------------------------------------------
// serve (owner): rowbuf_q[me][j] = table[inbox_ids[q][j]]
// ---------------------------------------------------------------------------------------------
struct ServeParams {
    const float* table;
    int64_t rs;                // floats between consecutive local rows
    const float* lr;           // first-order weight of local row r: lr[r * lr_stride] | NULL
    int64_t lr_stride;
    const int32_t* inbox_ids;  // this parity: [kW, cap]
    const int32_t* meta;       // this parity: [kW]
    int64_t cap;
    int rank, world;
    Peers rowbuf;              // requester q's: float [kW(owner), cap, D]
    Peers rowbuf_lr;           // requester q's: float [kW(owner), cap]
};

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_xs_serve(const __grid_constant__ ServeParams p) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR, U = 8;
    const int q = blockIdx.y;
    int64_t n = __ldg(p.meta + q);
    if (n > p.cap) n = p.cap;
    float* out = reinterpret_cast<float*>(p.rowbuf.p[q]) + (size_t)p.rank * p.cap * D;
    float* out_lr = (p.lr && p.rowbuf_lr.p[q]) ? reinterpret_cast<float*>(p.rowbuf_lr.p[q]) + (size_t)p.rank * p.cap : nullptr;
    const int32_t* ids = p.inbox_ids + (size_t)q * p.cap;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * kWarps;
    const uint64_t pol_keep = l2_policy_evict_last();
    for (int64_t base = warp0 * (RPW * U); base < n; base += nwarps * (RPW * U)) {
        int32_t r[U];
        float4 v[U];
        float l[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            r[u] = i < n ? __ldg(ids + i) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            l[u] = 0.f;
            if (r[u] >= 0) {
                v[u] = ld_row_f4_hint(p.table + (size_t)r[u] * p.rs + 4 * lig, pol_keep);
                if (out_lr && lig == 0) l[u] = __ldg(p.lr + (size_t)r[u] * p.lr_stride);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            if (r[u] >= 0) {
                st_stream_f4(out + (size_t)i * D + 4 * lig, v[u]);
                if (out_lr && lig == 0) out_lr[i] = l[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// consume (requester): rowbuf runs -> E tile in shared memory -> FM / LR sums, E out
// ---------------------------------------------------------------------------------------------
struct TileMeta {
    int16_t cat_pos[kMaxF];
    int16_t num_pos[RBX_MAX_SLOTS];
    int16_t num_widx[RBX_MAX_SLOTS];
};

struct ConsumeParams {
    const float* rowbuf;       // local: [kW(owner), cap, D]
    const float* rowbuf_lr;    // local: [kW(owner), cap] | NULL
    const int32_t* tile_base;
    const int32_t* tile_cnt;
    const uint16_t* pair_sorted;
    const float* dense_x;
    const float* dense_w;
    const float* dense_w_lr;
    const float* lr_bias;
    float* E;
    float* S;
    float* fm_out;
    float* lr_out;
    int64_t B, cap;
    int F, Fn, Ft, T, world;
    TileMeta meta;
};

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_xs_consume(const __grid_constant__ ConsumeParams p) {
    constexpr int D = 4 * LPR, SG = 32 / LPR, U = 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int32_t s_base[kW], s_off[kW + 1];
    const int F = p.F, Fn = p.Fn, Ft = p.Ft, T = p.T, W = p.world;
    float* Es = reinterpret_cast<float*>(smem_raw);                 // [T*F, D]
    float* lrs = Es + (size_t)T * F * D;                            // [T*F]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane & (LPR - 1), sg = lane / LPR;
    const float bias = p.lr_bias ? __ldg(p.lr_bias) : 0.f;
    const int64_t n_tiles = (p.B + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * T;
        const int th = (int)(p.B - b0 < T ? p.B - b0 : T);
        const int np = th * F;
        if (threadIdx.x == 0) {
            int run = 0;
            s_off[0] = 0;
            for (int o = 0; o < W; ++o) {
                s_base[o] = p.tile_base[tile * kW + o];
                run += p.tile_cnt[tile * kW + o];
                s_off[o + 1] = run;
            }
        }
        __syncthreads();
        const int nv = s_off[W];
        if (nv < np) {                                               // ids outside the table read as zero rows
            for (int i = threadIdx.x; i < np * LPR; i += kThreads) reinterpret_cast<float4*>(Es)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = threadIdx.x; i < np; i += kThreads) lrs[i] = 0.f;
            __syncthreads();
        }
        const uint16_t* ps = p.pair_sorted + tile * (int64_t)(T * F);
        const int items = nv * LPR;
        for (int i0 = threadIdx.x; i0 < items; i0 += kThreads * U) {
            float4 v[U];
            float l[U];
            int k[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * kThreads;
                k[u] = -1;
                l[u] = 0.f;
                if (i < items) {
                    const int pp = i / LPR, cc = i & (LPR - 1);
                    int o = 0;
                    while (pp >= s_off[o + 1]) ++o;
                    const int64_t slot = (int64_t)s_base[o] + (pp - s_off[o]);
                    if (slot < p.cap) {
                        k[u] = ps[pp];
                        v[u] = ld_stream_f4(p.rowbuf + ((size_t)o * p.cap + slot) * D + 4 * cc);
                        if (cc == 0 && p.rowbuf_lr) l[u] = ld_stream_f1(p.rowbuf_lr + (size_t)o * p.cap + slot);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k[u] >= 0) {
                    const int cc = (i0 + u * kThreads) & (LPR - 1);
                    *reinterpret_cast<float4*>(Es + (size_t)k[u] * D + 4 * cc) = v[u];
                    if (cc == 0) lrs[k[u]] = l[u];
                }
            }
        }
        __syncthreads();
        // per sample: lanes = (slot group sg, float4 column c); 32 / LPR slots per warp instruction
        for (int bl = warp; bl < th; bl += kWarps) {
            const int64_t b = b0 + bl;
            float4 S4 = make_float4(0.f, 0.f, 0.f, 0.f), Q4 = S4;
            float lr = 0.f;
            float* Eb = p.E ? p.E + (size_t)b * Ft * D + 4 * c : nullptr;
            for (int f = sg; f < F; f += SG) {
                const float4 e = *reinterpret_cast<const float4*>(Es + ((size_t)bl * F + f) * D + 4 * c);
                S4 = f4_add(S4, e);
                Q4 = f4_sqacc(e, Q4);
                if (Eb) st_stream_f4(Eb + (size_t)p.meta.cat_pos[f] * D, e);
                if (c == 0) lr += lrs[bl * F + f];
            }
            for (int n = sg; n < Fn; n += SG) {
                const float x = __ldg(p.dense_x + b * Fn + n);
                const int wi = p.meta.num_widx[n];
                const float4 e = f4_scale(ld_row_f4(p.dense_w + (size_t)wi * D + 4 * c), x);
                S4 = f4_add(S4, e);
                Q4 = f4_sqacc(e, Q4);
                if (Eb) st_stream_f4(Eb + (size_t)p.meta.num_pos[n] * D, e);
                if (c == 0 && p.dense_w_lr) lr = fmaf(x, __ldg(p.dense_w_lr + wi), lr);
            }
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
                S4.x += __shfl_xor_sync(0xffffffffu, S4.x, o);
                S4.y += __shfl_xor_sync(0xffffffffu, S4.y, o);
                S4.z += __shfl_xor_sync(0xffffffffu, S4.z, o);
                S4.w += __shfl_xor_sync(0xffffffffu, S4.w, o);
                Q4.x += __shfl_xor_sync(0xffffffffu, Q4.x, o);
                Q4.y += __shfl_xor_sync(0xffffffffu, Q4.y, o);
                Q4.z += __shfl_xor_sync(0xffffffffu, Q4.z, o);
                Q4.w += __shfl_xor_sync(0xffffffffu, Q4.w, o);
            }
            if (p.S && sg == 0) *reinterpret_cast<float4*>(p.S + (size_t)b * D + 4 * c) = S4;
            if (p.fm_out) {
                float fm = (S4.x * S4.x - Q4.x) * 0.5f + (S4.y * S4.y - Q4.y) * 0.5f + (S4.z * S4.z - Q4.z) * 0.5f +
                           (S4.w * S4.w - Q4.w) * 0.5f;
                fm = group_sum<LPR>(fm);
                if (lane == 0) p.fm_out[b] = fm;
            }
            if (p.lr_out) {
                lr = group_sum<32>(lr);
                if (lane == 0) p.lr_out[b] = lr + bias;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// grad push (requester): g = dE + d_fm (S - e), written in slot order into the owners' inboxes
// ---------------------------------------------------------------------------------------------
struct GradParams {
    const float* E;            // [B, Ft, D] | NULL (then e is re-read from rowbuf)
    const float* rowbuf;
    const float* S;
    const float* dE;
    const float* d_fm;
    const float* d_lr;
    const int32_t* rows;       // [B, F]: a slot whose id is its field's padding row gets a zero gradient
    const int32_t* tile_base;
    const int32_t* tile_cnt;
    const uint16_t* pair_sorted;
    int64_t B, cap;
    int F, Ft, T, world, rank;
    uint32_t f_magic;          // ceil(2^20 / F): k / F == (k * f_magic) >> 20 for k < 2048
    Peers ginbox;              // owner w's: float [kW(requester), cap, D]
    Peers ginbox_lr;           // owner w's: float [kW(requester), cap]
    int16_t cat_pos[kMaxF];
    int32_t pad_row[kMaxF];
};

template <int LPR>
__global__ void __launch_bounds__(kThreads, 3) k_xs_grad_push(const __grid_constant__ GradParams p) {
    constexpr int D = 4 * LPR, U = 4;
    __shared__ int32_t s_base[kW], s_off[kW + 1];
    const int F = p.F, Ft = p.Ft, T = p.T, W = p.world;
    const bool has_fm = p.d_fm != nullptr;
    const uint64_t pol_stream = l2_policy_evict_first();
    const int64_t n_tiles = (p.B + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * T;
        if (threadIdx.x == 0) {
            int run = 0;
            s_off[0] = 0;
            for (int o = 0; o < W; ++o) {
                s_base[o] = p.tile_base[tile * kW + o];
                run += p.tile_cnt[tile * kW + o];
                s_off[o + 1] = run;
            }
        }
        __syncthreads();
        const int items = s_off[W] * LPR;
        const uint16_t* ps = p.pair_sorted + tile * (int64_t)(T * F);
        for (int i0 = threadIdx.x; i0 < items; i0 += kThreads * U) {
            float4 g[U], e[U], S4[U];
            float dfm[U], dlr[U];
            float* dst[U];
            float* dst_lr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * kThreads;
                dst[u] = nullptr;
                dst_lr[u] = nullptr;
                g[u] = e[u] = S4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                dfm[u] = dlr[u] = 0.f;
                if (i < items) {
                    const int pp = i / LPR, cc = i & (LPR - 1);
                    int o = 0;
                    while (pp >= s_off[o + 1]) ++o;
                    const int64_t slot = (int64_t)s_base[o] + (pp - s_off[o]);
                    if (slot < p.cap) {
                        const uint32_t k = ps[pp];
                        const uint32_t bl = (k * p.f_magic) >> 20;
                        const int f = (int)(k - bl * F);
                        const int64_t b = b0 + bl;
                        const size_t in_off = ((size_t)p.rank * p.cap + slot);
                        dst[u] = reinterpret_cast<float*>(p.ginbox.p[o]) + in_off * D + 4 * cc;
                        if (cc == 0 && p.d_lr && p.ginbox_lr.p[o]) dst_lr[u] = reinterpret_cast<float*>(p.ginbox_lr.p[o]) + in_off;
                        if (__ldg(p.rows + b * F + f) != p.pad_row[f]) {  // padding rows receive a zero gradient
                            const size_t eo = ((size_t)b * Ft + p.cat_pos[f]) * D + 4 * cc;
                            if (p.dE) g[u] = ld_stream_f4_hint(p.dE + eo, pol_stream);
                            if (has_fm) {
                                e[u] = p.E ? ld_stream_f4_hint(p.E + eo, pol_stream)
                                           : ld_stream_f4(p.rowbuf + ((size_t)o * p.cap + slot) * D + 4 * cc);
                                S4[u] = ld_row_f4(p.S + (size_t)b * D + 4 * cc);
                                dfm[u] = __ldg(p.d_fm + b);
                            }
                            if (dst_lr[u]) dlr[u] = __ldg(p.d_lr + b);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (dst[u]) {
                    if (has_fm) g[u] = f4_fma(f4_sub(S4[u], e[u]), dfm[u], g[u]);
                    st_stream_f4(dst[u], g[u]);
                    if (dst_lr[u]) *dst_lr[u] = dlr[u];
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// apply (owner): g_table[inbox_ids[q][j]] += ginbox[q][j]
// ---------------------------------------------------------------------------------------------
struct ApplyParams {
    const float* ginbox;       // local: [kW, cap, D]
    const float* ginbox_lr;    // local: [kW, cap] | NULL
    const int32_t* inbox_ids;
    const int32_t* meta;
    float* g_table;
    int64_t rs;
    float* g_lr;
    int64_t lr_stride;
    int64_t cap;
};

template <int LPR>
__global__ void __launch_bounds__(kThreads) k_xs_apply(const __grid_constant__ ApplyParams p) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR, U = 8;
    const int q = blockIdx.y;
    int64_t n = __ldg(p.meta + q);
    if (n > p.cap) n = p.cap;
    const int32_t* ids = p.inbox_ids + (size_t)q * p.cap;
    const float* g = p.ginbox + (size_t)q * p.cap * D;
    const float* gl = (p.ginbox_lr && p.g_lr) ? p.ginbox_lr + (size_t)q * p.cap : nullptr;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * kWarps;
    const uint64_t pol_keep = l2_policy_evict_last();
    for (int64_t base = warp0 * (RPW * U); base < n; base += nwarps * (RPW * U)) {
        int32_t r[U];
        float4 v[U];
        float l[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            r[u] = i < n ? __ldg(ids + i) : -1;
            l[u] = 0.f;
            if (r[u] >= 0) {
                v[u] = ld_stream_f4(g + (size_t)i * D + 4 * lig);
                if (gl && lig == 0) l[u] = ld_stream_f1(gl + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r[u] >= 0) {
                red_add_f4_hint(p.g_table + (size_t)r[u] * p.rs + 4 * lig, v[u], pol_keep);
                if (gl && lig == 0) red_add_f1(p.g_lr + (size_t)r[u] * p.lr_stride, l[u]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
inline bool pow2_dim(int D) { return D % 4 == 0 && D >= 4 && D <= 128 && (D & (D - 1)) == 0; }

int fill_peers(Peers& pp, void* const* a, int world, bool required, const char* who, const char* what) {
    for (int w = 0; w < kW; ++w) pp.p[w] = nullptr;
    if (!a) return required ? rbx_fail(RBX_ERR_ARG, "%s: %s pointer array is null", who, what) : RBX_OK;
    for (int w = 0; w < world; ++w) {
        if (!a[w] && required) return rbx_fail(RBX_ERR_ARG, "%s: %s[%d] is null", who, what, w);
        pp.p[w] = a[w];
    }
    return RBX_OK;
}

int wlog2_of(int world) {
    int l = 0;
    while ((1 << l) < world) ++l;
    return ((1 << l) == world && world <= kW) ? l : -1;
}

size_t consume_smem(int T, int F, int D) { return (size_t)T * F * (D + 1) * 4; }

#define XS_DISPATCH_LPR(D, CALL)                          \
    switch ((D) / 4) {                                    \
        case 1: { constexpr int LPR = 1; CALL; } break;   \
        case 2: { constexpr int LPR = 2; CALL; } break;   \
        case 4: { constexpr int LPR = 4; CALL; } break;   \
        case 8: { constexpr int LPR = 8; CALL; } break;   \
        case 16: { constexpr int LPR = 16; CALL; } break; \
        default: { constexpr int LPR = 32; CALL; } break; \
    }

template <int LPR>
int launch_consume(const ConsumeParams& p, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_xs_consume<LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_xs_consume<LPR>, kThreads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    const int64_t n_tiles = (p.B + p.T - 1) / p.T;
    int64_t grid = (int64_t)rbx_sm_count() * occ;
    if (grid > n_tiles) grid = n_tiles;
    k_xs_consume<LPR><<<(int)grid, kThreads, smem, st>>>(p);
    return 0;
}

}  // namespace

extern "C" {

int rbx_xs_tile_samples(int F, int D) {
    if (F < 1 || F > kMaxF || !pow2_dim(D)) return 0;
    int T = 32;
    while (T > 1 && (consume_smem(T, F, D) > 57 * 1024 || T * F > kMaxPairs)) T >>= 1;
    return (consume_smem(T, F, D) <= 200 * 1024 && T * F <= kMaxPairs) ? T : 0;
}

int rbx_xs_route(const int32_t* rows, int64_t B, int F, int64_t R, int D, int rank, int world, int64_t cap, int32_t* cursor, int32_t* tile_base, int32_t* tile_cnt, uint16_t* pair_sorted,
                 int32_t* overflow, void* const* inbox_ids, rbx_stream_t stream) {
    const char* who = "rbx_xs_route";
    const int wl = wlog2_of(world);
    RBX_REQUIRE(wl >= 0 && rank >= 0 && rank < world, "%s: world must be a power of two <= %d (got %d, rank %d)", who, kW, world, rank);
    RBX_REQUIRE(B >= 0 && R >= 0 && cap >= 1, "%s: bad size", who);
    const int T = rbx_xs_tile_samples(F, D);
    if (T == 0) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: F=%d D=%d (covers 1..%d slots, D in {4..128} powers of two)", who, F, D, kMaxF);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(rows && cursor && tile_base && tile_cnt && pair_sorted && overflow, "%s: null pointer", who);
    RouteParams p;
    p.rows = rows; p.B = B; p.R = R; p.cap = cap; p.F = F; p.T = T; p.wlog2 = wl; p.rank = rank;
    p.cursor = cursor; p.tile_base = tile_base; p.tile_cnt = tile_cnt; p.pair_sorted = pair_sorted; p.overflow = overflow;
    if (int rc = fill_peers(p.inbox_ids, inbox_ids, world, true, who, "inbox_ids")) return rc;
    const int64_t n_tiles = (B + T - 1) / T;
    int64_t grid = (int64_t)rbx_sm_count() * 4;
    if (grid > n_tiles) grid = n_tiles;
    k_xs_route<<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_barrier(void* const* flags, void* const* meta, int32_t* cursor, int rank, int world, uint32_t epoch,
                   rbx_stream_t stream) {
    const char* who = "rbx_xs_barrier";
    RBX_REQUIRE(world >= 1 && world <= kW && rank >= 0 && rank < world, "%s: rank/world", who);
    RBX_REQUIRE((cursor == nullptr) == (meta == nullptr), "%s: cursor and meta come together", who);
    BarrierParams p;
    if (int rc = fill_peers(p.flags, flags, world, true, who, "flags")) return rc;
    if (int rc = fill_peers(p.meta, meta, world, cursor != nullptr, who, "meta")) return rc;
    p.cursor = cursor; p.rank = rank; p.world = world; p.epoch = epoch;
    k_xs_barrier<<<1, 32, 0, rbx_cast_stream(stream)>>>(p);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_serve(const float* table, int64_t row_stride, const float* lr, int64_t lr_stride, int D,
                 const int32_t* inbox_ids, const int32_t* meta, int64_t cap, int rank, int world,
                 void* const* rowbuf, void* const* rowbuf_lr, rbx_stream_t stream) {
    const char* who = "rbx_xs_serve";
    RBX_REQUIRE(world >= 1 && world <= kW && rank >= 0 && rank < world && cap >= 1, "%s: bad size", who);
    RBX_REQUIRE(table && inbox_ids && meta && row_stride >= D && row_stride % 4 == 0, "%s: null pointer / bad row stride", who);
    if (!pow2_dim(D)) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: D=%d (covers 4..128, powers of two)", who, D);
    ServeParams p;
    p.table = table; p.rs = row_stride; p.lr = lr; p.lr_stride = lr_stride; p.inbox_ids = inbox_ids; p.meta = meta;
    p.cap = cap; p.rank = rank; p.world = world;
    if (int rc = fill_peers(p.rowbuf, rowbuf, world, true, who, "rowbuf")) return rc;
    if (int rc = fill_peers(p.rowbuf_lr, rowbuf_lr, world, lr != nullptr, who, "rowbuf_lr")) return rc;
    int64_t per = (int64_t)rbx_sm_count() * 6 / world;
    if (per < 1) per = 1;
    XS_DISPATCH_LPR(D, (k_xs_serve<LPR><<<dim3((unsigned)per, world), kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_consume(const float* rowbuf, const float* rowbuf_lr, const int32_t* tile_base, const int32_t* tile_cnt,
                   const uint16_t* pair_sorted, const int32_t* cat_pos, const float* dense_x, const float* dense_w,
                   const float* dense_w_lr, const int32_t* num_pos, const int32_t* num_widx, const float* lr_bias,
                   float* E, float* S, float* fm_out, float* lr_out, int64_t B, int64_t cap, int F, int Fn, int D,
                   int n_slots, int world, rbx_stream_t stream) {
    const char* who = "rbx_xs_consume";
    RBX_REQUIRE(world >= 1 && world <= kW && cap >= 1 && B >= 0, "%s: bad size", who);
    const int T = rbx_xs_tile_samples(F, D);
    if (T == 0) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: F=%d D=%d", who, F, D);
    RBX_REQUIRE(Fn >= 0 && Fn <= RBX_MAX_SLOTS && n_slots >= F + Fn, "%s: slot counts", who);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(rowbuf && tile_base && tile_cnt && pair_sorted && cat_pos, "%s: null pointer", who);
    RBX_REQUIRE(Fn == 0 || (dense_x && dense_w && num_pos), "%s: numeric slots need dense_x / dense_w / num_pos", who);
    ConsumeParams p;
    p.rowbuf = rowbuf; p.rowbuf_lr = rowbuf_lr; p.tile_base = tile_base; p.tile_cnt = tile_cnt; p.pair_sorted = pair_sorted;
    p.dense_x = dense_x; p.dense_w = dense_w; p.dense_w_lr = dense_w_lr; p.lr_bias = lr_bias;
    p.E = E; p.S = S; p.fm_out = fm_out; p.lr_out = lr_out; p.B = B; p.cap = cap;
    p.F = F; p.Fn = Fn; p.Ft = n_slots; p.T = T; p.world = world;
    for (int f = 0; f < F; ++f) {
        RBX_REQUIRE(cat_pos[f] >= 0 && cat_pos[f] < n_slots, "%s: cat_pos[%d]", who, f);
        p.meta.cat_pos[f] = (int16_t)cat_pos[f];
    }
    for (int n = 0; n < Fn; ++n) {
        RBX_REQUIRE(num_pos[n] >= 0 && num_pos[n] < n_slots, "%s: num_pos[%d]", who, n);
        p.meta.num_pos[n] = (int16_t)num_pos[n];
        p.meta.num_widx[n] = (int16_t)(num_widx ? num_widx[n] : n);
    }
    const size_t smem = consume_smem(T, F, D);
    int rc = 0;
    XS_DISPATCH_LPR(D, (rc = launch_consume<LPR>(p, smem, rbx_cast_stream(stream))));
    if (rc != 0) return rbx_fail(RBX_ERR_CUDA, "%s: cannot reserve %zu bytes of shared memory", who, smem);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_grad_push(const float* E, const float* rowbuf, const float* S, const float* dE, const float* d_fm,
                     const float* d_lr, const int32_t* rows, const int32_t* pad_row, const int32_t* tile_base, const int32_t* tile_cnt, const uint16_t* pair_sorted,
                     const int32_t* cat_pos, int64_t B, int64_t cap, int F, int D, int n_slots, int rank, int world,
                     void* const* ginbox, void* const* ginbox_lr, rbx_stream_t stream) {
    const char* who = "rbx_xs_grad_push";
    RBX_REQUIRE(world >= 1 && world <= kW && rank >= 0 && rank < world && cap >= 1 && B >= 0, "%s: bad size", who);
    const int T = rbx_xs_tile_samples(F, D);
    if (T == 0) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: F=%d D=%d", who, F, D);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(rows && tile_base && tile_cnt && pair_sorted && cat_pos, "%s: null pointer", who);
    RBX_REQUIRE(!d_fm || (S && (E || rowbuf)), "%s: d_fm needs S and E (or the row buffer of the forward)", who);
    GradParams p;
    p.E = E; p.rowbuf = rowbuf; p.S = S; p.dE = dE; p.d_fm = d_fm; p.d_lr = d_lr; p.rows = rows;
    for (int f = 0; f < kMaxF; ++f) p.pad_row[f] = (pad_row && f < F) ? pad_row[f] : -1;
    p.tile_base = tile_base; p.tile_cnt = tile_cnt; p.pair_sorted = pair_sorted;
    p.B = B; p.cap = cap; p.F = F; p.Ft = n_slots; p.T = T; p.world = world; p.rank = rank;
    p.f_magic = ((1u << 20) + (uint32_t)F - 1u) / (uint32_t)F;
    if (int rc = fill_peers(p.ginbox, ginbox, world, true, who, "ginbox")) return rc;
    if (int rc = fill_peers(p.ginbox_lr, ginbox_lr, world, false, who, "ginbox_lr")) return rc;
    for (int f = 0; f < F; ++f) {
        RBX_REQUIRE(cat_pos[f] >= 0 && cat_pos[f] < n_slots, "%s: cat_pos[%d]", who, f);
        p.cat_pos[f] = (int16_t)cat_pos[f];
    }
    const int64_t n_tiles = (B + T - 1) / T;
    int64_t grid = (int64_t)rbx_sm_count() * 6;
    if (grid > n_tiles) grid = n_tiles;
    XS_DISPATCH_LPR(D, (k_xs_grad_push<LPR><<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_apply(const float* ginbox, const float* ginbox_lr, const int32_t* inbox_ids, const int32_t* meta, int64_t cap,
                 int world, float* g_table, int64_t row_stride, float* g_lr, int64_t lr_stride, int D, rbx_stream_t stream) {
    const char* who = "rbx_xs_apply";
    RBX_REQUIRE(world >= 1 && world <= kW && cap >= 1, "%s: bad size", who);
    RBX_REQUIRE(ginbox && inbox_ids && meta && g_table && row_stride >= D && row_stride % 4 == 0, "%s: null pointer / bad row stride", who);
    if (!pow2_dim(D)) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: D=%d (covers 4..128, powers of two)", who, D);
    ApplyParams p;
    p.ginbox = ginbox; p.ginbox_lr = ginbox_lr; p.inbox_ids = inbox_ids; p.meta = meta; p.g_table = g_table;
    p.rs = row_stride; p.g_lr = g_lr; p.lr_stride = lr_stride; p.cap = cap;
    int64_t per = (int64_t)rbx_sm_count() * 6 / world;
    if (per < 1) per = 1;
    XS_DISPATCH_LPR(D, (k_xs_apply<LPR><<<dim3((unsigned)per, world), kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
