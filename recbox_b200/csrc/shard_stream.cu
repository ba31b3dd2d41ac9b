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
#include <stdlib.h>
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
    __shared__ int32_t s_base[kW], s_tot[kW];
    __shared__ int32_t s_ids[kMaxPairs];
    __shared__ uint16_t s_pair[kMaxPairs];
    constexpr int kMyChunks = kMaxChunks / kWarps;
    const int W = 1 << p.wlog2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int F = p.F, T = p.T;
    const int64_t n_tiles = (p.B + T - 1) / T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * T;
        const int th = (int)(p.B - b0 < T ? p.B - b0 : T);
        const int np = th * F, nch = (np + 31) >> 5;
        const int32_t* rt = p.rows + b0 * F;
        // one pass over the ids: owner, rank among the chunk's ids of that owner (kept in registers), chunk counts
        int32_t my_r[kMyChunks];
        int my_o[kMyChunks], my_rank[kMyChunks];
#pragma unroll
        for (int ci = 0; ci < kMyChunks; ++ci) {
            const int c = warp + ci * kWarps;
            my_o[ci] = -1;
            my_r[ci] = 0;
            my_rank[ci] = 0;
            if (c < nch) {
                const int k = c * 32 + lane;
                const int32_t r = k < np ? __ldg(rt + k) : -1;
                const int o = ((uint32_t)r < (uint64_t)p.R) ? (r & (W - 1)) : -1;
                int rank_in = 0;
                for (int oo = 0; oo < W; ++oo) {
                    const unsigned m = __ballot_sync(0xffffffffu, o == oo);
                    if (lane == oo) s_pre[c][oo] = __popc(m);
                    if (o == oo) rank_in = __popc(m & ((1u << lane) - 1u));
                }
                my_o[ci] = o;
                my_r[ci] = r;
                my_rank[ci] = rank_in;
            }
        }
        __syncthreads();
        if (warp < W) {                            // warp o scans the chunk counts of owner o (nch <= 64: two per lane)
            const int o = warp;
            const int v0 = lane < nch ? s_pre[lane][o] : 0, v1 = lane + 32 < nch ? s_pre[lane + 32][o] : 0;
            int i0 = v0, i1 = v1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t0 = __shfl_up_sync(0xffffffffu, i0, d), t1 = __shfl_up_sync(0xffffffffu, i1, d);
                if (lane >= d) { i0 += t0; i1 += t1; }
            }
            const int tot0 = __shfl_sync(0xffffffffu, i0, 31), tot = tot0 + __shfl_sync(0xffffffffu, i1, 31);
            if (lane < nch) s_pre[lane][o] = i0 - v0;
            if (lane + 32 < nch) s_pre[lane + 32][o] = tot0 + i1 - v1;
            if (lane == 0) {
                const int base = atomicAdd(p.cursor + o, tot);
                s_base[o] = base;
                s_tot[o] = tot;
                p.tile_base[tile * kW + o] = base;
                p.tile_cnt[tile * kW + o] = tot;
                if ((int64_t)base + tot > p.cap) *p.overflow = 1;
            }
        }
        __syncthreads();
        int off[kW + 1];
        off[0] = 0;
#pragma unroll
        for (int o = 0; o < kW; ++o) off[o + 1] = off[o] + (o < W ? s_tot[o] : 0);
#pragma unroll
        for (int ci = 0; ci < kMyChunks; ++ci) {
            const int o = my_o[ci];
            if (o >= 0) {
                const int c = warp + ci * kWarps;
                int oo = 0;
#pragma unroll
                for (int q = 0; q < kW; ++q) oo += (q == o) ? off[q] : 0;
                const int pos = oo + s_pre[c][o] + my_rank[ci];
                s_ids[pos] = my_r[ci] >> p.wlog2;
                s_pair[pos] = (uint16_t)(c * 32 + lane);
            }
        }
        __syncthreads();
        const int nv = off[kW];
        uint16_t* ps = p.pair_sorted + tile * (int64_t)(T * F);
        for (int i = threadIdx.x; i < nv; i += kThreads) {
            int o = 0;
#pragma unroll
            for (int q = 1; q < kW; ++q) o += (i >= off[q]) ? 1 : 0;
            const int64_t slot = (int64_t)s_base[o] + (i - off[o]);
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
    // destination = blockIdx.x % world, so the CTAs resident at any moment feed every peer's ingress evenly (with the
    // destination on grid.y the last destinations only start when the first ones are done, and their ingress then takes
    // all eight sources at once: measured 316 us instead of ~190 for the same bytes at 8 GPUs)
    const int q = (blockIdx.x + p.rank) % p.world;
    const int bx = blockIdx.x / p.world, gx = gridDim.x / p.world;
    int64_t n = __ldg(p.meta + q);
    if (n > p.cap) n = p.cap;
    float* out = reinterpret_cast<float*>(p.rowbuf.p[q]) + (size_t)p.rank * p.cap * D;
    float* out_lr = (p.lr && p.rowbuf_lr.p[q]) ? reinterpret_cast<float*>(p.rowbuf_lr.p[q]) + (size_t)p.rank * p.cap : nullptr;
    const int32_t* ids = p.inbox_ids + (size_t)q * p.cap;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)bx * kWarps + (threadIdx.x >> 5), nwarps = (int64_t)gx * kWarps;
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
    constexpr int D = 4 * LPR, SG = 32 / LPR, U = 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int32_t s_base[kW], s_off[kW + 1], s_nv;
    const int F = p.F, Fn = p.Fn, Ft = p.Ft, T = p.T, W = p.world;
    float* Es = reinterpret_cast<float*>(smem_raw);                 // [T*F, D]
    float* lrs = Es + (size_t)T * F * D;                            // [T*F]
    int32_t* s_src = reinterpret_cast<int32_t*>(lrs + (size_t)T * F);    // [T*F]  row-buffer row of each pair | -1
    uint16_t* s_ps = reinterpret_cast<uint16_t*>(s_src + (size_t)T * F);  // [T*F]  b_local * F + f of each pair
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
            for (int o = 0; o < kW; ++o) {
                if (o < W) {
                    s_base[o] = p.tile_base[tile * kW + o];
                    run += p.tile_cnt[tile * kW + o];
                }
                s_off[o + 1] = o + 1 < W ? run : 0x7fffffff;     // runs of absent owners start "never"
                if (o + 1 == W) s_nv = run;
            }
        }
        __syncthreads();
        const int nv = s_nv;
        if (nv < np) {                                               // ids outside the table read as zero rows
            for (int i = threadIdx.x; i < np * LPR; i += kThreads) reinterpret_cast<float4*>(Es)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = threadIdx.x; i < np; i += kThreads) lrs[i] = 0.f;
            __syncthreads();
        }
        // per pair, once: where its row sits in the row buffer and where it goes in the E tile; the 16-byte items
        // below then cost a handful of instructions each
        const uint16_t* ps = p.pair_sorted + tile * (int64_t)(T * F);
        for (int i = threadIdx.x; i < nv; i += kThreads) {
            int o = 0;
#pragma unroll
            for (int q = 1; q < kW; ++q) o += (i >= s_off[q]) ? 1 : 0;
            const int64_t slot = (int64_t)s_base[o] + (i - s_off[o]);
            s_src[i] = slot < p.cap ? (int32_t)(o * p.cap + slot) : -1;
            s_ps[i] = ps[i];
        }
        __syncthreads();
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
                    const int32_t src = s_src[pp];
                    if (src >= 0) {
                        k[u] = s_ps[pp];
                        v[u] = ld_stream_f4(p.rowbuf + (size_t)src * D + 4 * cc);
                        if (cc == 0 && p.rowbuf_lr) l[u] = ld_stream_f1(p.rowbuf_lr + src);
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
    // numeric slots / bias: batch reductions folded into this launch (no exchange), see rbx_embed_fm_bwd
    const float* dense_x;
    const float* dense_w;
    float* g_dense_w;
    float* g_dense_w_lr;
    float* g_lr_bias;
    int64_t B, cap;
    int F, Fn, Ft, T, world, rank;
    uint32_t f_magic;          // ceil(2^20 / F): k / F == (k * f_magic) >> 20 for k < 2048
    Peers ginbox;              // owner w's: float [kW(requester), cap, D]
    Peers ginbox_lr;           // owner w's: float [kW(requester), cap]
    int16_t cat_pos[kMaxF];
    int32_t pad_row[kMaxF];
    int16_t num_pos[kMaxF];
    int16_t num_widx[kMaxF];
};

template <int LPR, int U, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_xs_grad_push(const __grid_constant__ GradParams p) {
    constexpr int D = 4 * LPR;
    __shared__ int32_t s_base[kW], s_off[kW + 1], s_nv;
    // per pair of the tile, filled once: E / dE offset (tile-relative, -1 = padding row: zero gradient), slot in the
    // owner's inbox (-1 = beyond capacity), owner, sample
    __shared__ int32_t s_eoff[kMaxPairs], s_slot[kMaxPairs];
    __shared__ uint8_t s_own[kMaxPairs], s_bl[kMaxPairs];
    __shared__ __align__(16) float s_S[2048];   // S of the tile's samples: T * D <= 2048 floats
    __shared__ float s_dfm[32], s_dlr[32];
    __shared__ float4 s_acc[kThreads];
    __shared__ float s_lr[kMaxF + 1];
    const int F = p.F, Fn = p.Fn, Ft = p.Ft, T = p.T, W = p.world;
    const bool has_fm = p.d_fm != nullptr;
    const uint64_t pol_stream = l2_policy_evict_first();
    const int64_t n_tiles = (p.B + T - 1) / T;
    // numeric roles: thread <-> (sample lane sub, slot n, float4 column c), accumulated over all tiles of this CTA
    const int R4 = Fn * LPR, SUB = R4 > 0 ? kThreads / R4 : 0;
    const bool num_on = R4 > 0 && SUB > 0 && p.g_dense_w && (p.dE || has_fm);
    const int nsub = (int)threadIdx.x / (R4 > 0 ? R4 : 1), nrole = (int)threadIdx.x - nsub * (R4 > 0 ? R4 : 1);
    const int nn = nrole / LPR, nc = nrole - nn * LPR;
    float4 nacc = make_float4(0.f, 0.f, 0.f, 0.f), nw = nacc;
    if (num_on && nsub < SUB && p.dense_w) nw = ld_row_f4(p.dense_w + (size_t)p.num_widx[nn] * D + 4 * nc);
    const bool lr_on = p.d_lr && (p.g_dense_w_lr || p.g_lr_bias) && Fn + 1 <= kThreads;
    const int lroles = Fn + 1, llanes = kThreads / lroles;
    const int lln = (int)threadIdx.x / lroles, lj = (int)threadIdx.x - lln * lroles;
    float lacc = 0.f;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b0 = tile * T;
        const int th = (int)(p.B - b0 < T ? p.B - b0 : T);
        if (threadIdx.x == 0) {
            int run = 0;
            s_off[0] = 0;
            for (int o = 0; o < kW; ++o) {
                if (o < W) {
                    s_base[o] = p.tile_base[tile * kW + o];
                    run += p.tile_cnt[tile * kW + o];
                }
                s_off[o + 1] = o + 1 < W ? run : 0x7fffffff;
                if (o + 1 == W) s_nv = run;
            }
        }
        if (has_fm)
            for (int i = threadIdx.x; i < th * LPR; i += kThreads)
                reinterpret_cast<float4*>(s_S)[i] = ld_stream_f4(p.S + (size_t)b0 * D + 4 * i);
        if ((int)threadIdx.x < th) {
            s_dfm[threadIdx.x] = has_fm ? __ldg(p.d_fm + b0 + threadIdx.x) : 0.f;
            s_dlr[threadIdx.x] = p.d_lr ? __ldg(p.d_lr + b0 + threadIdx.x) : 0.f;
        }
        __syncthreads();
        const int nv = s_nv;
        const uint16_t* ps = p.pair_sorted + tile * (int64_t)(T * F);
        for (int i = threadIdx.x; i < nv; i += kThreads) {
            const uint32_t k = ps[i];
            const uint32_t bl = (k * p.f_magic) >> 20;
            const int f = (int)(k - bl * F);
            const bool pad = __ldg(p.rows + b0 * F + k) == p.pad_row[f];
            int o = 0;
#pragma unroll
            for (int q = 1; q < kW; ++q) o += (i >= s_off[q]) ? 1 : 0;
            const int64_t slot = (int64_t)s_base[o] + (i - s_off[o]);
            s_eoff[i] = pad ? -1 : (int32_t)((bl * Ft + p.cat_pos[f]) * D);
            s_slot[i] = slot < p.cap ? (int32_t)slot : -1;
            s_own[i] = (uint8_t)o;
            s_bl[i] = (uint8_t)bl;
        }
        __syncthreads();
        const int items = nv * LPR;
        const float* dEt = p.dE ? p.dE + (size_t)b0 * Ft * D : nullptr;
        const float* Et = p.E ? p.E + (size_t)b0 * Ft * D : nullptr;
        for (int i0 = threadIdx.x; i0 < items; i0 += kThreads * U) {
            float4 g[U], e[U];
            int pp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * kThreads;
                pp[u] = -1;
                g[u] = e[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < items) {
                    const int q = i / LPR, cc = i & (LPR - 1);
                    if (s_slot[q] >= 0) {
                        pp[u] = q;
                        const int32_t eo = s_eoff[q];
                        if (eo >= 0) {
                            if (dEt) g[u] = ld_stream_f4_hint(dEt + eo + 4 * cc, pol_stream);
                            if (has_fm)
                                e[u] = Et ? ld_stream_f4_hint(Et + eo + 4 * cc, pol_stream)
                                          : ld_stream_f4(p.rowbuf + ((size_t)s_own[q] * p.cap + s_slot[q]) * D + 4 * cc);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = pp[u];
                if (q >= 0) {
                    const int cc = (i0 + u * kThreads) & (LPR - 1);
                    const int o = s_own[q], bl = s_bl[q];
                    const bool live = s_eoff[q] >= 0;
                    if (has_fm && live)
                        g[u] = f4_fma(f4_sub(*reinterpret_cast<const float4*>(s_S + bl * D + 4 * cc), e[u]), s_dfm[bl], g[u]);
                    const size_t in_off = (size_t)p.rank * p.cap + s_slot[q];
                    st_stream_f4(reinterpret_cast<float*>(p.ginbox.p[o]) + in_off * D + 4 * cc, g[u]);
                    if (cc == 0 && p.d_lr && p.ginbox_lr.p[o])
                        reinterpret_cast<float*>(p.ginbox_lr.p[o])[in_off] = live ? s_dlr[bl] : 0.f;
                }
            }
        }
        // numeric slots of this tile's samples: g_dense_w[n,:] += x (dE + d_fm (S - x w)); first-order / bias roles
        if (num_on && nsub < SUB) {
            for (int bl = nsub; bl < th; bl += SUB) {
                const int64_t b = b0 + bl;
                const float x = __ldg(p.dense_x + b * Fn + nn);
                float4 g = p.dE ? ld_stream_f4_hint(p.dE + ((size_t)b * Ft + p.num_pos[nn]) * D + 4 * nc, pol_stream)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
                if (has_fm) g = f4_fma(f4_sub(*reinterpret_cast<const float4*>(s_S + bl * D + 4 * nc), f4_scale(nw, x)), s_dfm[bl], g);
                nacc = f4_fma(g, x, nacc);
            }
        }
        if (lr_on && lln < llanes) {
            for (int bl = lln; bl < th; bl += llanes)
                lacc = fmaf(lj < Fn ? __ldg(p.dense_x + (b0 + bl) * Fn + lj) : 1.f, s_dlr[bl], lacc);
        }
        __syncthreads();
    }
    if (num_on) {
        s_acc[threadIdx.x] = nacc;
        __syncthreads();
        if ((int)threadIdx.x < R4) {
            float4 a = s_acc[threadIdx.x];
            for (int sub = 1; sub < SUB; ++sub) a = f4_add(a, s_acc[sub * R4 + threadIdx.x]);
            red_add_f4(p.g_dense_w + (size_t)p.num_widx[nn] * D + 4 * nc, a);
        }
    }
    if (lr_on) {
        for (int i = threadIdx.x; i <= Fn; i += kThreads) s_lr[i] = 0.f;
        __syncthreads();
        if (lln < llanes) atomicAdd(&s_lr[lj], lacc);
        __syncthreads();
        if ((int)threadIdx.x < Fn && p.g_dense_w_lr) red_add_f1(p.g_dense_w_lr + p.num_widx[threadIdx.x], s_lr[threadIdx.x]);
        if ((int)threadIdx.x == Fn && p.g_lr_bias) red_add_f1(p.g_lr_bias, s_lr[Fn]);
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

size_t consume_smem(int T, int F, int D) { return (size_t)T * F * (D + 2) * 4 + (((size_t)T * F * 2 + 15) & ~(size_t)15); }

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
    while (T > 1 && (consume_smem(T, F, D) > 62 * 1024 || T * F > kMaxPairs)) T >>= 1;
    return (consume_smem(T, F, D) <= 200 * 1024 && T * F <= kMaxPairs) ? T : 0;
}

int rbx_xs_route(const int32_t* rows, int64_t B, int F, int64_t R, int D, int rank, int world, int64_t cap, int32_t* cursor, int32_t* tile_base, int32_t* tile_cnt, uint16_t* pair_sorted,
                 int32_t* overflow, void* const* inbox_ids, rbx_stream_t stream) {
    const char* who = "rbx_xs_route";
    RBX_RANGE(who);
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
    int64_t grid = (int64_t)rbx_sm_count() * 8;
    if (grid > n_tiles) grid = n_tiles;
    k_xs_route<<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_barrier(void* const* flags, void* const* meta, int32_t* cursor, int rank, int world, uint32_t epoch,
                   rbx_stream_t stream) {
    const char* who = "rbx_xs_barrier";
    RBX_RANGE(who);
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
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= kW && rank >= 0 && rank < world && cap >= 1, "%s: bad size", who);
    RBX_REQUIRE(table && inbox_ids && meta && row_stride >= D && row_stride % 4 == 0, "%s: null pointer / bad row stride", who);
    if (!pow2_dim(D)) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: D=%d (covers 4..128, powers of two)", who, D);
    ServeParams p;
    p.table = table; p.rs = row_stride; p.lr = lr; p.lr_stride = lr_stride; p.inbox_ids = inbox_ids; p.meta = meta;
    p.cap = cap; p.rank = rank; p.world = world;
    if (int rc = fill_peers(p.rowbuf, rowbuf, world, true, who, "rowbuf")) return rc;
    if (int rc = fill_peers(p.rowbuf_lr, rowbuf_lr, world, lr != nullptr, who, "rowbuf_lr")) return rc;
    int64_t per = (int64_t)rbx_sm_count() * 4 / world;      // 4 resident CTAs per SM (64 registers): one wave, every destination in it
    if (per < 1) per = 1;
    XS_DISPATCH_LPR(D, (k_xs_serve<LPR><<<(unsigned)(per * world), kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_consume(const float* rowbuf, const float* rowbuf_lr, const int32_t* tile_base, const int32_t* tile_cnt,
                   const uint16_t* pair_sorted, const int32_t* cat_pos, const float* dense_x, const float* dense_w,
                   const float* dense_w_lr, const int32_t* num_pos, const int32_t* num_widx, const float* lr_bias,
                   float* E, float* S, float* fm_out, float* lr_out, int64_t B, int64_t cap, int F, int Fn, int D,
                   int n_slots, int world, rbx_stream_t stream) {
    const char* who = "rbx_xs_consume";
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= kW && cap >= 1 && B >= 0 && (int64_t)world * cap < INT32_MAX, "%s: bad size", who);
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
                     const float* d_lr, const int32_t* rows, const int32_t* pad_row, const int32_t* tile_base,
                     const int32_t* tile_cnt, const uint16_t* pair_sorted, const int32_t* cat_pos,
                     const float* dense_x, const float* dense_w, const int32_t* num_pos, const int32_t* num_widx, int Fn,
                     float* g_dense_w, float* g_dense_w_lr, float* g_lr_bias,
                     int64_t B, int64_t cap, int F, int D, int n_slots, int rank, int world,
                     void* const* ginbox, void* const* ginbox_lr, rbx_stream_t stream) {
    const char* who = "rbx_xs_grad_push";
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= kW && rank >= 0 && rank < world && cap >= 1 && B >= 0, "%s: bad size", who);
    const int T = rbx_xs_tile_samples(F, D);
    if (T == 0) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: F=%d D=%d", who, F, D);
    if (B == 0) return RBX_OK;
    RBX_REQUIRE(rows && tile_base && tile_cnt && pair_sorted && cat_pos, "%s: null pointer", who);
    RBX_REQUIRE(!d_fm || (S && (E || rowbuf)), "%s: d_fm needs S and E (or the row buffer of the forward)", who);
    RBX_REQUIRE(Fn >= 0 && Fn <= kMaxF && Fn * (D / 4) <= kThreads, "%s: %d numeric slots of D=%d exceed the fused reduction (use rbx_embed_fm_bwd with F = 0)", who, Fn, D);
    RBX_REQUIRE(Fn == 0 || !(g_dense_w || g_dense_w_lr) || (dense_x && num_pos), "%s: numeric slots need dense_x / num_pos", who);
    RBX_REQUIRE(Fn == 0 || !g_dense_w || !d_fm || dense_w, "%s: g_dense_w with d_fm needs dense_w", who);
    GradParams p;
    p.E = E; p.rowbuf = rowbuf; p.S = S; p.dE = dE; p.d_fm = d_fm; p.d_lr = d_lr; p.rows = rows;
    for (int f = 0; f < kMaxF; ++f) p.pad_row[f] = (pad_row && f < F) ? pad_row[f] : -1;
    p.tile_base = tile_base; p.tile_cnt = tile_cnt; p.pair_sorted = pair_sorted;
    p.dense_x = dense_x; p.dense_w = dense_w; p.g_dense_w = Fn ? g_dense_w : nullptr;
    p.g_dense_w_lr = Fn ? g_dense_w_lr : nullptr; p.g_lr_bias = g_lr_bias;
    p.B = B; p.cap = cap; p.F = F; p.Fn = Fn; p.Ft = n_slots; p.T = T; p.world = world; p.rank = rank;
    p.f_magic = ((1u << 20) + (uint32_t)F - 1u) / (uint32_t)F;
    if (int rc = fill_peers(p.ginbox, ginbox, world, true, who, "ginbox")) return rc;
    if (int rc = fill_peers(p.ginbox_lr, ginbox_lr, world, false, who, "ginbox_lr")) return rc;
    for (int f = 0; f < F; ++f) {
        RBX_REQUIRE(cat_pos[f] >= 0 && cat_pos[f] < n_slots, "%s: cat_pos[%d]", who, f);
        p.cat_pos[f] = (int16_t)cat_pos[f];
    }
    for (int n = 0; n < kMaxF; ++n) {
        p.num_pos[n] = 0;
        p.num_widx[n] = 0;
        if (n < Fn) {
            RBX_REQUIRE(num_pos[n] >= 0 && num_pos[n] < n_slots, "%s: num_pos[%d]", who, n);
            p.num_pos[n] = (int16_t)num_pos[n];
            p.num_widx[n] = (int16_t)(num_widx ? num_widx[n] : n);
        }
    }
    const int64_t n_tiles = (B + T - 1) / T;
    static const int variant = getenv("RBX_XS_GP_VARIANT") ? atoi(getenv("RBX_XS_GP_VARIANT")) : 1;   // tuning knob
    const int minb = variant == 1 ? 4 : (variant == 2 ? 3 : (variant == 3 ? 6 : (variant == 4 ? 5 : 2)));
    int64_t grid = (int64_t)rbx_sm_count() * minb;
    if (grid > n_tiles) grid = n_tiles;
    if (variant == 1) {
        XS_DISPATCH_LPR(D, (k_xs_grad_push<LPR, 2, 4><<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    } else if (variant == 2) {
        XS_DISPATCH_LPR(D, (k_xs_grad_push<LPR, 4, 3><<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    } else if (variant == 3) {
        XS_DISPATCH_LPR(D, (k_xs_grad_push<LPR, 2, 6><<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    } else if (variant == 4) {
        XS_DISPATCH_LPR(D, (k_xs_grad_push<LPR, 3, 5><<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    } else {
        XS_DISPATCH_LPR(D, (k_xs_grad_push<LPR, 4, 2><<<(int)grid, kThreads, 0, rbx_cast_stream(stream)>>>(p)));
    }
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_xs_apply(const float* ginbox, const float* ginbox_lr, const int32_t* inbox_ids, const int32_t* meta, int64_t cap,
                 int world, float* g_table, int64_t row_stride, float* g_lr, int64_t lr_stride, int D, rbx_stream_t stream) {
    const char* who = "rbx_xs_apply";
    RBX_RANGE(who);
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
