// (e) push-based exchange for the row-sharded table: the all-to-all of SURVEY.md section 8e done by
// the gather / scatter kernels themselves with stores over NVLink into the peers' inboxes
// (CUDA-IPC mapped, see peer.cu) -- no NCCL on the data path, no host-side split sizes.
//
// Why push and not remote gathers: random 64-byte reads into a multi-GB peer-mapped table collapse
// to ~10 GB/s on NVSwitch (the peer translation reach is tiny; measured, profiles/), while
// contiguous streams into a small inbox run at link speed.  So every random access stays LOCAL to
// the owner of the row and only dense, contiguous buffers cross NVLink:
//
//   requester q                                   owner w
//   rbx_shard_route: ids bucketed by owner
//   rbx_shard_push_ids  -- ids, (count, offset) -->  inbox_ids[q], inbox_meta[q]
//                                 ---- barrier ----
//                       <-- rows, lr values -------  rbx_shard_serve_rows: local gather, stored
//                                                    straight into q's row buffer (send order)
//                                 ---- barrier ----
//   rbx_embed_fm_fwd(table = row buffer, rows = pos)         (un-permute folded into the FM kernel)
//   rbx_embed_fm_bwd(g_table = grad send buffer, rows = pos)
//   rbx_shard_push_grads -- row grads, lr grads -->  ginbox[q]
//                                 ---- barrier ----
//                                                    rbx_shard_apply_grads: local scatter-add
//
// Slot capacity `cap` ids per (owner, requester) pair is fixed by the caller; a bucket that would
// overflow is truncated and flagged in inbox_meta[q][2] (the host checks the flag).
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMeta = 4;   // int32 per (owner, requester): count, offset in the requester's send order, overflow flag, pad

__device__ __forceinline__ void st_f4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// bucket start offsets from counts (world <= 8): start[w] = sum_{v<w} counts[v]
__device__ __forceinline__ void bucket_of(const int32_t* __restrict__ counts, int world, int w, int& start, int& cnt) {
    int s = 0;
    for (int v = 0; v < w; ++v) s += __ldg(counts + v);
    start = s;
    cnt = __ldg(counts + w);
}

struct PeerPtrs {
    void* p[RBX_MAX_WORLD];
    void* q[RBX_MAX_WORLD];
};

// grid.y = owner w.  inbox_ids[w][rank*cap + j] = send[start_w + j]
__global__ void __launch_bounds__(kThreads) k_push_ids(const int32_t* __restrict__ send, const int32_t* __restrict__ counts,
                                                      const __grid_constant__ PeerPtrs peers, int rank, int world, int64_t cap) {
    const int w = blockIdx.y;
    int start, cnt;
    bucket_of(counts, world, w, start, cnt);
    int32_t* ids = reinterpret_cast<int32_t*>(peers.p[w]) + (size_t)rank * cap;
    int32_t* meta = reinterpret_cast<int32_t*>(peers.q[w]) + rank * kMeta;
    const int n = cnt < cap ? cnt : (int)cap;
    for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += (int64_t)gridDim.x * kThreads)
        ids[j] = __ldg(send + start + j);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        meta[0] = n;
        meta[1] = start;
        meta[2] = cnt > cap ? 1 : 0;
    }
}

// owner: grid.y = requester q.  rows[q][(off_q + j) * D ..] = table[inbox_ids[q*cap + j]]
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_serve_rows(const float* __restrict__ table, const float* __restrict__ table_lr,
                                                        const int32_t* __restrict__ inbox_ids, const int32_t* __restrict__ inbox_meta,
                                                        const __grid_constant__ PeerPtrs outs, int64_t cap) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR, U = 4;
    const int q = blockIdx.y;
    const int n = __ldg(inbox_meta + q * kMeta), off = __ldg(inbox_meta + q * kMeta + 1);
    float* out = reinterpret_cast<float*>(outs.p[q]) + (size_t)off * D;
    float* out_lr = outs.q[q] ? reinterpret_cast<float*>(outs.q[q]) + off : nullptr;
    const int32_t* ids = inbox_ids + (size_t)q * cap;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
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
            if (r[u] >= 0) {
                v[u] = ld_row_f4(table + (size_t)r[u] * D + 4 * lig);
                if (out_lr && lig == 0) l[u] = __ldg(table_lr + r[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            if (r[u] >= 0) {
                st_f4(out + (size_t)i * D + 4 * lig, v[u]);
                if (out_lr && lig == 0) out_lr[i] = l[u];
            }
        }
    }
}

// requester: grid.y = owner w.  ginbox[w][(rank*cap + j) * D ..] = gsend[(start_w + j) * D ..]
__global__ void __launch_bounds__(kThreads) k_push_grads(const float* __restrict__ gsend, const float* __restrict__ gsend_lr,
                                                        const int32_t* __restrict__ counts, const __grid_constant__ PeerPtrs peers,
                                                        int rank, int world, int64_t cap, int D) {
    const int w = blockIdx.y;
    int start, cnt;
    bucket_of(counts, world, w, start, cnt);
    const int n = cnt < cap ? cnt : (int)cap;
    const int V = D / 4;
    const float4* src = reinterpret_cast<const float4*>(gsend + (size_t)start * D);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(peers.p[w]) + (size_t)rank * cap * D);
    const int64_t total = (int64_t)n * V;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads)
        dst[t] = __ldg(src + t);
    if (gsend_lr && peers.q[w]) {
        float* dl = reinterpret_cast<float*>(peers.q[w]) + (size_t)rank * cap;
        for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += (int64_t)gridDim.x * kThreads)
            dl[j] = __ldg(gsend_lr + start + j);
    }
}

// owner: grid.y = requester q.  g_table[inbox_ids[q][j]] += ginbox[q][j]
template <int LPR>
__global__ void __launch_bounds__(kThreads) k_apply_grads(const float* __restrict__ ginbox, const float* __restrict__ ginbox_lr,
                                                         const int32_t* __restrict__ inbox_ids, const int32_t* __restrict__ inbox_meta,
                                                         float* __restrict__ g_table, float* __restrict__ g_table_lr, int64_t cap) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR, U = 4;
    const int q = blockIdx.y;
    const int n = __ldg(inbox_meta + q * kMeta);
    const int32_t* ids = inbox_ids + (size_t)q * cap;
    const float* g = ginbox + (size_t)q * cap * D;
    const float* gl = ginbox_lr ? ginbox_lr + (size_t)q * cap : nullptr;
    const int lane = threadIdx.x & 31, lig = lane & (LPR - 1), gi = lane / LPR;
    const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
    for (int64_t base = warp0 * (RPW * U); base < n; base += nwarps * (RPW * U)) {
        int32_t r[U];
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            r[u] = i < n ? __ldg(ids + i) : -1;
            if (r[u] >= 0) v[u] = ld_stream_f4(g + (size_t)i * D + 4 * lig);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * RPW + gi;
            if (r[u] >= 0) {
                red_add_f4(g_table + (size_t)r[u] * D + 4 * lig, v[u]);
                if (gl && g_table_lr && lig == 0) red_add_f1(g_table_lr + r[u], ld_stream_f1(gl + i));
            }
        }
    }
}

__global__ void k_zero_rows(float* __restrict__ g_table, float* __restrict__ g_table_lr, const int32_t* __restrict__ rows, int n, int D) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const int32_t r = rows[i];
    for (int d = threadIdx.x; d < D; d += blockDim.x) g_table[(size_t)r * D + d] = 0.f;
    if (g_table_lr && threadIdx.x == 0) g_table_lr[r] = 0.f;
}

inline bool pow2_dim(int D) { return D % 4 == 0 && D <= 128 && (D & (D - 1)) == 0; }

int fill_peers(PeerPtrs& pp, void* const* a, void* const* b, int world, const char* who) {
    for (int w = 0; w < RBX_MAX_WORLD; ++w) pp.p[w] = pp.q[w] = nullptr;
    for (int w = 0; w < world; ++w) {
        if (!a || !a[w]) return rbx_fail(RBX_ERR_ARG, "%s: peer pointer %d is null", who, w);
        pp.p[w] = a[w];
        pp.q[w] = b ? b[w] : nullptr;
    }
    return RBX_OK;
}

#define RBX_DISPATCH_LPR(D, CALL)                   \
    switch ((D) / 4) {                              \
        case 1: { constexpr int LPR = 1; CALL; } break;   \
        case 2: { constexpr int LPR = 2; CALL; } break;   \
        case 4: { constexpr int LPR = 4; CALL; } break;   \
        case 8: { constexpr int LPR = 8; CALL; } break;   \
        case 16: { constexpr int LPR = 16; CALL; } break; \
        default: { constexpr int LPR = 32; CALL; } break; \
    }

}  // namespace

extern "C" {

int rbx_shard_push_ids(const int32_t* send, const int32_t* counts, int32_t* const* inbox_ids, int32_t* const* inbox_meta,
                       int rank, int world, int64_t cap, int64_t N, rbx_stream_t stream) {
    const char* who = "rbx_shard_push_ids";
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= RBX_MAX_WORLD && rank >= 0 && rank < world, "%s: rank/world", who);
    RBX_REQUIRE(cap >= 1 && N >= 0 && counts, "%s: bad size", who);
    RBX_REQUIRE(N == 0 || send, "%s: null send", who);
    PeerPtrs pp;
    if (int rc = fill_peers(pp, (void* const*)inbox_ids, (void* const*)inbox_meta, world, who)) return rc;
    RBX_REQUIRE(inbox_meta != nullptr, "%s: inbox_meta required", who);
    int64_t per = (N / world + kThreads - 1) / kThreads + 1;
    const int64_t capx = (int64_t)rbx_sm_count() * 2;
    if (per > capx) per = capx;
    k_push_ids<<<dim3((unsigned)per, world), kThreads, 0, rbx_cast_stream(stream)>>>(send, counts, pp, rank, world, cap);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_shard_serve_rows(const float* table, const float* table_lr, int D, const int32_t* inbox_ids,
                         const int32_t* inbox_meta, float* const* out_rows, float* const* out_lr, int world, int64_t cap,
                         rbx_stream_t stream) {
    const char* who = "rbx_shard_serve_rows";
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= RBX_MAX_WORLD && cap >= 1, "%s: bad size", who);
    RBX_REQUIRE(table && inbox_ids && inbox_meta, "%s: null pointer", who);
    RBX_REQUIRE(!out_lr || table_lr, "%s: table_lr required with out_lr", who);
    if (!pow2_dim(D)) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: D=%d (covers 4..128, powers of two)", who, D);
    PeerPtrs pp;
    if (int rc = fill_peers(pp, (void* const*)out_rows, (void* const*)out_lr, world, who)) return rc;
    int64_t per = (int64_t)rbx_sm_count() * 8 / world;
    if (per < 1) per = 1;
    RBX_DISPATCH_LPR(D, (k_serve_rows<LPR><<<dim3((unsigned)per, world), kThreads, 0, rbx_cast_stream(stream)>>>(
                            table, table_lr, inbox_ids, inbox_meta, pp, cap)));
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_shard_push_grads(const float* gsend, const float* gsend_lr, const int32_t* counts, float* const* ginbox,
                         float* const* ginbox_lr, int rank, int world, int64_t cap, int D, int64_t N, rbx_stream_t stream) {
    const char* who = "rbx_shard_push_grads";
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= RBX_MAX_WORLD && rank >= 0 && rank < world, "%s: rank/world", who);
    RBX_REQUIRE(cap >= 1 && N >= 0 && counts && D % 4 == 0, "%s: bad size (D %% 4 == 0 required)", who);
    RBX_REQUIRE(N == 0 || gsend, "%s: null gsend", who);
    PeerPtrs pp;
    if (int rc = fill_peers(pp, (void* const*)ginbox, (void* const*)ginbox_lr, world, who)) return rc;
    int64_t per = (int64_t)rbx_sm_count() * 8 / world;
    if (per < 1) per = 1;
    k_push_grads<<<dim3((unsigned)per, world), kThreads, 0, rbx_cast_stream(stream)>>>(gsend, gsend_lr, counts, pp, rank, world, cap, D);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_shard_apply_grads(const float* ginbox, const float* ginbox_lr, const int32_t* inbox_ids, const int32_t* inbox_meta,
                          float* g_table, float* g_table_lr, int world, int64_t cap, int D, const int32_t* pad_local,
                          int n_pad, rbx_stream_t stream) {
    const char* who = "rbx_shard_apply_grads";
    RBX_RANGE(who);
    RBX_REQUIRE(world >= 1 && world <= RBX_MAX_WORLD && cap >= 1, "%s: bad size", who);
    RBX_REQUIRE(ginbox && inbox_ids && inbox_meta && g_table, "%s: null pointer", who);
    if (!pow2_dim(D)) return rbx_fail(RBX_ERR_UNSUPPORTED, "%s: D=%d (covers 4..128, powers of two)", who, D);
    int64_t per = (int64_t)rbx_sm_count() * 8 / world;
    if (per < 1) per = 1;
    cudaStream_t st = rbx_cast_stream(stream);
    RBX_DISPATCH_LPR(D, (k_apply_grads<LPR><<<dim3((unsigned)per, world), kThreads, 0, st>>>(ginbox, ginbox_lr, inbox_ids, inbox_meta,
                                                                                        g_table, g_table_lr, cap)));
    RBX_LAUNCH_CHECK(who);
    if (n_pad > 0 && pad_local) {   // nn.Embedding(padding_idx): the padding rows' gradient is defined as zero
        k_zero_rows<<<n_pad, 32, 0, st>>>(g_table, g_table_lr, pad_local, n_pad, D);
        RBX_LAUNCH_CHECK(who);
    }
    return RBX_OK;
}

}  // extern "C"
