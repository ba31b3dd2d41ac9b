// (e) row-shard routing around the all-to-all: stable bucket-by-owner of global row ids, and the
// payload permutations either side of the exchange.  owner(r) = r % world, local row = r / world.
// The partition is STABLE (first-come order inside a bucket), hence deterministic and equal to the
// oracle's argsort(kind="stable"): per-CTA histograms -> one exclusive scan -> ranked placement.
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;                        // sub-tiles per CTA
constexpr int kChunk = kThreads * kItems;        // ids per CTA
constexpr int kMaxWorld = 64;

__global__ void __launch_bounds__(kThreads) k_shard_hist(const int32_t* __restrict__ rows, int64_t N, int world, int nblk,
                                                        int32_t* __restrict__ hist /*[world][nblk]*/) {
    __shared__ int32_t s[kMaxWorld];
    if (threadIdx.x < world) s[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kChunk;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const int64_t i = base + k * kThreads + threadIdx.x;
        if (i < N) atomicAdd(&s[(uint32_t)__ldg(rows + i) % (uint32_t)world], 1);
    }
    __syncthreads();
    if (threadIdx.x < world) hist[(size_t)threadIdx.x * nblk + blockIdx.x] = s[threadIdx.x];
}

// single CTA: exclusive scan of hist (owner-major) in place; counts[w] = bucket size
__global__ void __launch_bounds__(1024) k_shard_scan(int32_t* __restrict__ hist, int world, int nblk, int32_t* __restrict__ counts) {
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_carry;
    const int64_t total = (int64_t)world * nblk;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int64_t base = 0; base < total; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const int32_t x = i < total ? hist[i] : 0;
        int32_t inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int32_t carry = s_carry;
        const int32_t excl = carry + (wid ? s_warp[wid - 1] : 0) + inc - x;
        if (i < total) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    // bucket sizes: start of bucket w+1 minus start of bucket w
    if (threadIdx.x < world) {
        const int32_t start = hist[(size_t)threadIdx.x * nblk];
        const int32_t end = threadIdx.x + 1 < world ? hist[(size_t)(threadIdx.x + 1) * nblk] : s_carry;
        counts[threadIdx.x] = end - start;
    }
}

__global__ void __launch_bounds__(kThreads) k_shard_place(const int32_t* __restrict__ rows, int64_t N, int world, int nblk,
                                                         const int32_t* __restrict__ offs /*[world][nblk]*/,
                                                         int32_t* __restrict__ send, int32_t* __restrict__ pos) {
    __shared__ int32_t s_base[kMaxWorld];
    __shared__ int32_t s_cnt[kThreads / 32][kMaxWorld];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < world) s_base[threadIdx.x] = offs[(size_t)threadIdx.x * nblk + blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kChunk;
    for (int k = 0; k < kItems; ++k) {
        for (int t = threadIdx.x; t < (kThreads / 32) * kMaxWorld; t += kThreads) (&s_cnt[0][0])[t] = 0;
        __syncthreads();
        const int64_t i = base + k * kThreads + threadIdx.x;
        const bool valid = i < N;
        const int32_t r = valid ? __ldg(rows + i) : 0;
        const int owner = valid ? (int)((uint32_t)r % (uint32_t)world) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) s_cnt[wid][owner] = __popc(peers);
        __syncthreads();
        if (valid) {
            int pre = s_base[owner] + rank;
            for (int w = 0; w < wid; ++w) pre += s_cnt[w][owner];
            send[pre] = (int32_t)((uint32_t)r / (uint32_t)world);
            pos[i] = pre;
        }
        __syncthreads();
        if (threadIdx.x < world) {
            int add = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) add += s_cnt[w][threadIdx.x];
            s_base[threadIdx.x] += add;
        }
        __syncthreads();
    }
}

// payload permutations: rows of D floats.  forward: out[pos[i]] = in[i]; inverse: out[i] = in[pos[i]]
template <bool kInverse>
__global__ void __launch_bounds__(kThreads) k_permute_rows(const float* __restrict__ in, const int32_t* __restrict__ pos,
                                                          float* __restrict__ out, int64_t N, int D) {
    const int vec = D / 4;
    const int64_t total = N * vec;
    if (D % 4 == 0) {
        for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
            const int64_t i = t / vec;
            const int c = (int)(t - i * vec);
            const int64_t j = __ldg(pos + i);
            const size_t src = kInverse ? (size_t)j * D + 4 * c : (size_t)i * D + 4 * c;
            const size_t dst = kInverse ? (size_t)i * D + 4 * c : (size_t)j * D + 4 * c;
            st_stream_f4(out + dst, ld_stream_f4(in + src));
        }
    } else {
        const int64_t tot = N * D;
        for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < tot; t += (int64_t)gridDim.x * kThreads) {
            const int64_t i = t / D;
            const int d = (int)(t - i * D);
            const int64_t j = __ldg(pos + i);
            if (kInverse) out[(size_t)i * D + d] = in[(size_t)j * D + d];
            else out[(size_t)j * D + d] = in[(size_t)i * D + d];
        }
    }
}

inline int nblk_for(int64_t N) { return (int)((N + kChunk - 1) / kChunk); }

}  // namespace

extern "C" {

size_t rbx_shard_ws_bytes(int64_t N, int world) {
    if (N <= 0 || world <= 0) return 0;
    return (size_t)nblk_for(N) * world * sizeof(int32_t);
}

int rbx_shard_route(const int32_t* rows, int64_t N, int world, void* ws, size_t ws_bytes, int32_t* send, int32_t* pos,
                    int32_t* counts, rbx_stream_t stream) {
    const char* who = "rbx_shard_route";
    RBX_RANGE(who);
    RBX_REQUIRE(N >= 0 && N <= INT32_MAX, "%s: N outside int32", who);
    RBX_REQUIRE(world >= 1 && world <= kMaxWorld, "%s: world=%d outside [1,%d]", who, world, kMaxWorld);
    RBX_REQUIRE(counts != nullptr, "%s: counts required", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (N == 0) {
        cudaMemsetAsync(counts, 0, world * sizeof(int32_t), st);
        return RBX_OK;
    }
    RBX_REQUIRE(rows && send && pos && ws, "%s: null pointer", who);
    RBX_REQUIRE(ws_bytes >= rbx_shard_ws_bytes(N, world), "%s: workspace too small (%zu < %zu)", who, ws_bytes,
                rbx_shard_ws_bytes(N, world));
    RBX_REQUIRE((uintptr_t)ws % 4 == 0, "%s: workspace misaligned", who);
    const int nblk = nblk_for(N);
    int32_t* hist = reinterpret_cast<int32_t*>(ws);
    k_shard_hist<<<nblk, kThreads, 0, st>>>(rows, N, world, nblk, hist);
    RBX_LAUNCH_CHECK(who);
    k_shard_scan<<<1, 1024, 0, st>>>(hist, world, nblk, counts);
    RBX_LAUNCH_CHECK(who);
    k_shard_place<<<nblk, kThreads, 0, st>>>(rows, N, world, nblk, hist, send, pos);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

static int permute(const char* who, bool inverse, const float* in, const int32_t* pos, float* out, int64_t N, int D,
                   rbx_stream_t stream) {
    RBX_RANGE(who);
    RBX_REQUIRE(N >= 0 && D >= 1, "%s: bad size", who);
    if (N == 0) return RBX_OK;
    RBX_REQUIRE(in && pos && out, "%s: null pointer", who);
    RBX_REQUIRE(D % 4 != 0 || ((uintptr_t)in % 16 == 0 && (uintptr_t)out % 16 == 0), "%s: misaligned payload", who);
    const int64_t work = D % 4 == 0 ? N * (D / 4) : N * D;
    int64_t ctas = (work + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    if (inverse) k_permute_rows<true><<<(int)ctas, kThreads, 0, rbx_cast_stream(stream)>>>(in, pos, out, N, D);
    else k_permute_rows<false><<<(int)ctas, kThreads, 0, rbx_cast_stream(stream)>>>(in, pos, out, N, D);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_shard_permute(const float* in, const int32_t* pos, float* out, int64_t N, int D, rbx_stream_t stream) {
    return permute("rbx_shard_permute", false, in, pos, out, N, D, stream);
}

int rbx_shard_unroute(const float* recv, const int32_t* pos, float* out, int64_t N, int D, rbx_stream_t stream) {
    return permute("rbx_shard_unroute", true, recv, pos, out, N, D, stream);
}

}  // extern "C"
