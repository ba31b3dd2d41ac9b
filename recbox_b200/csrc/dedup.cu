// a14: sorted unique + inverse + first occurrence of a bag of item ids, for ids bounded by a
// known vocabulary (every id the reference's loaders produce is < vocab_size of its feature).
//
// Replaces torch.unique(item_indexes.flatten(), return_inverse=True, sorted=True) and the
// flip/scatter_ "return_index" trick of collate_fn_unique
// (recbox/matching/pytorch/dataloaders/h5_generator.py:45-53), which sorts on the CPU inside the
// DataLoader worker.  No sort here: because ids live in [0, vocab) the sorted-unique set is a
// BITMAP of vocab bits (1.25 MB for 10 M items -- it lives in L2), the rank of an id is a
// popcount prefix over that bitmap, and the whole job is five streaming passes:
//   1 clear bitmap            2 mark: atomicOr(bitmap[id>>5], 1 << (id&31))
//   3 per-CTA popcount sums   4 scan of the CTA sums (one CTA), then the exclusive prefix of every word
//   5 rank(id) = prefix[id>>5] + popc(bitmap[id>>5] & lower bits): uniq[rank] = id, inverse[i] = rank, first[rank] = min(i)
// Integer work: results are bit-identical to numpy/torch unique (tests compare with the oracle).
#include "rbx_common.cuh"

namespace {

constexpr int kT = 256;             // threads per CTA
constexpr int kWPT = 8;             // bitmap words per thread in the count / emit passes
constexpr int kChunk = kT * kWPT;   // words per CTA (= 65 536 ids of vocabulary)

// workspace: counters (64 B) | wp[W] = (bitmap word, exclusive popcount prefix) pairs | blocksum[NB] | byte map [32 W]
struct UniqueWs {
    uint2* wp;           // [W]   .x bitmap word, .y rank of the word's first id
    uint32_t* blocksum;  // [NB]  popcount of every chunk, then its exclusive prefix
    uint8_t* bytemap;    // [32 W] dense mode: one byte per id, written with plain stores (no atomics)
    int64_t* counters;   // [2]   reserved
};

__host__ __device__ inline int64_t words_of(int64_t vocab) { return (vocab + 31) / 32; }

size_t ws_bytes(int64_t vocab) {
    const int64_t W = words_of(vocab), NB = (W + kChunk - 1) / kChunk;
    return 64 + (((size_t)W * 8 + 15) & ~(size_t)15) + (((size_t)NB * 4 + 15) & ~(size_t)15) + (size_t)W * 32;
}

UniqueWs carve(void* ws, int64_t vocab) {
    const int64_t W = words_of(vocab), NB = (W + kChunk - 1) / kChunk;
    UniqueWs u;
    char* p = reinterpret_cast<char*>(ws);
    u.counters = reinterpret_cast<int64_t*>(p);                               // 64 bytes reserved
    u.wp = reinterpret_cast<uint2*>(p + 64);
    const size_t wp_bytes = ((size_t)W * 8 + 15) & ~(size_t)15;
    u.blocksum = reinterpret_cast<uint32_t*>(p + 64 + wp_bytes);
    u.bytemap = reinterpret_cast<uint8_t*>(p + 64 + wp_bytes + (((size_t)NB * 4 + 15) & ~(size_t)15));   // 16-byte aligned
    return u;
}

// Sparse batches (n << vocab): set bits with atomicOr -- few ids share a word, the bitmap is all that is cleared.
template <typename IdT>
__global__ void __launch_bounds__(kT) k_mark(const IdT* __restrict__ ids, int64_t n, int64_t vocab, uint2* wp, int64_t* n_out) {
    int64_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT) {
        const int64_t id = (int64_t)ids[i];
        if (id < 0 || id >= vocab) { ++bad; continue; }
        const uint32_t bit = 1u << (id & 31);
        uint32_t* word = &wp[id >> 5].x;
        // most ids of a batch are distinct: test first so repeated (hot) ids do not serialise on the atomic
        if (!(__ldcg(word) & bit)) atomicOr(word, bit);
    }
    if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(n_out + 1), (unsigned long long)bad);
}

// Dense batches (n comparable to vocab, e.g. the rows a 65 536 x 26 batch touches in a 1 M-row table): dozens of ids
// fall into every bitmap word and their atomics serialise in L2 (31 us of the 90 us total, ncu launch list r1z).  Plain
// byte stores of the same value race harmlessly instead; k_count folds the bytes into bitmap words.
template <typename IdT>
__global__ void __launch_bounds__(kT) k_mark_bytes(const IdT* __restrict__ ids, int64_t n, int64_t vocab, uint8_t* bytemap,
                                                   int64_t* n_out) {
    int64_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT) {
        const int64_t id = (int64_t)ids[i];
        if (id < 0 || id >= vocab) { ++bad; continue; }
        bytemap[id] = 1;
    }
    if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(n_out + 1), (unsigned long long)bad);
}

// 32 flag bytes (0 / 1) -> one bitmap word; (x * 0x01020408) >> 24 gathers the four flags of a 32-bit group
__device__ __forceinline__ uint32_t fold_flags(const uint8_t* p) {
    const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
    const uint32_t g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) word |= ((((g[k] & 0x01010101u) * 0x01020408u) >> 24) & 0xfu) << (4 * k);
    return word;
}

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t s_warp[kT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) {
        const uint32_t s = s_warp[w];
        if (w < warp) base += s;
        tot += s;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kT) k_count(uint2* wp, const uint8_t* __restrict__ bytemap, int64_t W, uint32_t* blocksum) {
    const int64_t w0 = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * kWPT;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kWPT; ++k) {
        if (w0 + k < W) {
            uint32_t word;
            if (bytemap) {
                word = fold_flags(bytemap + (w0 + k) * 32);
                wp[w0 + k].x = word;
            } else {
                word = wp[w0 + k].x;
            }
            c += __popc(word);
        }
    }
    uint32_t total;
    (void)block_exclusive_scan(c, &total);
    if (threadIdx.x == 0) blocksum[blockIdx.x] = total;
}

// one CTA: exclusive scan of the chunk sums in place, total -> n_out[0]
__global__ void __launch_bounds__(kT) k_scan_chunks(uint32_t* blocksum, int64_t NB, int64_t* n_out) {
    uint32_t carry = 0;
    for (int64_t base = 0; base < NB; base += kT) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < NB ? blocksum[i] : 0u;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, &total);
        if (i < NB) blocksum[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) n_out[0] = (int64_t)carry;
}

// exclusive popcount prefix of every bitmap word (rank of the first id of the word), stored next to the word
__global__ void __launch_bounds__(kT) k_prefix(uint2* wp, int64_t W, const uint32_t* __restrict__ blocksum) {
    const int64_t w0 = (int64_t)blockIdx.x * kChunk + (int64_t)threadIdx.x * kWPT;
    uint32_t cnt[kWPT], c = 0;
#pragma unroll
    for (int k = 0; k < kWPT; ++k) {
        cnt[k] = (w0 + k < W) ? __popc(wp[w0 + k].x) : 0u;
        c += cnt[k];
    }
    uint32_t total;
    uint32_t rank = blocksum[blockIdx.x] + block_exclusive_scan(c, &total);
#pragma unroll
    for (int k = 0; k < kWPT; ++k) {
        if (w0 + k < W) wp[w0 + k].y = rank;
        rank += cnt[k];
    }
}

__global__ void __launch_bounds__(kT) k_fill_i64(int64_t* p, int64_t n, int64_t v) {
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT) p[i] = v;
}

// rank of every id = word prefix + popcount of the lower bits.  The sorted unique list is written from HERE
// (uniq[rank] = id: all duplicates store the same value), so no pass has to walk the bitmap bit by bit -- the first
// version emitted the uniques with a serial per-bit loop over 8 words per thread in a 15-CTA launch, which was most
// of its 134 us on a 1 M-row vocabulary.
template <typename IdT>
__global__ void __launch_bounds__(kT) k_inverse(const IdT* __restrict__ ids, int64_t n, int64_t vocab, const uint2* __restrict__ wp,
                                                IdT* uniq, IdT* inverse, int64_t* first) {
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < n; i += (int64_t)gridDim.x * kT) {
        const int64_t id = (int64_t)ids[i];
        if (id < 0 || id >= vocab) {
            if (inverse) inverse[i] = (IdT)-1;
            continue;
        }
        const uint2 e = __ldg(wp + (id >> 5));               // one 8-byte read: the word and its prefix
        const uint32_t rank = e.y + __popc(e.x & ((1u << (id & 31)) - 1u));
        if (uniq) uniq[rank] = (IdT)id;
        if (inverse) inverse[i] = (IdT)rank;
        if (first) atomicMin(reinterpret_cast<unsigned long long*>(first + rank), (unsigned long long)i);
    }
}

template <typename IdT>
int unique_impl(const char* who, const IdT* ids, int64_t n, int64_t vocab, void* ws, size_t ws_have, IdT* uniq, int64_t* first,
                IdT* inverse, int64_t* n_out, rbx_stream_t stream) {
    RBX_RANGE(who);
    RBX_REQUIRE(n >= 0 && vocab >= 1, "%s: n=%lld vocab=%lld", who, (long long)n, (long long)vocab);
    RBX_REQUIRE(vocab <= (int64_t)INT32_MAX * 32, "%s: vocab=%lld too large for the bitmap", who, (long long)vocab);
    RBX_REQUIRE(sizeof(IdT) == 8 || vocab <= (int64_t)INT32_MAX, "%s: int32 ids need vocab < 2^31", who);
    RBX_REQUIRE(n == 0 || ids, "%s: ids is null", who);
    RBX_REQUIRE(n_out, "%s: n_out (device int64[2]) is required", who);
    RBX_REQUIRE(ws && ws_have >= ws_bytes(vocab), "%s: workspace of %zu bytes needed (rbx_unique_ws_bytes), got %zu", who,
                ws_bytes(vocab), ws_have);
    RBX_REQUIRE((uintptr_t)ws % 16 == 0, "%s: workspace must be 16-byte aligned", who);
    cudaStream_t st = rbx_cast_stream(stream);
    const int64_t W = words_of(vocab), NB = (W + kChunk - 1) / kChunk;
    const UniqueWs u = carve(ws, vocab);
    const bool dense = n * 4 >= vocab;                     // >= 1 id per 4 vocabulary entries: byte map, no atomics
    cudaError_t e = dense ? cudaMemsetAsync(u.bytemap, 0, (size_t)W * 32, st)
                          : cudaMemsetAsync(ws, 0, 64 + (size_t)W * 8, st);   // counters + (word, prefix) pairs
    if (e == cudaSuccess) e = cudaMemsetAsync(n_out, 0, 16, st);
    if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: memset: %s", who, cudaGetErrorString(e));
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    int64_t g = (n + kT - 1) / kT;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    if (n > 0) {
        if (dense) k_mark_bytes<IdT><<<(int)g, kT, 0, st>>>(ids, n, vocab, u.bytemap, n_out);
        else k_mark<IdT><<<(int)g, kT, 0, st>>>(ids, n, vocab, u.wp, n_out);
    }
    k_count<<<(int)NB, kT, 0, st>>>(u.wp, dense ? u.bytemap : nullptr, W, u.blocksum);
    k_scan_chunks<<<1, kT, 0, st>>>(u.blocksum, NB, n_out);
    k_prefix<<<(int)NB, kT, 0, st>>>(u.wp, W, u.blocksum);
    if (n > 0 && first) {
        const int64_t cap_u = n < vocab ? n : vocab;          // capacity of `first` (rbx_unique_ids_* contract)
        k_fill_i64<<<(int)((cap_u + kT - 1) / kT < cap ? (cap_u + kT - 1) / kT : cap), kT, 0, st>>>(first, cap_u, n);
    }
    if (n > 0 && (uniq || inverse || first))
        k_inverse<IdT><<<(int)g, kT, 0, st>>>(ids, n, vocab, u.wp, uniq, inverse, first);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // namespace

extern "C" {

size_t rbx_unique_ws_bytes(int64_t vocab) { return vocab < 1 ? 0 : ws_bytes(vocab); }

int rbx_unique_ids_i64(const int64_t* ids, int64_t n, int64_t vocab, void* ws, size_t ws_bytes_have, int64_t* uniq,
                       int64_t* first, int64_t* inverse, int64_t* n_out, rbx_stream_t stream) {
    return unique_impl<int64_t>("rbx_unique_ids_i64", ids, n, vocab, ws, ws_bytes_have, uniq, first, inverse, n_out, stream);
}

int rbx_unique_ids_i32(const int32_t* ids, int64_t n, int64_t vocab, void* ws, size_t ws_bytes_have, int32_t* uniq,
                       int64_t* first, int32_t* inverse, int64_t* n_out, rbx_stream_t stream) {
    return unique_impl<int32_t>("rbx_unique_ids_i32", ids, n, vocab, ws, ws_bytes_have, uniq, first, inverse, n_out, stream);
}

}  // extern "C"
