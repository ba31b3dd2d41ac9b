// f2: per-epoch negative sampling on the GPU.  Replaces sampling_block
// (recbox/matching/pytorch/dataloaders/h5_generator.py:72-95): np.random.choice(num_items, size=(n_queries, num_negs),
// replace=True) -- uniform over the corpus, or, with ignore_pos_items, uniform over the items the query user has NOT
// interacted with (the reference zeroes their probabilities and renormalises, one O(num_items) pass per query row;
// rejection sampling draws from exactly that distribution).  The reference then hstacks the positive column in front
// (h5_generator.py:176-177); `pos` does that in the same launch.
//
// RNG: counter-based (one independent stream per output element: key = seed, counter = element index, retry number), a
// 64-bit SplitMix/Murmur-style mixer; ids come from the high bits by multiply-shift (bias < 2^-32 * num_items).  The
// draws are reproducible from (seed, element index) alone and independent of the launch geometry.  They are NOT numpy's
// MT19937 stream: parity with the reference is distributional (tests: range, rejection of positives, chi-square
// uniformity, determinism), as it is between any two seeds of the reference itself.
#include "rbx_common.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ int64_t draw(uint64_t seed, uint64_t elem, uint32_t attempt, int64_t num_items) {
    const uint64_t r = mix64(mix64(seed ^ (elem * 0xd1342543de82ef95ull)) + attempt);
    return (int64_t)__umul64hi(r, (uint64_t)num_items);        // uniform in [0, num_items)
}

__device__ __forceinline__ bool in_sorted(const int64_t* a, int64_t lo, int64_t hi, int64_t x) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = a[mid];
        if (v == x) return true;
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return false;
}

// out [n_queries, lead + num_negs]: column 0 = pos[q] when lead == 1, then the sampled negatives
__global__ void __launch_bounds__(256) k_sample_negs(int64_t n_queries, int num_negs, int64_t num_items, uint64_t seed,
                                                     const int64_t* __restrict__ pos, const int64_t* __restrict__ user_of_query,
                                                     const int64_t* __restrict__ pos_ptr, const int64_t* __restrict__ pos_items,
                                                     int64_t* __restrict__ out, int lead, int* __restrict__ gave_up) {
    const int64_t total = n_queries * num_negs;
    const int ld = lead + num_negs;
    const bool small = total < (1ll << 31);                    // 32-bit division: the 64-bit one costs more than the draw
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t qi = small ? (int64_t)((uint32_t)e / (uint32_t)num_negs) : e / num_negs;
        const int j = (int)(e - qi * num_negs);
        int64_t id = draw(seed, (uint64_t)e, 0, num_items);
        if (pos_ptr) {
            const int64_t u = user_of_query ? user_of_query[qi] : qi;
            const int64_t lo = pos_ptr[u], hi = pos_ptr[u + 1];
            uint32_t attempt = 0;
            while (in_sorted(pos_items, lo, hi, id)) {
                if (++attempt > 4096u) { atomicAdd(gave_up, 1); break; }   // the user interacted with ~every item
                id = draw(seed, (uint64_t)e, attempt, num_items);
            }
        }
        out[qi * ld + lead + j] = id;
        if (lead && j == 0) out[qi * ld] = pos[qi];
    }
}

}  // namespace

extern "C" {

int rbx_sample_negatives(int64_t n_queries, int num_negs, int64_t num_items, uint64_t seed, const int64_t* pos,
                         const int64_t* user_of_query, const int64_t* pos_ptr, const int64_t* pos_items, int64_t* out,
                         int* gave_up, rbx_stream_t stream) {
    const char* who = "rbx_sample_negatives";
    RBX_RANGE(who);
    RBX_REQUIRE(n_queries >= 0 && num_negs >= 0 && num_items >= 1, "%s: bad size", who);
    if (n_queries == 0 || (num_negs == 0 && !pos)) return RBX_OK;
    RBX_REQUIRE(out != nullptr, "%s: null output", who);
    RBX_REQUIRE((pos_ptr == nullptr) == (pos_items == nullptr), "%s: pos_ptr / pos_items must come together", who);
    RBX_REQUIRE(!pos_ptr || gave_up, "%s: rejection sampling needs the gave_up counter", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (num_negs == 0) {                      // only the positive column
        cudaMemcpyAsync(out, pos, (size_t)n_queries * 8, cudaMemcpyDeviceToDevice, st);
        return RBX_OK;
    }
    const int64_t total = n_queries * num_negs;
    int64_t grid = (total + 255) / 256;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (grid > cap) grid = cap;
    k_sample_negs<<<(int)grid, 256, 0, st>>>(n_queries, num_negs, num_items, seed, pos, user_of_query, pos_ptr, pos_items, out,
                                            pos ? 1 : 0, gave_up);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
