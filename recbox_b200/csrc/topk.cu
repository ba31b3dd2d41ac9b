// f3: retrieval evaluation -- exact inner-product top-k of every query user against the whole item corpus, the
// "mask the train items, re-rank, score" step and the ranking metrics.  Replaces faiss.IndexFlatIP.search on the CPU
// (recbox/utils/ann/faiss.py:3-14), evaluate_block's numpy mask / argsort (recbox/core/metrics.py:52-68) and the
// per-user Python metric objects (core/metrics.py:71-200) that MatchingModel.evaluate runs every epoch over all users
// x all items (matching/pytorch/models/match_model.py:205-225).
//
// Top-k without ever materialising the [U, N] score matrix (U = 1000 users x N = 10 M items would be 40 GB):
//   pass A  k_ip_filter : fp32 SIMT GEMM tile (128 users x 128 items, D <= 128 resident in shared memory, 8x8 register
//                         micro-tiles, item tiles double-buffered with cp.async) whose epilogue compares every score
//                         with the user's current k-th best (tau) and appends the survivors -- as sortable 64-bit keys
//                         (score bits | ~index) -- to the user's queue.  After the first few thousand items only
//                         ~k * chunk / seen candidates per user survive a chunk, so the queue traffic is negligible and
//                         the kernel is a plain GEMM bound by the fp32 FMA pipe (2*U*N*D flop; the corpus is read from
//                         HBM once per 128..1024 users and shared through L2 by the co-resident user tiles).
//   pass B  k_select    : one CTA per user merges its queue into the running sorted top-Kp list with a shared-memory
//                         bitonic sort (2048 keys per round) and publishes the new tau.
// Item chunks grow geometrically (4096, 8192, ... up to `chunk`) so that the number of survivors per pass stays ~k even
// while tau is still loose.  Everything is stream-ordered; the host never reads a count.
// Accuracy: fp32 level -- the default pass A forms every product as a 3xTF32 split on the tensor cores with fp32 accumulators
// (k_ip_filter_mma); RBX_TOPK_MMA=0 builds the plain fp32 FMA-chain kernel.  Ties go to the smaller item index.
#include <stdlib.h>
#include "rbx_common.cuh"

namespace {

constexpr int kT = 256;
constexpr int TM = 128, TN = 128;     // CTA tile: users x items
constexpr int kSortN = 2048;          // keys per bitonic round in pass B
constexpr int kMaxK = 1024;

__device__ __forceinline__ uint32_t flip_f32(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float unflip_f32(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// larger key = better: higher score first, then smaller index.  0 is "empty" (every real key is >= 0x007fffff00000000).
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t idx) {
    return ((unsigned long long)flip_f32(score) << 32) | (unsigned long long)(0xffffffffu - idx);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 16 : 0;                       // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct IpParams {
    const float* q;        // [U, D]
    const float* items;    // [N, D]
    int64_t U, N, n0, n1;  // this pass covers items [n0, n1)
    int D;
    const float* tau;                 // [U]
    int* count;                       // [U]
    unsigned long long* queue;        // [U, qcap]
    int64_t qcap;
};

// rows [r0, r0 + 128) of a row-major [rows, D] matrix -> shared tile [128][D + 4] (16-byte chunks, zero past `rows`)
__device__ __forceinline__ void load_tile(float* s, const float* g, int64_t r0, int64_t rows, int D, int LD) {
    const int cpr = D >> 2;                             // 16-byte chunks per row
    for (int t = threadIdx.x; t < 128 * cpr; t += kT) {
        const int r = t / cpr, c = t - r * cpr;
        const bool ok = r0 + r < rows;
        cp_async16(s + r * LD + 4 * c, g + (ok ? (size_t)(r0 + r) * D + 4 * c : 0), ok);
    }
}

__global__ void __launch_bounds__(kT, 2) k_ip_filter(const __grid_constant__ IpParams p) {
    extern __shared__ __align__(16) float smem[];
    const int D = p.D, LD = D + 4;
    float* sA = smem;
    float* sTau = sA + 3 * TM * LD;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t m_base = (int64_t)blockIdx.y * TM;
    const int64_t n_tiles = (p.n1 - p.n0 + TN - 1) / TN;

    load_tile(sA, p.q, m_base, p.U, D, LD);
    if (tid < TM) sTau[tid] = (m_base + tid < p.U) ? __ldcg(p.tau + m_base + tid) : __int_as_float(0x7f800000);
    int64_t t = blockIdx.x;
    if (t < n_tiles) load_tile(sA + TM * LD, p.items, p.n0 + t * TN, p.n1, D, LD);
    cp_async_commit();

    int buf = 0;
    for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
        const int64_t tn = t + gridDim.x;
        if (tn < n_tiles) load_tile(sA + (2 - buf) * TM * LD, p.items, p.n0 + tn * TN, p.n1, D, LD);
        cp_async_commit();
        cp_async_wait<1>();                              // everything but the tile just requested has landed
        __syncthreads();

        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const float* a0 = sA + ty * LD;
        const float* b0 = sA + (1 + buf) * TM * LD + tx * LD;
#pragma unroll 2
        for (int c = 0; c < D; c += 4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * 16 * LD + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(b0 + j * 16 * LD + c);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float s = acc[i][j];
                    s = fmaf(a[i].x, b.x, s);
                    s = fmaf(a[i].y, b.y, s);
                    s = fmaf(a[i].z, b.z, s);
                    s = fmaf(a[i].w, b.w, s);
                    acc[i][j] = s;
                }
            }
        }
        // epilogue: keep what beats the user's current k-th best
        const int64_t n_base = p.n0 + t * TN;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = ty + 16 * i;
            const float tau = sTau[m];                   // +inf for rows past U: nothing survives
            const int64_t u = m_base + m;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t n = n_base + tx + 16 * j;
                if (acc[i][j] > tau && n < p.n1) {
                    const int pos = atomicAdd(p.count + u, 1);
                    if (pos < p.qcap) p.queue[(size_t)u * p.qcap + pos] = make_key(acc[i][j], (uint32_t)n);
                }
            }
        }
        __syncthreads();                                 // sB[buf] is refilled by the next iteration's prefetch
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Tensor-core form of pass A (default, RBX_TOPK_MMA=1): the same 128 x 128 tile, cp.async pipeline and filter epilogue,
// with the products on the tensor cores in "3xTF32": x = hi + lo with hi = tf32(x), lo = tf32(x - hi), and
//   <a, b> ~= sum a_lo*b_hi + a_hi*b_lo + a_hi*b_hi      (fp32 accumulators; the dropped a_lo*b_lo term is ~2^-22 relative)
// which keeps fp32-level accuracy (the score matrix is a true GEMM: [U,D] x [D,N]).  mma.sync.m16n8k8 tf32 fragments:
//   A (16x8, row): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B (8x8, col): b0 (k=t, n=g) b1 (k=t+4, n=g);
//   C (16x8): c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)          with g = lane >> 2, t = lane & 3.
// Both operand tiles are row-major [row][k] in shared memory with a row stride of D + 4 floats, so a fragment load is a
// conflict-free LDS.32 (bank = 4 g + t when (D + 4) mod 32 is an odd multiple of 4).  8 warps = 2 (users) x 4 (items),
// a warp owns 64 x 32 = 4 x 4 mma tiles.  The kernel is written against the legacy warp-level mma; a tcgen05 / TMEM
// version is the next step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(kT, 2) k_ip_filter_mma(const __grid_constant__ IpParams p) {
    extern __shared__ __align__(16) float smem[];
    const int D = p.D, LD = D + 4, KE = (D + 7) & ~7;
    float* sA = smem;
    float* sTau = sA + 3 * TM * LD;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
    const int64_t m_base = (int64_t)blockIdx.y * TM;
    const int64_t n_tiles = (p.n1 - p.n0 + TN - 1) / TN;

    // the 4 padding columns of every tile row are read when D % 8 == 4: keep them zero (cp.async never writes them)
    for (int r = tid; r < 3 * TM; r += kT) *reinterpret_cast<float4*>(sA + r * LD + D) = make_float4(0.f, 0.f, 0.f, 0.f);
    load_tile(sA, p.q, m_base, p.U, D, LD);
    if (tid < TM) sTau[tid] = (m_base + tid < p.U) ? __ldcg(p.tau + m_base + tid) : __int_as_float(0x7f800000);
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) load_tile(sA + TM * LD, p.items, p.n0 + tile * TN, p.n1, D, LD);
    cp_async_commit();

    int buf = 0;
    for (; tile < n_tiles; tile += gridDim.x, buf ^= 1) {
        const int64_t tn = tile + gridDim.x;
        if (tn < n_tiles) load_tile(sA + (2 - buf) * TM * LD, p.items, p.n0 + tn * TN, p.n1, D, LD);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        float acc[4][4][4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                for (int x = 0; x < 4; ++x) acc[mi][ni][x] = 0.f;
        const float* pa = sA + (wm * 64 + g) * LD + t;
        const float* pb = sA + (1 + buf) * TM * LD + (wn * 32 + g) * LD + t;
        for (int k0 = 0; k0 < KE; k0 += 8) {
            uint32_t ah[4][4], al[4][4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const float* r0 = pa + mi * 16 * LD + k0;
                split_tf32(r0[0], ah[mi][0], al[mi][0]);
                split_tf32(r0[8 * LD], ah[mi][1], al[mi][1]);
                split_tf32(r0[4], ah[mi][2], al[mi][2]);
                split_tf32(r0[8 * LD + 4], ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const float* c0 = pb + ni * 8 * LD + k0;
                uint32_t bh[2], bl[2];
                split_tf32(c0[0], bh[0], bl[0]);
                split_tf32(c0[4], bh[1], bl[1]);
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    mma_tf32(acc[mi][ni], al[mi], bh);          // small terms first
                    mma_tf32(acc[mi][ni], ah[mi], bl);
                    mma_tf32(acc[mi][ni], ah[mi], bh);
                }
            }
        }
        // epilogue: keep what beats the user's current k-th best
        const int64_t n_base = p.n0 + tile * TN;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = wm * 64 + mi * 16 + g + 8 * h;
                const float tau = sTau[m];
                const int64_t u = m_base + m;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        const float sc = acc[mi][ni][2 * h + x];
                        const int64_t n = n_base + wn * 32 + ni * 8 + 2 * t + x;
                        if (sc > tau && n < p.n1) {
                            const int pos = atomicAdd(p.count + u, 1);
                            if (pos < p.qcap) p.queue[(size_t)u * p.qcap + pos] = make_key(sc, (uint32_t)n);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

#ifndef RBX_TOPK_MMA
#define RBX_TOPK_MMA 1      // 1: k_ip_filter_mma (3xTF32 on the tensor cores); 0: k_ip_filter (plain fp32 FMA chains)
#endif

// descending bitonic sort of n (power of two) keys in shared memory, whole CTA
__device__ void bitonic_desc(unsigned long long* s, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (n >> 1); t += kT) {
                const int i = 2 * t - (t & (j - 1)), l = i + j;
                const bool desc = (i & k) == 0;
                const unsigned long long a = s[i], b = s[l];
                if ((a < b) == desc) { s[i] = b; s[l] = a; }
            }
            __syncthreads();
        }
    }
}

// pass B: merge user u's queue into its running top list R[u][0..Kp) (sorted, descending), publish tau, clear the queue
__global__ void __launch_bounds__(kT) k_select(unsigned long long* __restrict__ R, const unsigned long long* __restrict__ queue,
                                               int* __restrict__ count, float* __restrict__ tau, int64_t qcap, int Kp, int k) {
    __shared__ unsigned long long s[kSortN];
    const int64_t u = blockIdx.x;
    int cnt = count[u];
    if (cnt == 0) return;                                // uniform per CTA
    if (cnt > qcap) cnt = (int)qcap;
    const int T = kSortN - Kp;
    for (int i = threadIdx.x; i < Kp; i += kT) s[i] = R[(size_t)u * Kp + i];
    for (int off = 0; off < cnt; off += T) {
        const int here = min(T, cnt - off);
        // sort only as many keys as there are: late in a search a user keeps a handful of candidates per pass, and a
        // 256-key network is ~15x cheaper than the 2048-key one
        int n_sort = 2 * Kp;
        while (n_sort < Kp + here) n_sort <<= 1;
        for (int i = threadIdx.x; i < n_sort - Kp; i += kT) s[Kp + i] = i < here ? queue[(size_t)u * qcap + off + i] : 0ull;
        __syncthreads();
        bitonic_desc(s, n_sort);                         // ends with a barrier; the best Kp keys are s[0..Kp)
    }
    for (int i = threadIdx.x; i < Kp; i += kT) R[(size_t)u * Kp + i] = s[i];
    if (threadIdx.x == 0) {
        const unsigned long long kth = s[k - 1];
        tau[u] = kth ? unflip_f32((uint32_t)(kth >> 32)) : __int_as_float(0xff800000);
        count[u] = 0;
    }
}

__global__ void k_topk_init(float* tau, int64_t U) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (int64_t)gridDim.x * blockDim.x)
        tau[i] = __int_as_float(0xff800000);             // -inf
}

__global__ void k_topk_emit(const unsigned long long* __restrict__ R, int Kp, int k, int64_t U, float* __restrict__ scores,
                            int64_t* __restrict__ idx) {
    const int64_t n = U * k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t u = i / k;
        const int r = (int)(i - u * k);
        const unsigned long long key = R[(size_t)u * Kp + r];
        if (scores) scores[i] = key ? unflip_f32((uint32_t)(key >> 32)) : __int_as_float(0xff800000);
        if (idx) idx[i] = key ? (int64_t)(0xffffffffu - (uint32_t)key) : -1;
    }
}

inline int pow2_at_least(int k) {
    int p = 1;
    while (p < k) p <<= 1;
    return p;
}

struct TopkWs {
    float* tau;
    int* count;
    unsigned long long* R;
    unsigned long long* queue;
};
inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
inline size_t topk_fixed_bytes(int64_t U, int Kp) { return align256(U * 4) * 2 + align256((size_t)U * Kp * 8); }

// ---------------------------------------------------------------------------------------------
// mask + re-rank + metrics (evaluate_block, core/metrics.py:52-68, and the metric classes :71-200)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_sorted(const int64_t* a, int64_t lo, int64_t hi, int64_t x) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = a[mid];
        if (v == x) return true;
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return false;
}

// One warp per user.  cand[u, 0..T) is the top-T list (descending); items the user clicked in train are pushed behind
// every other candidate (scores += -1e9 * mask; argsort(-scores)), the first kmax of the result are `ranked`, hit[u,r]
// says whether ranked[u,r] is one of the user's valid items.
__global__ void __launch_bounds__(kT) k_mask_rerank(const int64_t* __restrict__ cand, int T, int64_t U,
                                                    const int64_t* __restrict__ train_ptr, const int64_t* __restrict__ train_items,
                                                    const int64_t* __restrict__ valid_ptr, const int64_t* __restrict__ valid_items,
                                                    int kmax, int64_t* __restrict__ ranked, uint8_t* __restrict__ hit) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * (kT / 32) + (threadIdx.x >> 5);
    if (u >= U) return;
    const int64_t* c = cand + u * T;
    const int64_t t0 = train_ptr ? train_ptr[u] : 0, t1 = train_ptr ? train_ptr[u + 1] : 0;
    const int64_t v0 = valid_ptr[u], v1 = valid_ptr[u + 1];
    int out = 0;
    // pass 0: candidates not clicked in train, in order; pass 1: the clicked ones (they only surface when fewer than
    // kmax others exist)
    for (int pass = 0; pass < 2 && out < kmax; ++pass) {
        for (int base = 0; base < T && out < kmax; base += 32) {
            const int i = base + lane;
            const int64_t item = i < T ? c[i] : -1;
            const bool masked = item >= 0 && in_sorted(train_items, t0, t1, item);
            const bool take = item >= 0 && (masked == (pass == 1));
            const uint32_t bal = __ballot_sync(0xffffffffu, take);
            const int pos = out + __popc(bal & ((1u << lane) - 1));
            if (take && pos < kmax) {
                ranked[u * kmax + pos] = item;
                hit[u * kmax + pos] = in_sorted(valid_items, v0, v1, item) ? 1 : 0;
            }
            out += __popc(bal);
        }
    }
    for (int r = min(out, kmax) + lane; r < kmax; r += 32) {      // corpus smaller than kmax
        ranked[u * kmax + r] = -1;
        hit[u * kmax + r] = 0;
    }
}

// metric kinds (core/metrics.py): 0 Recall, 1 nRecall, 2 Precision, 3 F1, 4 DCG, 5 NDCG, 6 MRR, 7 HitRate, 8 MAP.
// One thread per (user, metric); float64 like the reference's Python floats.
__global__ void k_rank_metrics(const uint8_t* __restrict__ hit, int kmax, int64_t U, const int64_t* __restrict__ valid_ptr,
                               const int* __restrict__ kinds, const int* __restrict__ ks, int M, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U * M) return;
    const int64_t u = i / M;
    const int m = (int)(i - u * M);
    const int kind = kinds[m], k = min(ks[m], kmax);
    const double n_true = (double)(valid_ptr[u + 1] - valid_ptr[u]);
    const uint8_t* h = hit + u * kmax;
    double hits = 0, dcg = 0, mrr = 0, prec = 0;
    for (int r = 0; r < k; ++r) {
        if (h[r]) {
            hits += 1;
            dcg += 1.0 / log(2.0 + r);
            mrr += 1.0 / (r + 1.0);
            prec += hits / (r + 1.0);
        }
    }
    double v = 0;
    const double recall = hits / (n_true + 1e-12), precision = hits / (ks[m] + 1e-12);
    switch (kind) {
        case 0: v = recall; break;
        case 1: v = hits / fmin((double)ks[m], n_true + 1e-12); break;
        case 2: v = precision; break;
        case 3: v = 2 * precision * recall / (precision + recall + 1e-12); break;
        case 4: v = dcg; break;
        case 5: {
            double idcg = 0;                             // dcg_fn(true_items[:k], true_items): every one of them hits
            const int nt = (int)fmin((double)ks[m], n_true);
            for (int r = 0; r < nt; ++r) idcg += 1.0 / log(2.0 + r);
            v = dcg / (idcg + 1e-12);
        } break;
        case 6: v = mrr; break;
        case 7: v = hits > 0 ? 1.0 : 0.0; break;
        default: v = prec / (hits + 1e-12); break;
    }
    out[i] = v;
}

}  // namespace

extern "C" {

size_t rbx_topk_ws_bytes(int64_t U, int k, int64_t chunk) {
    if (U <= 0 || k <= 0 || k > kMaxK || chunk <= 0) return 0;
    chunk = (chunk + TN - 1) / TN * TN;
    return topk_fixed_bytes(U, pow2_at_least(k)) + align256((size_t)U * chunk * 8);
}

int rbx_topk_ip(const float* q, const float* items, int64_t U, int64_t N, int D, int k, int64_t chunk, float* out_scores,
                int64_t* out_idx, void* ws, size_t ws_bytes, rbx_stream_t stream) {
    const char* who = "rbx_topk_ip";
    RBX_RANGE(who);
    RBX_REQUIRE(U >= 0 && N >= 0 && N < 0xffffffffll, "%s: bad size", who);
    RBX_REQUIRE(k >= 1 && k <= kMaxK, "%s: k=%d outside [1, %d]", who, k, kMaxK);
    RBX_REQUIRE(D >= 4 && D <= 128 && D % 4 == 0, "%s: D=%d (needs a multiple of 4 in [4, 128])", who, D);
    if (U == 0) return RBX_OK;
    RBX_REQUIRE(q && (items || N == 0) && (out_scores || out_idx) && ws, "%s: null pointer", who);
    RBX_REQUIRE((uintptr_t)q % 16 == 0 && (uintptr_t)items % 16 == 0 && (uintptr_t)ws % 256 == 0, "%s: unaligned pointer", who);
    chunk = (chunk + TN - 1) / TN * TN;
    RBX_REQUIRE(chunk >= TN && ws_bytes >= rbx_topk_ws_bytes(U, k, chunk), "%s: workspace too small (%zu < %zu)", who,
                ws_bytes, rbx_topk_ws_bytes(U, k, chunk));
    cudaStream_t st = rbx_cast_stream(stream);
    const int Kp = pow2_at_least(k);
    char* w = reinterpret_cast<char*>(ws);
    TopkWs t;
    t.tau = reinterpret_cast<float*>(w);               w += align256(U * 4);
    t.count = reinterpret_cast<int*>(w);               w += align256(U * 4);
    t.R = reinterpret_cast<unsigned long long*>(w);    w += align256((size_t)U * Kp * 8);
    t.queue = reinterpret_cast<unsigned long long*>(w);
    cudaError_t me = cudaMemsetAsync(t.count, 0, (size_t)U * 4, st);
    if (me == cudaSuccess) me = cudaMemsetAsync(t.R, 0, (size_t)U * Kp * 8, st);
    if (me != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: memset: %s", who, cudaGetErrorString(me));
    k_topk_init<<<(int)((U + 255) / 256 < 1024 ? (U + 255) / 256 : 1024), 256, 0, st>>>(t.tau, U);
    RBX_LAUNCH_CHECK(who);

    const size_t smem = ((size_t)3 * TM * (D + 4) + TM) * 4;
    cudaError_t e = cudaFuncSetAttribute(k_ip_filter_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ip_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: %s", who, cudaGetErrorString(e));
    const int64_t UT = (U + TM - 1) / TM;
    RBX_REQUIRE(UT <= 65535, "%s: U too large for one call", who);
    const int resident = rbx_sm_count() * 2;
    IpParams p;
    p.q = q; p.items = items; p.U = U; p.N = N; p.D = D; p.tau = t.tau; p.count = t.count; p.queue = t.queue; p.qcap = chunk;
    // pass A engine: 2 (default) = tcgen05 GEMM with the filter epilogue (csrc/gemm.cu); 1 = warp-level mma.sync; 0 = fp32 FMA
    const char* eng_s = getenv("RBX_TOPK_ENGINE");
    const int engine = (eng_s && *eng_s) ? atoi(eng_s) : 2;
    int64_t c = chunk < 4096 ? chunk : 4096;
    for (int64_t n0 = 0; n0 < N;) {
        const int64_t n1 = n0 + c < N ? n0 + c : N;
        p.n0 = n0; p.n1 = n1;
        const int64_t tiles = (n1 - n0 + TN - 1) / TN;
        int64_t gx = resident / UT;
        if (gx < 1) gx = 1;
        if (gx > tiles) gx = tiles;
        if (engine == 2) {
            if (int rc = rbx_topk_filter_tc(q, items, U, n0, n1, D, t.tau, t.count, t.queue, chunk, st)) return rc;
        } else if (engine == 1) {
            k_ip_filter_mma<<<dim3((unsigned)gx, (unsigned)UT), kT, smem, st>>>(p);
        } else {
            k_ip_filter<<<dim3((unsigned)gx, (unsigned)UT), kT, smem, st>>>(p);
        }
        RBX_LAUNCH_CHECK(who);
        k_select<<<(unsigned)U, kT, 0, st>>>(t.R, t.queue, t.count, t.tau, chunk, Kp, k);
        RBX_LAUNCH_CHECK(who);
        n0 = n1;
        c = 2 * c < chunk ? 2 * c : chunk;
    }
    const int64_t n_out = U * k;
    k_topk_emit<<<(int)((n_out + 255) / 256 < 2048 ? (n_out + 255) / 256 : 2048), 256, 0, st>>>(t.R, Kp, k, U, out_scores, out_idx);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_rank_metrics(const int64_t* cand, int T, int64_t U, const int64_t* train_ptr, const int64_t* train_items,
                     const int64_t* valid_ptr, const int64_t* valid_items, int kmax, const int* kinds, const int* ks, int M,
                     int64_t* ranked, uint8_t* hit, double* out, rbx_stream_t stream) {
    const char* who = "rbx_rank_metrics";
    RBX_RANGE(who);
    RBX_REQUIRE(U >= 0 && T >= 1 && kmax >= 1 && M >= 0, "%s: bad size", who);
    if (U == 0) return RBX_OK;
    RBX_REQUIRE(cand && valid_ptr && valid_items && ranked && hit, "%s: null pointer", who);
    RBX_REQUIRE((train_ptr == nullptr) == (train_items == nullptr), "%s: train_ptr / train_items must come together", who);
    RBX_REQUIRE(M == 0 || (kinds && ks && out), "%s: null metric arrays", who);
    cudaStream_t st = rbx_cast_stream(stream);
    k_mask_rerank<<<(unsigned)((U + kT / 32 - 1) / (kT / 32)), kT, 0, st>>>(cand, T, U, train_ptr, train_items, valid_ptr,
                                                                            valid_items, kmax, ranked, hit);
    RBX_LAUNCH_CHECK(who);
    if (M > 0) {
        k_rank_metrics<<<(unsigned)((U * M + 255) / 256), 256, 0, st>>>(hit, kmax, U, valid_ptr, kinds, ks, M, out);
        RBX_LAUNCH_CHECK(who);
    }
    return RBX_OK;
}

}  // extern "C"
