// a13  dense tail of the hot path: fp32 GEMM with fused bias / ReLU / ReLU-mask epilogue on the 5th-generation tensor
// cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA), sm_100a only.
//
// Replaces nn.Linear + activation of MLP_Block (recbox/ranking/pytorch/layers/blocks/mlp_block.py:43-61) and their
// autograd backward inside loss.backward() (ranking_model.py:191-197):
//     forward   H' = relu(H W^T + b)            A = H  [M,K] K-major,  B = W  [N,K] K-major
//     backward  dH = (dZ W) * (H > 0)           A = dZ [M,N] K-major,  B = W  [N,K] MN-major (the reduction runs over N)
//               dW = dZ^T H                      A = dZ MN-major,       B = H  MN-major     (reduction over the batch, split-K)
//
// fp32 contract (logits / grads within 1e-5 rel): every operand is split x = hi + lo with hi = the 19 bits the tensor
// core reads of an fp32 word (kind::tf32 ignores the low 13 mantissa bits) and lo = x - hi (exact in fp32); the product
// is a_lo b_hi + a_hi b_lo + a_hi b_hi accumulated in fp32 in TMEM (3xTF32).  The split costs no global traffic: TMA lands
// the raw fp32 tile in shared memory, four "transform" warps write the lo plane next to it (element-wise, so it inherits
// the 128B / 64B swizzle of the tile), and the MMA warp issues three MMAs per k-step: hi planes straight from the TMA
// buffer, lo planes from the transform buffer.  precision = 1 skips the transform (plain TF32, ~1e-3 rel), stated option.
//
// One 128 x BN output tile per CTA (BN <= 256 runtime), 192 threads:
//   warp 0      TMA producer (one lane): cp.async.bulk.tensor.2d -> smem stage, mbarrier complete_tx
//   warp 1      TMEM allocator + MMA issuer (one lane): tcgen05.mma.cta_group::1.kind::tf32, tcgen05.commit -> mbarriers
//   warps 2..5  transform (lo planes) during the main loop, then the epilogue: tcgen05.ld 32x32b -> bias / ReLU / mask ->
//               global stores (or red.add for split-K)
// Stage ring: full[s] (TMA landed) -> xform[s] (lo plane written, fence.proxy.async) -> MMA -> empty[s] (tcgen05.commit).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "rbx_common.cuh"

namespace {

constexpr int kBM = 128;            // rows of one output tile = UMMA M (cta_group::1)
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct GemmArgs {
    CUtensorMap ta, tb;
    float* C;
    const float* bias;
    const float* mask;
    int64_t ldc, ldmask;
    int M, N, K;
    int BN, KB, stages;
    int a_mn, b_mn, prec, act, atomic;
    int kb_total, kb_per_split;
    uint32_t tmem_cols, idesc;
    uint32_t a_bytes, b_bytes;        // bytes of one raw A / B tile in a stage
    uint32_t a_desc_hi, b_desc_hi;    // upper 32 bits of the shared-memory matrix descriptors
    uint32_t a_lbo_sbo, b_lbo_sbo;    // bits [16,30) LBO>>4 of the lower word (start address is OR-ed in)
    uint32_t a_kstep, b_kstep;        // descriptor start-address increment (>>4) per UMMA k-step (8 tf32)
    uint32_t a_boxes, b_boxes;        // TMA boxes per tile (1 for K-major, tile/32 for MN-major)
    uint32_t a_box_bytes, b_box_bytes;
    int write_hi;                     // debug: also overwrite the raw tile with the masked hi words
    // top-k filter epilogue (f3): scores above the row's running k-th best go to the row's candidate queue, nothing else is stored
    int filter;
    const float* tau;                 // [M]
    int* count;                       // [M]
    unsigned long long* queue;        // [M, qcap]
    int64_t qcap, col_off;
};

// sortable 64-bit key of (score, item): larger = better, ties to the smaller index (same encoding as csrc/topk.cu)
__device__ __forceinline__ unsigned long long topk_key(float score, uint32_t idx) {
    const uint32_t b = __float_as_uint(score);
    const uint32_t f = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)f << 32) | (unsigned long long)(0xffffffffu - idx);
}

// ---- raw PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x = hi + lo, hi = the tf32 the tensor core reads of x (low 13 mantissa bits ignored), lo exact in fp32
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

__global__ void __launch_bounds__(kThreads, 1) k_gemm_tc(const __grid_constant__ GemmArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzle atoms need 1024-byte aligned tiles
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t plane_bytes = p.a_bytes + p.b_bytes;                   // raw A | raw B
    const uint32_t stage_bytes = p.prec == 3 ? 2 * plane_bytes : plane_bytes;
    const uint32_t bars = smem_base + (uint32_t)p.stages * stage_bytes;   // full[S] | xform[S] | empty[S] | accum | tmem ptr
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xform_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + s); };
    const uint32_t accum_bar = bars + 8u * (3 * kMaxStages);
    const uint32_t tmem_slot = accum_bar + 8u;
    const int n0 = blockIdx.x * p.BN, m0 = blockIdx.y * kBM;
    const int kb0 = blockIdx.z * p.kb_per_split;
    int nkb = p.kb_total - kb0;
    if (nkb > p.kb_per_split) nkb = p.kb_per_split;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xform_bar(s), 4);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1u);
                mbar_arrive_expect_tx(full_bar(s), plane_bytes);
                const uint32_t dst = smem_base + (uint32_t)s * stage_bytes;
                const int k0 = (kb0 + i) * p.KB;
                if (p.a_mn) {
                    for (uint32_t b = 0; b < p.a_boxes; ++b) tma_load_2d(dst + b * p.a_box_bytes, &p.ta, full_bar(s), m0 + 32 * (int)b, k0);
                } else {
                    tma_load_2d(dst, &p.ta, full_bar(s), k0, m0);
                }
                if (p.b_mn) {
                    for (uint32_t b = 0; b < p.b_boxes; ++b)
                        tma_load_2d(dst + p.a_bytes + b * p.b_box_bytes, &p.tb, full_bar(s), n0 + 32 * (int)b, k0);
                } else {
                    tma_load_2d(dst + p.a_bytes, &p.tb, full_bar(s), k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const int ksteps = p.KB / 8;
            uint32_t acc = 0;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
                mbar_wait(p.prec == 3 ? xform_bar(s) : full_bar(s), ph);
                tc_fence_after();
                const uint32_t a_hi = smem_base + (uint32_t)s * stage_bytes, b_hi = a_hi + p.a_bytes;
                const uint32_t a_lo = a_hi + plane_bytes, b_lo = b_hi + plane_bytes;
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t da_hi = ((uint64_t)p.a_desc_hi << 32) | (p.a_lbo_sbo | (((a_hi >> 4) + k * p.a_kstep) & 0x3fffu));
                    const uint64_t db_hi = ((uint64_t)p.b_desc_hi << 32) | (p.b_lbo_sbo | (((b_hi >> 4) + k * p.b_kstep) & 0x3fffu));
                    if (p.prec == 3) {
                        const uint64_t da_lo = ((uint64_t)p.a_desc_hi << 32) | (p.a_lbo_sbo | (((a_lo >> 4) + k * p.a_kstep) & 0x3fffu));
                        const uint64_t db_lo = ((uint64_t)p.b_desc_hi << 32) | (p.b_lbo_sbo | (((b_lo >> 4) + k * p.b_kstep) & 0x3fffu));
                        tc_mma_tf32(tmem_base, da_lo, db_hi, p.idesc, acc);     // small terms first
                        tc_mma_tf32(tmem_base, da_hi, db_lo, p.idesc, 1u);
                        tc_mma_tf32(tmem_base, da_hi, db_hi, p.idesc, 1u);
                    } else {
                        tc_mma_tf32(tmem_base, da_hi, db_hi, p.idesc, acc);
                    }
                    acc = 1u;
                }
                tc_commit(empty_bar(s));       // frees the stage when the MMAs above have read it
            }
            tc_commit(accum_bar);              // accumulator complete
        }
    } else {
        // ===== transform warps (3xTF32 lo planes), then the epilogue =====
        const int t = threadIdx.x - 64;        // 0..127
        if (p.prec == 3) {
            const uint32_t n16 = plane_bytes >> 4;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % p.stages;
                const uint32_t ph = (uint32_t)(i / p.stages) & 1u;
                mbar_wait(full_bar(s), ph);
                const uint32_t raw = smem_base + (uint32_t)s * stage_bytes;
                // four independent 16-byte loads in flight per thread before the first store (the asm stores order memory)
                for (uint32_t j0 = t; j0 < n16; j0 += 4 * 128) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t j = j0 + 128u * u;
                        if (j < n16)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                         : "r"(raw + 16u * j));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t j = j0 + 128u * u;
                        if (j < n16) {
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(raw + plane_bytes + 16u * j), "f"(tf32_lo(v[u].x)),
                                         "f"(tf32_lo(v[u].y)), "f"(tf32_lo(v[u].z)), "f"(tf32_lo(v[u].w))
                                         : "memory");
                            if (p.write_hi)
                                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(raw + 16u * j), "f"(tf32_hi(v[u].x)),
                                             "f"(tf32_hi(v[u].y)), "f"(tf32_hi(v[u].z)), "f"(tf32_hi(v[u].w))
                                             : "memory");
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(xform_bar(s));
            }
        }
        mbar_wait(accum_bar, 0u);
        tc_fence_after();
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                            (!p.mask || (((p.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0)));
        const bool first_split = blockIdx.z == 0;
        if (p.filter) {
            // ---- f3: keep what beats the row's running k-th best; nothing is stored ----
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            const float tau = row_ok ? __ldcg(p.tau + row) : __int_as_float(0x7f800000);
            for (int c = 0; c < p.BN; c += 16) {
                uint32_t r[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
                if (!row_ok) continue;
                float mx = __uint_as_float(r[0]);
#pragma unroll
                for (int e = 1; e < 16; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
                if (!(mx > tau)) continue;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float sc = __uint_as_float(r[e]);
                    const int col = n0 + c + e;
                    if (sc > tau && col < p.N) {
                        const int pos = atomicAdd(p.count + row, 1);
                        if (pos < p.qcap) p.queue[(size_t)row * p.qcap + pos] = topk_key(sc, (uint32_t)(p.col_off + col));
                    }
                }
            }
        } else {
            // ---- a13: tcgen05.ld hands a thread 16 columns of ONE row (a warp store would touch 32 different lines); the
            // chunk is transposed through a per-warp shared-memory slab (the stage ring is idle by now) so that four lanes
            // cover one row's 64 bytes and a warp instruction writes 8 rows of whole sectors -- and reads the ReLU mask and the
            // bias the same way ----
            const uint32_t slab = smem_base + (uint32_t)(warp - 2) * (32u * 80u);       // 32 rows x 20 floats
            const int rsub = lane >> 2, cg = lane & 3;
            for (int c = 0; c < p.BN; c += 16) {
                uint32_t r[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(slab + (uint32_t)lane * 80u + 16u * g), "r"(r[4 * g]),
                                 "r"(r[4 * g + 1]), "r"(r[4 * g + 2]), "r"(r[4 * g + 3])
                                 : "memory");
                __syncwarp();
                const int col = n0 + c + 4 * cg;
                const bool col_ok = col < p.N;
                const bool full4 = vec_ok && col + 4 <= p.N;
                float bv[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                if (p.bias && first_split && col_ok) {
                    if (col + 4 <= p.N && ((reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)) {
                        const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                        bv[0] = t4.x; bv[1] = t4.y; bv[2] = t4.z; bv[3] = t4.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (col + e < p.N) bv[e] = __ldg(p.bias + col + e);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = 8 * i + rsub;
                    const int row = m0 + q * 32 + rr;
                    float v[4];
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                                 : "r"(slab + (uint32_t)rr * 80u + 16u * cg));
                    if (row >= p.M || !col_ok) continue;
                    float* crow = p.C + (int64_t)row * p.ldc;
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] += bv[e];
                    if (p.act == 1) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.0f);
                    }
                    if (p.mask) {
                        const float* mrow = p.mask + (int64_t)row * p.ldmask;
                        if (full4) {
                            const float4 mk = ld_stream_f4(mrow + col);
                            v[0] = mk.x > 0.0f ? v[0] : 0.0f;
                            v[1] = mk.y > 0.0f ? v[1] : 0.0f;
                            v[2] = mk.z > 0.0f ? v[2] : 0.0f;
                            v[3] = mk.w > 0.0f ? v[3] : 0.0f;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col + e < p.N) v[e] = mrow[col + e] > 0.0f ? v[e] : 0.0f;
                        }
                    }
                    if (p.atomic) {
                        if (full4) red_add_f4(crow + col, make_float4(v[0], v[1], v[2], v[3]));
                        else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (col + e < p.N) red_add_f1(crow + col + e, v[e]);
                        }
                    } else if (full4) {
                        *reinterpret_cast<float4*>(crow + col) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (col + e < p.N) crow[col + e] = v[e];
                    }
                }
                __syncwarp();                              // the slab is rewritten by the next chunk
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ---- skinny shapes (the Linear(hidden, 1) head of the MLP): memory-bound, no tensor cores ------------------------------
// C[M,N] = act(A[M,K] B[N,K]^T + bias), N <= 8: one warp per row, 16-byte streaming loads of the row
template <int NMAX>
__global__ void __launch_bounds__(256) k_skinny_nt(const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
                                                   float* __restrict__ C, int64_t ldc, int M, int N, int K, const float* __restrict__ bias,
                                                   int act, int accumulate) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool vec = ((lda | ldb) & 3) == 0 && (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0;
    for (int64_t m = warp0; m < M; m += nw) {
        float acc[NMAX];
#pragma unroll
        for (int n = 0; n < NMAX; ++n) acc[n] = 0.0f;
        const float* a = A + m * lda;
        if (vec) {
            for (int k = 4 * lane; k < K; k += 128) {
                const float4 x = ld_stream_f4(a + k);
#pragma unroll
                for (int n = 0; n < NMAX; ++n)
                    if (n < N) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(B + n * ldb + k));
                        acc[n] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[n]))));
                    }
            }
        } else {
            for (int k = lane; k < K; k += 32) {
                const float x = ld_stream_f1(a + k);
#pragma unroll
                for (int n = 0; n < NMAX; ++n)
                    if (n < N) acc[n] = fmaf(x, __ldg(B + n * ldb + k), acc[n]);
            }
        }
#pragma unroll
        for (int n = 0; n < NMAX; ++n) acc[n] = group_sum<32>(acc[n]);
        if (lane == 0) {
            for (int n = 0; n < N; ++n) {
                float v = acc[n] + (bias ? bias[n] : 0.0f);
                if (act == 1) v = fmaxf(v, 0.0f);
                if (accumulate) v += C[m * ldc + n];
                C[m * ldc + n] = v;
            }
        }
    }
}

// C[M,N] = (A[M,K] Bt[K,N]) * (mask > 0), K <= 8 (rank-K update: dH = dy w for the head): a thread owns 4 columns of a row
__global__ void __launch_bounds__(256) k_skinny_k(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bt, int64_t ldb,
                                                  float* __restrict__ C, int64_t ldc, int64_t M, int N, int K, const float* __restrict__ mask,
                                                  int64_t ldmask, int accumulate) {
    const int n4 = (N + 3) >> 2;
    const bool vec = (N & 3) == 0 && ((ldb | ldc) & 3) == 0 && (!mask || (ldmask & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(Bt) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0;
    const int64_t total = M * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / n4;
        const int n = 4 * (int)(i - m * n4);
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (int k = 0; k < K; ++k) {
            const float a = __ldg(A + m * lda + k);
            if (vec) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(Bt + k * ldb + n));
                v[0] = fmaf(a, w.x, v[0]); v[1] = fmaf(a, w.y, v[1]); v[2] = fmaf(a, w.z, v[2]); v[3] = fmaf(a, w.w, v[3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (n + e < N) v[e] = fmaf(a, __ldg(Bt + k * ldb + n + e), v[e]);
            }
        }
        if (vec) {
            if (mask) {
                const float4 mk = ld_stream_f4(mask + m * ldmask + n);
                v[0] = mk.x > 0.0f ? v[0] : 0.0f; v[1] = mk.y > 0.0f ? v[1] : 0.0f;
                v[2] = mk.z > 0.0f ? v[2] : 0.0f; v[3] = mk.w > 0.0f ? v[3] : 0.0f;
            }
            float4* c = reinterpret_cast<float4*>(C + m * ldc + n);
            if (accumulate) {
                const float4 o = *c;
                v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
            }
            *c = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (n + e < N) {
                    float x = v[e];
                    if (mask) x = mask[m * ldmask + n + e] > 0.0f ? x : 0.0f;
                    if (accumulate) x += C[m * ldc + n + e];
                    C[m * ldc + n + e] = x;
                }
        }
    }
}

// C[M,N] += At[Kr,M]^T Bt[Kr,N], M <= 8 (dw = dy^T H for the head): rows of the reduction are split over the CTAs, four rows
// in flight per thread
template <int MMAX>
__global__ void __launch_bounds__(256) k_skinny_m(const float* __restrict__ At, int64_t lda, const float* __restrict__ Bt, int64_t ldb,
                                                  float* __restrict__ C, int64_t ldc, int M, int N, int64_t Kr, int64_t rows_per_cta) {
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    int64_t r1 = r0 + rows_per_cta;
    if (r1 > Kr) r1 = Kr;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float acc[MMAX];
#pragma unroll
        for (int m = 0; m < MMAX; ++m) acc[m] = 0.0f;
        int64_t r = r0;
        for (; r + 4 <= r1; r += 4) {
            float x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = ld_stream_f1(Bt + (r + u) * ldb + n);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int m = 0; m < MMAX; ++m)
                    if (m < M) acc[m] = fmaf(__ldg(At + (r + u) * lda + m), x[u], acc[m]);
        }
        for (; r < r1; ++r) {
            const float x = ld_stream_f1(Bt + r * ldb + n);
#pragma unroll
            for (int m = 0; m < MMAX; ++m)
                if (m < M) acc[m] = fmaf(__ldg(At + r * lda + m), x, acc[m]);
        }
        for (int m = 0; m < M; ++m) red_add_f1(C + m * ldc + n, acc[m]);
    }
}

// out[n] (+)= sum_m X[m, n]  (bias gradient: column sums of dZ)
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ X, int64_t ldx, float* __restrict__ out, int64_t M, int N,
                                                int64_t rows_per_cta) {
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
    int64_t r1 = r0 + rows_per_cta;
    if (r1 > M) r1 = M;
    const int n = blockIdx.x * 64 + (threadIdx.x & 63), sub = threadIdx.x >> 6;        // 4 row phases x 64 columns
    __shared__ float part[4][64];
    float acc = 0.0f;
    if (n < N)
        for (int64_t r = r0 + sub; r < r1; r += 4) acc += ld_stream_f1(X + r * ldx + n);
    part[sub][threadIdx.x & 63] = acc;
    __syncthreads();
    if (sub == 0 && n < N) red_add_f1(out + n, part[0][threadIdx.x] + part[1][threadIdx.x] + part[2][threadIdx.x] + part[3][threadIdx.x]);
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2D fp32 tensor map over a row-major [outer, inner] view with row stride ld (elements); box = [box_outer, box_inner]
int make_map(CUtensorMap* m, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer,
             CUtensorMapSwizzle sw, const char* who) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return rbx_fail(RBX_ERR_CUDA, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return rbx_fail(RBX_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld %lld box [%d x %d]", who, (int)r,
                                           (long long)outer, (long long)inner, (long long)ld, box_outer, box_inner);
    return RBX_OK;
}

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

// Fills tile shape, descriptors and tensor maps of `g` (epilogue fields already set) and launches k_gemm_tc.
// bn_force / kb_force / max_stages = 0: chosen here.  allow_split: split-K (red.add epilogue) when the output has too few tiles.
int launch_tc(const char* who, GemmArgs& g, const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
              int bn_force, int kb_force, int max_stages, bool allow_split, int accumulate, cudaStream_t st) {
    const int sms = rbx_sm_count();
    const int precision = g.prec;
    g.M = (int)M; g.N = (int)N; g.K = (int)K;
    g.write_hi = env_int("RBX_GEMM_WRITE_HI", 0);
    // tile width: the fewest tiles that cover N, then the narrowest such tile (N = 400 -> 2 x 208; 624 -> 3 x 208)
    const int gran = g.b_mn ? 32 : 16;
    int tiles_n = (int)((N + 255) / 256);
    int BN = (int)(((N + tiles_n - 1) / tiles_n + gran - 1) / gran * gran);
    if (BN < 16) BN = 16;
    if (bn_force) BN = bn_force;
    if (int f = env_int("RBX_GEMM_BN", 0)) BN = f;
    RBX_REQUIRE(BN >= 16 && BN <= 256 && BN % gran == 0, "%s: tile width %d not supported", who, BN);
    tiles_n = (int)((N + BN - 1) / BN);
    const int tiles_m = (int)((M + kBM - 1) / kBM);
    RBX_REQUIRE(tiles_m <= 65535, "%s: M too large for one launch", who);
    g.BN = BN;
    int KB = kb_force ? kb_force : env_int("RBX_GEMM_KB", 0);
    if (KB != 16 && KB != 32) KB = (precision == 3 && BN > 128) ? 16 : 32;
    g.KB = KB;
    g.kb_total = (int)((K + KB - 1) / KB);
    // split-K: when the output has too few tiles to fill the machine (dW: reduction over the batch), and always so that one
    // TMEM accumulation chain stays at or below 1024 products (the tensor core's fp32 adds are not round-to-nearest: at K = 65 536
    // one chain of 5461 products per output drifts to 4.6e-5 of max|C|, 2048 to 1.4e-5; chains of <= 1024 combined by fp32
    // red.add stay within 1e-5)
    int splits = 1;
    if (allow_split) {
        const int tiles = tiles_m * tiles_n, slots = 2 * sms;            // two CTAs per SM (see the stage count below)
        int want = (int)((K + 1023) / 1024);                             // chains of <= 1024 products
        if (tiles * 2 <= sms && want < slots / tiles) want = slots / tiles;
        if (want > 1) {                                                  // whole waves: 12 tiles x 74 splits = 3 x 296 CTAs
            const int waves = (want * tiles + slots - 1) / slots;
            splits = waves * slots / tiles;
        }
        if (splits > g.kb_total / 4) splits = g.kb_total / 4;
        if (splits < 1) splits = 1;
    }
    if (int f = env_int("RBX_GEMM_SPLITS", 0)) splits = allow_split ? f : 1;
    g.kb_per_split = (g.kb_total + splits - 1) / splits;
    splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
    g.atomic = (splits > 1 || accumulate) ? 1 : 0;
    if (splits > 1 && !accumulate) {
        cudaError_t e = cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
        if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: memset: %s", who, cudaGetErrorString(e));
    }
    g.a_bytes = (uint32_t)(kBM * KB * 4);
    g.b_bytes = (uint32_t)(BN * KB * 4);
    const uint32_t stage_bytes = (precision == 3 ? 2u : 1u) * (g.a_bytes + g.b_bytes);
    const uint32_t bar_bytes = 8u * (3 * kMaxStages + 2);
    // two CTAs per SM when two stages of each fit (one CTA's prologue / epilogue then runs under the other's MMAs: measured
    // 225 us against 280 us with one 5-stage CTA on the 65 536 x 400 x 624 layer); otherwise one CTA with every stage that fits
    int stages = (int)((113u * 1024u - 1024u - bar_bytes) / stage_bytes);
    if (stages < 2) stages = (int)((227u * 1024u - 1024u - bar_bytes) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (max_stages && stages > max_stages) stages = max_stages;
    if (int f = env_int("RBX_GEMM_STAGES", 0)) stages = f < stages ? f : stages;
    if (stages > g.kb_per_split) stages = g.kb_per_split < 2 ? 2 : g.kb_per_split;      // no deeper than the k loop
    RBX_REQUIRE(stages >= 2, "%s: tile does not fit shared memory", who);
    g.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + bar_bytes + 1024;
    uint32_t cols = 32;
    while (cols < (uint32_t)BN) cols <<= 1;
    g.tmem_cols = cols;
    g.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)g.a_mn << 15) | ((uint32_t)g.b_mn << 16) | ((uint32_t)(BN >> 3) << 17) |
              ((uint32_t)(kBM >> 4) << 24);
    // shared-memory matrix descriptors (version 1 in bits [46,48), swizzle mode in bits [61,64))
    const CUtensorMapSwizzle k_sw = KB == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const uint32_t k_layout = KB == 32 ? 2u : 4u;
    auto desc_hi = [](uint32_t sbo_bytes, uint32_t layout) { return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | (layout << 29); };
    // MN-major tf32 operands have ONE legal shared-memory layout: 128-byte rows swizzled in 32-byte chunks (descriptor
    // layout type 1 = SWIZZLE_128B_BASE32B, TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); its atom is 32 elements (MN) x 4 k-rows.
    if (g.a_mn) {       // [KB k-rows][32 m] boxes, 128 B rows: LBO = box pitch (next 32 m), SBO = 4 k-rows = 512 B
        g.a_boxes = kBM / 32; g.a_box_bytes = (uint32_t)(KB * 128);
        g.a_desc_hi = desc_hi(512, 1); g.a_lbo_sbo = ((g.a_box_bytes >> 4) & 0x3fffu) << 16; g.a_kstep = 1024 >> 4;
        if (int rc = make_map(&g.ta, A, M, K, lda, 32, KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, who)) return rc;
    } else {            // [128 rows][KB k] one box, rows of KB*4 B: SBO = 8 rows
        g.a_boxes = 1; g.a_box_bytes = g.a_bytes;
        g.a_desc_hi = desc_hi((uint32_t)(8 * KB * 4), k_layout); g.a_lbo_sbo = 1u << 16; g.a_kstep = 32 >> 4;
        if (int rc = make_map(&g.ta, A, K, M, lda, KB, kBM, k_sw, who)) return rc;
    }
    if (g.b_mn) {
        g.b_boxes = (uint32_t)(BN / 32); g.b_box_bytes = (uint32_t)(KB * 128);
        g.b_desc_hi = desc_hi(512, 1); g.b_lbo_sbo = ((g.b_box_bytes >> 4) & 0x3fffu) << 16; g.b_kstep = 1024 >> 4;
        if (int rc = make_map(&g.tb, B, N, K, ldb, 32, KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, who)) return rc;
    } else {
        g.b_boxes = 1; g.b_box_bytes = g.b_bytes;
        g.b_desc_hi = desc_hi((uint32_t)(8 * KB * 4), k_layout); g.b_lbo_sbo = 1u << 16; g.b_kstep = 32 >> 4;
        if (int rc = make_map(&g.tb, B, K, N, ldb, KB, BN, k_sw, who)) return rc;
    }
    static bool attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    k_gemm_tc<<<dim3(tiles_n, tiles_m, splits), kThreads, smem, st>>>(g);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

// =====================================================================================================================
// f3 pass A, persistent form: one CTA per SM keeps its 128-user tile (raw + lo planes, every k-block) resident in shared
// memory and streams 256-item tiles through a stage ring; two 256-column TMEM accumulators alternate, so the filter epilogue of
// tile i runs under the MMAs of tile i + 1, and the per-tile fixed costs of the one-tile-per-CTA kernel (barrier init, TMEM
// allocation, first TMA round trip, re-loading and re-splitting the user tile: ~4 us against 1.6 us of MMAs at D = 64) are
// paid once per CTA.  10 warps: 0 TMA producer, 1 MMA issuer, 2..5 transform (lo planes), 6..9 filter epilogue.
// =====================================================================================================================
constexpr int kTkThreads = 320;
constexpr int kTkBN = 256, kTkKB = 16;
constexpr uint32_t kTkABlk = kBM * kTkKB * 4;       // 8 KB: one k-block of the user tile
constexpr uint32_t kTkBBlk = kTkBN * kTkKB * 4;     // 16 KB: one k-block of an item tile

struct TopkArgs {
    CUtensorMap ta, tb;
    int M, N, nkb, stages, tiles_n;
    uint32_t idesc, desc_hi;
    const float* tau;
    int* count;
    unsigned long long* queue;
    int64_t qcap, col_off;
};

__global__ void __launch_bounds__(kTkThreads, 1) k_topk_tc(const __grid_constant__ TopkArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_raw = smem_base, a_lo = a_raw + (uint32_t)p.nkb * kTkABlk;
    const uint32_t b_ring = a_lo + (uint32_t)p.nkb * kTkABlk;
    const uint32_t stage_bytes = 2 * kTkBBlk;                                    // raw | lo
    const uint32_t bars = b_ring + (uint32_t)p.stages * stage_bytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto xform_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
    auto empty_bar = [&](int s) { return bars + 8u * (2 * kMaxStages + s); };
    const uint32_t a_full = bars + 8u * (3 * kMaxStages), a_ready = a_full + 8u;
    auto acc_full = [&](int b) { return a_ready + 8u + 8u * b; };
    auto acc_empty = [&](int b) { return a_ready + 24u + 8u * b; };
    const uint32_t tmem_slot = a_ready + 40u;
    const int m0 = blockIdx.y * kBM;
    int my_tiles = 0;
    if ((int)blockIdx.x < p.tiles_n) my_tiles = (p.tiles_n - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(xform_bar(s), 4);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_ready, 4);
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full(b), 1);
            mbar_init(acc_empty(b), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===== TMA producer: the user tile once, then the item tiles =====
        if (lane == 0 && my_tiles > 0) {
            mbar_arrive_expect_tx(a_full, (uint32_t)p.nkb * kTkABlk);
            for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(a_raw + (uint32_t)kb * kTkABlk, &p.ta, a_full, kb * kTkKB, m0);
            int it = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int n0 = ((int)blockIdx.x + j * (int)gridDim.x) * kTkBN;
                for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_arrive_expect_tx(full_bar(s), kTkBBlk);
                    tma_load_2d(b_ring + (uint32_t)s * stage_bytes, &p.tb, full_bar(s), kb * kTkKB, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && my_tiles > 0) {
            mbar_wait(a_ready, 0u);
            tc_fence_after();
            int it = 0;
            for (int j = 0; j < my_tiles; ++j) {
                const int buf = j & 1;
                mbar_wait(acc_empty(buf), ((uint32_t)(j >> 1) & 1u) ^ 1u);       // the epilogue drained this accumulator
                tc_fence_after();
                const uint32_t acc_addr = tmem_base + (uint32_t)buf * kTkBN;
                uint32_t acc = 0;
                for (int kb = 0; kb < p.nkb; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                    mbar_wait(xform_bar(s), ph);
                    tc_fence_after();
                    const uint32_t b_hi = b_ring + (uint32_t)s * stage_bytes, b_lo = b_hi + kTkBBlk;
                    const uint32_t ah = a_raw + (uint32_t)kb * kTkABlk, al = a_lo + (uint32_t)kb * kTkABlk;
#pragma unroll
                    for (int k = 0; k < kTkKB / 8; ++k) {
                        const uint64_t hi64 = (uint64_t)p.desc_hi << 32;
                        const uint64_t da_hi = hi64 | ((1u << 16) | (((ah >> 4) + 2u * k) & 0x3fffu));
                        const uint64_t da_lo = hi64 | ((1u << 16) | (((al >> 4) + 2u * k) & 0x3fffu));
                        const uint64_t db_hi = hi64 | ((1u << 16) | (((b_hi >> 4) + 2u * k) & 0x3fffu));
                        const uint64_t db_lo = hi64 | ((1u << 16) | (((b_lo >> 4) + 2u * k) & 0x3fffu));
                        tc_mma_tf32(acc_addr, da_lo, db_hi, p.idesc, acc);
                        tc_mma_tf32(acc_addr, da_hi, db_lo, p.idesc, 1u);
                        tc_mma_tf32(acc_addr, da_hi, db_hi, p.idesc, 1u);
                        acc = 1u;
                    }
                    tc_commit(empty_bar(s));
                }
                tc_commit(acc_full(buf));
            }
        }
    } else if (warp < 6) {
        // ===== transform warps: lo plane of the user tile once, then of every item stage =====
        const int t = threadIdx.x - 64;          // 0..127
        if (my_tiles > 0) {
            auto split_plane = [&](uint32_t raw, uint32_t lo, uint32_t n16) {
                for (uint32_t j0 = t; j0 < n16; j0 += 4 * 128) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t j = j0 + 128u * u;
                        if (j < n16)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                         : "r"(raw + 16u * j));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t j = j0 + 128u * u;
                        if (j < n16)
                            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo + 16u * j), "f"(tf32_lo(v[u].x)), "f"(tf32_lo(v[u].y)),
                                         "f"(tf32_lo(v[u].z)), "f"(tf32_lo(v[u].w))
                                         : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            };
            mbar_wait(a_full, 0u);
            split_plane(a_raw, a_lo, ((uint32_t)p.nkb * kTkABlk) >> 4);
            if (lane == 0) mbar_arrive(a_ready);
            const int total = my_tiles * p.nkb;
            for (int it = 0; it < total; ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
                mbar_wait(full_bar(s), ph);
                const uint32_t raw = b_ring + (uint32_t)s * stage_bytes;
                split_plane(raw, raw + kTkBBlk, kTkBBlk >> 4);
                if (lane == 0) mbar_arrive(xform_bar(s));
            }
        }
    } else {
        // ===== filter epilogue: a thread owns one user row of the tile =====
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < p.M;
        const float tau = row_ok ? __ldcg(p.tau + row) : __int_as_float(0x7f800000);
        for (int j = 0; j < my_tiles; ++j) {
            const int buf = j & 1;
            const int n0 = ((int)blockIdx.x + j * (int)gridDim.x) * kTkBN;
            mbar_wait(acc_full(buf), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * kTkBN;
            for (int c = 0; c < kTkBN; c += 16) {
                uint32_t r[16];
                tc_ld16(taddr + (uint32_t)c, r);
                if (!row_ok) continue;
                // survivors are rare once tau has tightened (~k * chunk / seen per user and pass): one max over the 16 scores
                // decides whether the slow path runs at all
                float mx = __uint_as_float(r[0]);
#pragma unroll
                for (int e = 1; e < 16; ++e) mx = fmaxf(mx, __uint_as_float(r[e]));
                if (!(mx > tau)) continue;
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float sc = __uint_as_float(r[e]);
                    const int col = n0 + c + e;
                    if (sc > tau && col < p.N) {
                        const int pos = atomicAdd(p.count + row, 1);
                        if (pos < p.qcap) p.queue[(size_t)row * p.qcap + pos] = topk_key(sc, (uint32_t)(p.col_off + col));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(buf));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

// f3 pass A on the tensor cores (called from csrc/topk.cu): scores of users [0, U) against items [n0, n1) as a 3xTF32 GEMM
// whose epilogue keeps only what beats the user's running k-th best.  Default: the persistent kernel above; RBX_TOPK_PERSIST=0
// (or a chunk of fewer item tiles than CTAs) runs the one-tile-per-CTA GEMM with the filter epilogue.
int rbx_topk_filter_tc(const float* q, const float* items, int64_t U, int64_t n0, int64_t n1, int D, const float* tau, int* count,
                       unsigned long long* queue, int64_t qcap, cudaStream_t st) {
    const char* who = "rbx_topk_ip";
    const int64_t N = n1 - n0;
    const int tiles_n = (int)((N + kTkBN - 1) / kTkBN), tiles_m = (int)((U + kBM - 1) / kBM);
    const int sms = rbx_sm_count();
    if (env_int("RBX_TOPK_PERSIST", 1) && tiles_m <= 65535 && (int64_t)tiles_n * tiles_m >= 2 * sms) {
        TopkArgs a;
        memset(&a, 0, sizeof(a));
        a.M = (int)U; a.N = (int)N; a.nkb = (D + kTkKB - 1) / kTkKB; a.tiles_n = tiles_n;
        a.tau = tau; a.count = count; a.queue = queue; a.qcap = qcap; a.col_off = n0;
        const uint32_t bar_bytes = 8u * (3 * kMaxStages + 8);
        const uint32_t fixed = 2u * (uint32_t)a.nkb * kTkABlk + bar_bytes + 1024u;
        int stages = (int)((227u * 1024u - fixed) / (2 * kTkBBlk));
        if (stages > kMaxStages) stages = kMaxStages;
        RBX_REQUIRE(stages >= 2, "%s: D=%d does not fit the persistent top-k kernel", who, D);
        a.stages = stages;
        a.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTkBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        a.desc_hi = ((uint32_t)(8 * kTkKB * 4) >> 4) | (1u << 14) | (4u << 29);              // K-major, 64-byte swizzle, SBO = 8 rows
        if (int rc = make_map(&a.ta, q, D, U, D, kTkKB, kBM, CU_TENSOR_MAP_SWIZZLE_64B, who)) return rc;
        if (int rc = make_map(&a.tb, items + (size_t)n0 * D, D, N, D, kTkKB, kTkBN, CU_TENSOR_MAP_SWIZZLE_64B, who)) return rc;
        static bool attr_set[64];
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(k_topk_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
            attr_set[dev] = true;
        }
        int gx = sms / tiles_m;                     // one CTA per SM; user tiles of the same item range run side by side (L2)
        if (gx < 1) gx = 1;
        if (gx > tiles_n) gx = tiles_n;
        const size_t smem = fixed + (size_t)stages * 2 * kTkBBlk;
        k_topk_tc<<<dim3(gx, tiles_m), kTkThreads, smem, st>>>(a);
        RBX_LAUNCH_CHECK(who);
        return RBX_OK;
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.prec = 3;
    g.filter = 1; g.tau = tau; g.count = count; g.queue = queue; g.qcap = qcap; g.col_off = n0;
    return launch_tc(who, g, q, D, items + (size_t)n0 * D, D, U, n1 - n0, D, 256, 16, 2, false, 0, st);
}

extern "C" {

int rbx_gemm_f32(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, float* C, int64_t ldc, int64_t M, int64_t N,
                 int64_t K, const float* bias, int act, const float* mask, int64_t ldmask, int precision, int accumulate,
                 rbx_stream_t stream) {
    const char* who = "rbx_gemm_f32";
    RBX_RANGE("rbx_gemm_f32");
    RBX_REQUIRE(M >= 0 && N >= 0 && K >= 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "%s: bad shape", who);
    if (M == 0 || N == 0) return RBX_OK;
    RBX_REQUIRE(A && B && C, "%s: null operand", who);
    RBX_REQUIRE(precision == 1 || precision == 3, "%s: precision must be 1 (TF32) or 3 (3xTF32, fp32-level)", who);
    RBX_REQUIRE(act == 0 || act == 1, "%s: act must be 0 (none) or 1 (ReLU)", who);
    RBX_REQUIRE(ldc >= N && (!mask || ldmask >= N), "%s: ldc / ldmask shorter than a row", who);
    cudaStream_t st = rbx_cast_stream(stream);
    const int sms = rbx_sm_count();
    // ---- skinny shapes: SIMT, memory-bound ----
    if (N <= 8 && !a_mn && !b_mn && !mask) {
        RBX_REQUIRE(lda >= K && ldb >= K, "%s: leading dimension shorter than a row", who);
        int grid = (int)((M + 7) / 8);
        if (grid > sms * 8) grid = sms * 8;
        k_skinny_nt<8><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, (int)M, (int)N, (int)K, bias, act, accumulate);
        RBX_LAUNCH_CHECK(who);
        return RBX_OK;
    }
    if (K <= 8 && !a_mn && b_mn && !bias && act == 0) {
        int64_t grid = (M * ((N + 3) / 4) + 255) / 256;
        if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
        k_skinny_k<<<(unsigned)grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, (int)N, (int)K, mask, ldmask, accumulate);
        RBX_LAUNCH_CHECK(who);
        return RBX_OK;
    }
    if (M <= 8 && a_mn && b_mn && !bias && act == 0 && !mask) {
        if (!accumulate) {
            cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
            if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: memset: %s", who, cudaGetErrorString(e));
        }
        const int bt = N >= 256 ? 256 : (N >= 128 ? 128 : 64);         // N = 400: 4 x 128 threads cover the columns, 112 idle
        const int gx = (int)((N + bt - 1) / bt);
        int gy = sms * 8 / gx;
        if (gy < 1) gy = 1;
        if (gy > K) gy = (int)(K > 0 ? K : 1);
        const int64_t per = (K + gy - 1) / gy;
        if (K > 0) k_skinny_m<8><<<dim3(gx, gy), bt, 0, st>>>(A, lda, B, ldb, C, ldc, (int)M, (int)N, K, per);
        RBX_LAUNCH_CHECK(who);
        return RBX_OK;
    }
    // ---- tensor-core path ----
    RBX_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0, "%s: lda / ldb must be multiples of 4 floats (TMA row pitch is 16-byte granular)", who);
    RBX_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0, "%s: A / B must be 16-byte aligned", who);
    RBX_REQUIRE(lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K), "%s: leading dimension shorter than a row", who);
    if (K == 0) {
        RBX_REQUIRE(!bias && act == 0, "%s: K = 0 with an epilogue is not supported", who);
        if (!accumulate) {
            cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
            if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: memset: %s", who, cudaGetErrorString(e));
        }
        return RBX_OK;
    }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.C = C; g.bias = bias; g.mask = mask; g.ldc = ldc; g.ldmask = ldmask;
    g.a_mn = a_mn ? 1 : 0; g.b_mn = b_mn ? 1 : 0; g.prec = precision; g.act = act;
    return launch_tc(who, g, A, lda, B, ldb, M, N, K, 0, 0, 0, act == 0 && !mask, accumulate, st);
}

int rbx_colsum_f32(const float* X, int64_t ldx, float* out, int64_t M, int64_t N, int accumulate, rbx_stream_t stream) {
    const char* who = "rbx_colsum_f32";
    RBX_RANGE("rbx_colsum_f32");
    RBX_REQUIRE(M >= 0 && N >= 0 && N < (1ll << 31), "%s: bad shape", who);
    if (N == 0) return RBX_OK;
    RBX_REQUIRE(X && out && ldx >= N, "%s: bad argument", who);
    cudaStream_t st = rbx_cast_stream(stream);
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(out, 0, (size_t)N * 4, st);
        if (e != cudaSuccess) return rbx_fail(RBX_ERR_CUDA, "%s: memset: %s", who, cudaGetErrorString(e));
    }
    if (M == 0) return RBX_OK;
    const int gx = (int)((N + 63) / 64);
    int gy = rbx_sm_count() * 8 / gx;
    if (gy < 1) gy = 1;
    if (gy > M) gy = (int)M;
    const int64_t per = (M + gy - 1) / gy;
    k_colsum<<<dim3(gx, gy), 256, 0, st>>>(X, ldx, out, M, (int)N, per);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
