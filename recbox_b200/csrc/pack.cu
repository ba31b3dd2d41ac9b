// a1: batch matrix / dict-of-columns -> dense slot blocks ([B,F] int32 global rows, [B,Fn] fp32,
// [B] fp32 label).  One coalesced pass; replaces 39 column slices + 39 H2D copies + 26 .long()
// casts per batch (ranking_model.py:106-122, feature_embedding.py:201,204).
#include "rbx_common.cuh"

namespace {

constexpr int kMaxCols = 512;

struct SplitParams {
    const double* batch;
    int32_t* rows;
    float* dense_x;
    float* label;
    int64_t B, ld;
    int n_cols, F, Fn;
    int32_t* n_bad;              // [1] device counter of ids outside their slot's vocabulary | NULL
    int32_t col_add[kMaxCols];   // field offset of the column's slot (kind 1)
    int32_t col_rows[kMaxCols];  // vocabulary size of the column's slot (kind 1), 0 = unchecked
    int16_t col_slot[kMaxCols];
    int8_t col_kind[kMaxCols];
};

// A CTA takes blocks of kSplitRows consecutive rows and walks their elements in memory order (thread <-> element, so
// every lane of every load is busy and a warp reads 256 contiguous bytes); row / column of an element come from one
// 32-bit division.  The per-column metadata is staged in shared memory once per CTA: indexing the __grid_constant__
// arrays with a lane-dependent column serialises on the constant cache (the first version of this kernel did that, plus
// a 64-bit division per element, and ran at 0.08 of the HBM roofline; a warp-per-row version idled 24 of 32 lanes on the
// second pass over a 40-column row: 0.19).
constexpr int kSplitRows = 32;
__global__ void __launch_bounds__(256) k_split_batch(const __grid_constant__ SplitParams p) {
    __shared__ int32_t s_add[kMaxCols];
    __shared__ int32_t s_rows[kMaxCols];
    __shared__ int16_t s_slot[kMaxCols];
    __shared__ int8_t s_kind[kMaxCols];
    for (int c = threadIdx.x; c < p.n_cols; c += blockDim.x) {
        s_add[c] = p.col_add[c];
        s_rows[c] = p.col_rows[c];
        s_slot[c] = p.col_slot[c];
        s_kind[c] = p.col_kind[c];
    }
    __syncthreads();
    const uint32_t n_cols = (uint32_t)p.n_cols;
    for (int64_t b0 = (int64_t)blockIdx.x * kSplitRows; b0 < p.B; b0 += (int64_t)gridDim.x * kSplitRows) {
        const uint32_t rows_here = (uint32_t)(p.B - b0 < kSplitRows ? p.B - b0 : kSplitRows);
        const uint32_t n = rows_here * n_cols;
#pragma unroll 4
        for (uint32_t e = threadIdx.x; e < n; e += 256) {
            const uint32_t r = e / n_cols, c = e - r * n_cols;
            const int kind = s_kind[c];
            if (kind == 0) continue;
            const int64_t b = b0 + r;
            const double v = __ldcs(p.batch + b * p.ld + c);
            const int s = s_slot[c];
            if (kind == 1) {
                const int64_t id = (int64_t)v;                          // .long(): truncation toward zero
                int32_t row = (int32_t)id + s_add[c];
                if (s_rows[c] > 0 && (id < 0 || id >= s_rows[c])) {     // nn.Embedding would raise: never alias a neighbour's rows
                    row = -1;
                    if (p.n_bad) atomicAdd(p.n_bad, 1);
                }
                p.rows[b * p.F + s] = row;
            } else if (kind == 2) {
                p.dense_x[b * p.Fn + s] = (float)v;                     // .float(): round to nearest
            } else {
                p.label[b] = (float)v;
            }
        }
    }
}

constexpr int kPackCols = 64;
struct PackParams {
    const void* col[kPackCols];
    int64_t stride[kPackCols];
    int64_t add[kPackCols];
    int64_t rows[kPackCols];     // vocabulary size per column (as_rows), 0 = unchecked
    int8_t dtype[kPackCols];
    void* out;
    int32_t* n_bad;
    int64_t B;
    int n_total, c0, n_here, as_rows;
};

__device__ __forceinline__ double load_any(const void* p, int64_t i, int dtype) {
    switch (dtype) {
        case 0: return reinterpret_cast<const double*>(p)[i];
        case 1: return (double)reinterpret_cast<const float*>(p)[i];
        case 2: return (double)reinterpret_cast<const int64_t*>(p)[i];
        default: return (double)reinterpret_cast<const int32_t*>(p)[i];
    }
}

// 32 samples x n_here columns per CTA, transposed through shared memory so both the column reads
// (along b) and the block writes (along the slot axis) are coalesced.
__global__ void __launch_bounds__(256) k_pack_columns(const __grid_constant__ PackParams p) {
    __shared__ int32_t tile[32][kPackCols + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int64_t b0 = (int64_t)blockIdx.x * 32; b0 < p.B; b0 += (int64_t)gridDim.x * 32) {
        const int64_t b = b0 + tx;
        for (int c = ty; c < p.n_here; c += 8) {
            int32_t bits = 0;
            if (b < p.B) {
                const int dt = p.dtype[c];
                if (p.as_rows) {
                    int64_t id;
                    if (dt == 2) id = reinterpret_cast<const int64_t*>(p.col[c])[b * p.stride[c]];
                    else if (dt == 3) id = reinterpret_cast<const int32_t*>(p.col[c])[b * p.stride[c]];
                    else id = (int64_t)load_any(p.col[c], b * p.stride[c], dt);
                    bits = (int32_t)(id + p.add[c]);
                    if (p.rows[c] > 0 && (id < 0 || id >= p.rows[c])) {
                        bits = -1;
                        if (p.n_bad) atomicAdd(p.n_bad, 1);
                    }
                } else {
                    bits = __float_as_int((float)load_any(p.col[c], b * p.stride[c], dt));
                }
            }
            tile[tx][c] = bits;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int64_t bb = b0 + r;
            if (bb < p.B)
                for (int c = tx; c < p.n_here; c += 32)
                    reinterpret_cast<int32_t*>(p.out)[bb * p.n_total + p.c0 + c] = tile[r][c];
        }
        __syncthreads();
    }
}

// compact ids: uint16 local ids (every vocabulary < 65 536) -> int32 fused-table rows
__global__ void __launch_bounds__(256) k_unpack_ids_u16(const uint16_t* __restrict__ ids, int32_t* __restrict__ rows, int64_t n,
                                                        int F, const __grid_constant__ SplitParams p) {
    __shared__ int32_t s_add[kMaxCols];
    for (int c = threadIdx.x; c < F; c += blockDim.x) s_add[c] = p.col_add[c];
    __syncthreads();
    // 8 ids (16 bytes) per thread; F is arbitrary, so the slot of element e is e % F
    const int64_t n8 = ((uintptr_t)ids % 16 == 0 && (uintptr_t)rows % 16 == 0) ? n / 8 : 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4*>(ids) + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        int f = (int)((i * 8) % F);
        int32_t out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            out[j] = (int32_t)((w[j >> 1] >> (16 * (j & 1))) & 0xffffu) + s_add[f];
            f = f + 1 == F ? 0 : f + 1;
        }
        reinterpret_cast<int4*>(rows)[2 * i] = make_int4(out[0], out[1], out[2], out[3]);
        reinterpret_cast<int4*>(rows)[2 * i + 1] = make_int4(out[4], out[5], out[6], out[7]);
    }
    for (int64_t e = 8 * n8 + (int64_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (int64_t)gridDim.x * 256)
        rows[e] = (int32_t)ids[e] + s_add[e % F];
}

__global__ void __launch_bounds__(256) k_zero_f32(float* __restrict__ p, int64_t n) {
    const int64_t head = ((16 - (uintptr_t)p % 16) % 16) / 4;          // floats up to the first 16-byte boundary
    const int64_t h = head < n ? head : n;
    const int64_t n4 = (n - h) / 4;
    float4* p4 = reinterpret_cast<float4*>(p + h);
    const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x, nth = (int64_t)gridDim.x * 256;
    for (int64_t i = tid; i < n4; i += nth) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = tid; i < h; i += nth) p[i] = 0.f;
    for (int64_t i = h + 4 * n4 + tid; i < n; i += nth) p[i] = 0.f;
}

}  // namespace

extern "C" {

int rbx_unpack_ids_u16(const uint16_t* ids, int64_t B, int F, const int64_t* field_off, int32_t* rows, rbx_stream_t stream) {
    const char* who = "rbx_unpack_ids_u16";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && F >= 0 && F <= kMaxCols, "%s: bad shape", who);
    if (B == 0 || F == 0) return RBX_OK;
    RBX_REQUIRE(ids && rows, "%s: null pointer", who);
    SplitParams p;
    for (int f = 0; f < F; ++f) {
        const int64_t off = field_off ? field_off[f] : 0;
        RBX_REQUIRE(off >= 0 && off <= INT32_MAX - 65535, "%s: field_off[%d] outside int32", who, f);
        p.col_add[f] = (int32_t)off;
    }
    const int64_t n = B * F;
    int64_t ctas = (n / 8 + 255) / 256 + 1;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    k_unpack_ids_u16<<<(int)ctas, 256, 0, rbx_cast_stream(stream)>>>(ids, rows, n, F, p);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_zero_f32(float* ptr, int64_t n, rbx_stream_t stream) {
    const char* who = "rbx_zero_f32";
    RBX_RANGE(who);
    RBX_REQUIRE(n >= 0, "%s: negative size", who);
    if (n == 0) return RBX_OK;
    RBX_REQUIRE(ptr != nullptr, "%s: null pointer", who);
    int64_t ctas = (n / 4 + 255) / 256 + 1;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    k_zero_f32<<<(int)ctas, 256, 0, rbx_cast_stream(stream)>>>(ptr, n);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_split_batch_f64(const double* batch, int64_t B, int n_cols, int64_t ld, const int8_t* col_kind,
                        const int16_t* col_slot, const int64_t* field_off, const int64_t* field_rows, int F, int Fn,
                        int32_t* rows, float* dense_x, float* label, int32_t* n_bad, rbx_stream_t stream) {
    const char* who = "rbx_split_batch_f64";
    RBX_RANGE(who);
    RBX_REQUIRE(B >= 0 && n_cols >= 0 && ld >= n_cols, "%s: bad shape", who);
    RBX_REQUIRE(n_cols <= kMaxCols, "%s: n_cols=%d > %d", who, n_cols, kMaxCols);
    if (B == 0 || n_cols == 0) return RBX_OK;
    RBX_REQUIRE(batch && col_kind && col_slot, "%s: null pointer", who);
    SplitParams p;
    p.batch = batch; p.rows = rows; p.dense_x = dense_x; p.label = label; p.B = B; p.ld = ld;
    p.n_cols = n_cols; p.F = F; p.Fn = Fn; p.n_bad = n_bad;
    for (int c = 0; c < n_cols; ++c) {
        const int k = col_kind[c], s = col_slot[c];
        RBX_REQUIRE(k >= 0 && k <= 3, "%s: col_kind[%d]=%d", who, c, k);
        if (k == 1) {
            RBX_REQUIRE(rows && s >= 0 && s < F, "%s: column %d -> categorical slot %d of %d", who, c, s, F);
            const int64_t off = field_off ? field_off[s] : 0;
            RBX_REQUIRE(off >= 0 && off <= INT32_MAX, "%s: field_off[%d] outside int32", who, s);
            p.col_add[c] = (int32_t)off;
            const int64_t nr = field_rows ? field_rows[s] : 0;
            RBX_REQUIRE(nr >= 0 && nr <= INT32_MAX, "%s: field_rows[%d] outside int32", who, s);
            p.col_rows[c] = (int32_t)nr;
        } else {
            p.col_add[c] = 0;
            p.col_rows[c] = 0;
        }
        if (k == 2) RBX_REQUIRE(dense_x && s >= 0 && s < Fn, "%s: column %d -> numeric slot %d of %d", who, c, s, Fn);
        if (k == 3) RBX_REQUIRE(label != nullptr, "%s: label column without label output", who);
        p.col_kind[c] = (int8_t)k;
        p.col_slot[c] = (int16_t)s;
    }
    int64_t ctas = (B + kSplitRows - 1) / kSplitRows;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    k_split_batch<<<(int)ctas, 256, 0, rbx_cast_stream(stream)>>>(p);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_pack_columns(const void* const* cols, const int64_t* strides, const int8_t* dtypes, const int64_t* add,
                     const int64_t* vocab, int n, int64_t B, int as_rows, void* out, int32_t* n_bad, rbx_stream_t stream) {
    const char* who = "rbx_pack_columns";
    RBX_RANGE(who);
    RBX_REQUIRE(n >= 0 && B >= 0, "%s: negative size", who);
    if (n == 0 || B == 0) return RBX_OK;
    RBX_REQUIRE(cols && strides && dtypes && out, "%s: null pointer", who);
    for (int c0 = 0; c0 < n; c0 += kPackCols) {
        PackParams p;
        p.out = out; p.B = B; p.n_total = n; p.c0 = c0; p.as_rows = as_rows; p.n_bad = n_bad;
        p.n_here = (n - c0 < kPackCols) ? n - c0 : kPackCols;
        for (int c = 0; c < p.n_here; ++c) {
            RBX_REQUIRE(cols[c0 + c] != nullptr, "%s: column %d is null", who, c0 + c);
            RBX_REQUIRE(dtypes[c0 + c] >= 0 && dtypes[c0 + c] <= 3, "%s: dtype code %d", who, dtypes[c0 + c]);
            p.col[c] = cols[c0 + c];
            p.stride[c] = strides[c0 + c];
            p.dtype[c] = dtypes[c0 + c];
            p.add[c] = add ? add[c0 + c] : 0;
            p.rows[c] = (vocab && as_rows) ? vocab[c0 + c] : 0;
        }
        int64_t ctas = (B + 31) / 32;
        const int64_t cap = (int64_t)rbx_sm_count() * 8;
        if (ctas > cap) ctas = cap;
        k_pack_columns<<<(int)ctas, 256, 0, rbx_cast_stream(stream)>>>(p);
        RBX_LAUNCH_CHECK(who);
    }
    return RBX_OK;
}

}  // extern "C"
