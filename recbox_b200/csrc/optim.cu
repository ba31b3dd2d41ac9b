// a12: global-norm clip + dense Adam on the fused tables, device-resident end to end (no host sync:
// the clip coefficient is computed and consumed on the device).  Replaces
// nn.utils.clip_grad_norm_(self.parameters(), 10.) + torch.optim.Adam.step over the embedding tables
// (ranking_model.py:195-196, match_model.py:197-198).  Streaming, HBM-bound: 16 B read + 12 B
// written per element for Adam, 4 B read for the norm.
#include "rbx_common.cuh"

namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_sqnorm(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
    __shared__ float s_part[kThreads / 32];
    float acc = 0.f;
    const int64_t n4 = ((uintptr_t)g % 16 == 0) ? n / 4 : 0;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    for (int64_t i = tid; i < n4; i += nth) {
        const float4 v = ld_stream_f4(g + 4 * i);
        acc = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, acc))));
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nth) acc = fmaf(g[i], g[i], acc);
    acc = group_sum<32>(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += (double)s_part[w];
        atomicAdd(out, t);
    }
}

// coef = min(1, max_norm / (sqrt(sqnorm) + 1e-6))      (torch.nn.utils.clip_grad_norm_)
__global__ void k_clip_coef(const double* __restrict__ sqnorm, float max_norm, float* __restrict__ coef,
                            float* __restrict__ norm_out) {
    const float total = (float)sqrt(*sqnorm);
    const float c = max_norm / (total + 1e-6f);
    *coef = c < 1.f ? c : 1.f;
    if (norm_out) *norm_out = total;
}

struct AdamConst {
    float one_minus_b1, b2, one_minus_b2, neg_step_size, bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_elem(float& w, float g, float& m, float& v, const AdamConst& c) {
    m = m + c.one_minus_b1 * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = v * c.b2 + c.one_minus_b2 * g * g;                  // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) / c.bc2_sqrt + c.eps;     // (sqrt / sqrt(bc2)).add_(eps)
    w = w + c.neg_step_size * (m / denom);                  // addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(kThreads) k_adam_dense(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const float* __restrict__ clip,
                                                        AdamConst c, int vec_ok) {
    const float coef = clip ? __ldg(clip) : 1.f;
    const int64_t n4 = vec_ok ? n / 4 : 0;
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 W = *reinterpret_cast<float4*>(w + 4 * i);
        float4 G = ld_stream_f4(g + 4 * i);
        float4 M = *reinterpret_cast<float4*>(m + 4 * i);
        float4 V = *reinterpret_cast<float4*>(v + 4 * i);
        adam_elem(W.x, G.x * coef, M.x, V.x, c);
        adam_elem(W.y, G.y * coef, M.y, V.y, c);
        adam_elem(W.z, G.z * coef, M.z, V.z, c);
        adam_elem(W.w, G.w * coef, M.w, V.w, c);
        *reinterpret_cast<float4*>(w + 4 * i) = W;
        *reinterpret_cast<float4*>(m + 4 * i) = M;
        *reinterpret_cast<float4*>(v + 4 * i) = V;
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nth) adam_elem(w[i], g[i] * coef, m[i], v[i], c);
}


// ---------------------------------------------------------------------------------------------
// f1: touched-rows variants.  `rows` is the sorted-unique list of table rows a batch touched
// (rbx_unique_ids_i32 over the batch's global row ids), its length lives in DEVICE memory
// (n_rows_dev[0]) so nothing synchronises with the host.  Work is O(touched rows * D) instead of
// O(table): 1.7 M rows x 64 B per step instead of 7 passes over the whole table.
//   kind 0  SGD       w -= lr * g'                                   (exact: untouched rows have g = 0)
//   kind 1  Adagrad   s += g'^2 ; w -= lr * g' / (sqrt(s) + eps)     (exact, torch.optim.Adagrad, lr_decay = 0)
//   kind 2  Adam, dense formula of rbx_adam_dense applied to the touched rows only ("lazy" Adam:
//           NOT what the reference's dense torch.optim.Adam does to untouched rows, whose moments keep decaying)
//   kind 3  torch.optim.SparseAdam formula: denom = sqrt(v) + eps ; w -= lr * sqrt(bc2) / bc1 * m / denom
// g' = g * clip[0].  zero_grad != 0 writes zeros back over the consumed gradient rows, which keeps the
// dense gradient table all-zero between steps without the per-step O(table) memset.
// ---------------------------------------------------------------------------------------------
struct RowsOptConst {
    AdamConst a;
    float lr, sparse_step;   // kind 3: lr * sqrt(bc2) / bc1
    int kind, zero_grad;
};

__device__ __forceinline__ void rows_elem(float& w, float g, float& m, float& v, const RowsOptConst& c) {
    if (c.kind == 0) {
        w = w - c.lr * g;
    } else if (c.kind == 1) {
        v = fmaf(g, g, v);                                   // state_sum.addcmul_(grad, grad, value=1)
        w = w - c.lr * (g / (sqrtf(v) + c.a.eps));           // param.addcdiv_(grad, std, value=-clr)
    } else if (c.kind == 2) {
        adam_elem(w, g, m, v, c.a);
    } else {
        m = m + c.a.one_minus_b1 * (g - m);                  // exp_avg.add_((g - old) * (1 - beta1))
        v = v + c.a.one_minus_b2 * (g * g - v);              // exp_avg_sq.add_((g^2 - old) * (1 - beta2))
        w = w - c.sparse_step * (m / (sqrtf(v) + c.a.eps));  // numer / (sqrt + eps) * step_size
    }
}

template <int VEC>   // VEC = 4: D % 4 == 0 and 16-byte aligned; VEC = 1: anything
__global__ void __launch_bounds__(kThreads) k_optim_rows(float* __restrict__ w, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, const int32_t* __restrict__ rows,
                                                        const int64_t* __restrict__ n_rows_dev, int64_t max_rows, int D,
                                                        const float* __restrict__ clip, RowsOptConst c) {
    int64_t n_rows = n_rows_dev ? *n_rows_dev : max_rows;
    if (n_rows > max_rows) n_rows = max_rows;
    const float coef = clip ? __ldg(clip) : 1.f;
    const int per_row = D / VEC;
    const int64_t total = n_rows * per_row;
    const bool has_m = c.kind >= 2, has_v = c.kind >= 1;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        const int64_t i = t / per_row;
        const int col = (int)(t - i * per_row) * VEC;
        const size_t o = (size_t)__ldg(rows + i) * D + col;
        if (VEC == 4) {
            float4 W = *reinterpret_cast<float4*>(w + o);
            float4 G = *reinterpret_cast<const float4*>(g + o);
            float4 M = has_m ? *reinterpret_cast<float4*>(m + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 V = has_v ? *reinterpret_cast<float4*>(v + o) : make_float4(0.f, 0.f, 0.f, 0.f);
            rows_elem(W.x, G.x * coef, M.x, V.x, c);
            rows_elem(W.y, G.y * coef, M.y, V.y, c);
            rows_elem(W.z, G.z * coef, M.z, V.z, c);
            rows_elem(W.w, G.w * coef, M.w, V.w, c);
            *reinterpret_cast<float4*>(w + o) = W;
            if (has_m) *reinterpret_cast<float4*>(m + o) = M;
            if (has_v) *reinterpret_cast<float4*>(v + o) = V;
            if (c.zero_grad) *reinterpret_cast<float4*>(g + o) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float W = w[o], M = has_m ? m[o] : 0.f, V = has_v ? v[o] : 0.f;
            rows_elem(W, g[o] * coef, M, V, c);
            w[o] = W;
            if (has_m) m[o] = M;
            if (has_v) v[o] = V;
            if (c.zero_grad) g[o] = 0.f;
        }
    }
}

// out[0] += sum over the touched rows of |g[row,:]|^2  (== the dense sum: untouched rows are zero)
template <int VEC>
__global__ void __launch_bounds__(kThreads) k_sqnorm_rows(const float* __restrict__ g, const int32_t* __restrict__ rows,
                                                         const int64_t* __restrict__ n_rows_dev, int64_t max_rows, int D,
                                                         double* __restrict__ out) {
    __shared__ float s_part[kThreads / 32];
    int64_t n_rows = n_rows_dev ? *n_rows_dev : max_rows;
    if (n_rows > max_rows) n_rows = max_rows;
    const int per_row = D / VEC;
    const int64_t total = n_rows * per_row;
    float acc = 0.f;
    for (int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x; t < total; t += (int64_t)gridDim.x * kThreads) {
        const int64_t i = t / per_row;
        const int col = (int)(t - i * per_row) * VEC;
        const size_t o = (size_t)__ldg(rows + i) * D + col;
        if (VEC == 4) {
            const float4 x = *reinterpret_cast<const float4*>(g + o);
            acc = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, acc))));
        } else {
            acc = fmaf(g[o], g[o], acc);
        }
    }
    acc = group_sum<32>(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += (double)s_part[w];
        if (t != 0.0) atomicAdd(out, t);
    }
}

int rows_grid(int64_t max_rows, int per_row) {
    int64_t ctas = (max_rows * per_row + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    return ctas < 1 ? 1 : (int)ctas;
}

}  // namespace

extern "C" {

int rbx_sqnorm(const float* g, int64_t n, double* out, rbx_stream_t stream) {
    const char* who = "rbx_sqnorm";
    RBX_RANGE(who);
    RBX_REQUIRE(n >= 0 && out, "%s: bad argument", who);
    if (n == 0) return RBX_OK;
    RBX_REQUIRE(g != nullptr, "%s: null pointer", who);
    int64_t ctas = (n / 4 + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    k_sqnorm<<<(int)ctas, kThreads, 0, rbx_cast_stream(stream)>>>(g, n, out);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_clip_coef(const double* sqnorm, float max_norm, float* coef, float* norm_out, rbx_stream_t stream) {
    const char* who = "rbx_clip_coef";
    RBX_RANGE(who);
    RBX_REQUIRE(sqnorm && coef, "%s: null pointer", who);
    k_clip_coef<<<1, 1, 0, rbx_cast_stream(stream)>>>(sqnorm, max_norm, coef, norm_out);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_adam_dense(float* w, const float* g, float* m, float* v, int64_t n, const float* clip, float lr, float beta1,
                   float beta2, float eps, int step, rbx_stream_t stream) {
    const char* who = "rbx_adam_dense";
    RBX_RANGE(who);
    RBX_REQUIRE(n >= 0 && step >= 1, "%s: bad argument (step counts from 1)", who);
    if (n == 0) return RBX_OK;
    RBX_REQUIRE(w && g && m && v, "%s: null pointer", who);
    // scalar prefactors in double, as torch computes them in Python floats
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    AdamConst c;
    c.one_minus_b1 = (float)(1.0 - (double)beta1);
    c.b2 = beta2;
    c.one_minus_b2 = (float)(1.0 - (double)beta2);
    c.neg_step_size = (float)(-((double)lr / bc1));
    c.bc2_sqrt = (float)sqrt(bc2);
    c.eps = eps;
    const int vec_ok = ((uintptr_t)w % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) && ((uintptr_t)v % 16 == 0);
    int64_t ctas = (n / 4 + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)rbx_sm_count() * 8;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    k_adam_dense<<<(int)ctas, kThreads, 0, rbx_cast_stream(stream)>>>(w, g, m, v, n, clip, c, vec_ok);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_sqnorm_rows(const float* g, const int32_t* rows, const int64_t* n_rows_dev, int64_t max_rows, int D, double* out,
                    rbx_stream_t stream) {
    const char* who = "rbx_sqnorm_rows";
    RBX_RANGE(who);
    RBX_REQUIRE(max_rows >= 0 && D >= 1 && out, "%s: bad argument", who);
    if (max_rows == 0) return RBX_OK;
    RBX_REQUIRE(g && rows, "%s: null pointer", who);
    const bool vec = D % 4 == 0 && (uintptr_t)g % 16 == 0;
    if (vec) k_sqnorm_rows<4><<<rows_grid(max_rows, D / 4), kThreads, 0, rbx_cast_stream(stream)>>>(g, rows, n_rows_dev, max_rows, D, out);
    else k_sqnorm_rows<1><<<rows_grid(max_rows, D), kThreads, 0, rbx_cast_stream(stream)>>>(g, rows, n_rows_dev, max_rows, D, out);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_optim_rows(float* w, float* g, float* m, float* v, const int32_t* rows, const int64_t* n_rows_dev, int64_t max_rows,
                   int D, const float* clip, int kind, float lr, float beta1, float beta2, float eps, int step, int zero_grad,
                   rbx_stream_t stream) {
    const char* who = "rbx_optim_rows";
    RBX_RANGE(who);
    RBX_REQUIRE(max_rows >= 0 && D >= 1 && step >= 1, "%s: bad argument (step counts from 1)", who);
    RBX_REQUIRE(kind >= 0 && kind <= 3, "%s: kind=%d (0 sgd, 1 adagrad, 2 adam rows, 3 sparse adam)", who, kind);
    if (max_rows == 0) return RBX_OK;
    RBX_REQUIRE(w && g && rows, "%s: null pointer", who);
    RBX_REQUIRE(kind < 1 || v, "%s: kind %d needs the second-moment / sum state v", who, kind);
    RBX_REQUIRE(kind < 2 || m, "%s: kind %d needs the first-moment state m", who, kind);
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    RowsOptConst c;
    c.a.one_minus_b1 = (float)(1.0 - (double)beta1);
    c.a.b2 = beta2;
    c.a.one_minus_b2 = (float)(1.0 - (double)beta2);
    c.a.neg_step_size = (float)(-((double)lr / bc1));
    c.a.bc2_sqrt = (float)sqrt(bc2);
    c.a.eps = eps;
    c.lr = lr;
    c.sparse_step = (float)((double)lr * sqrt(bc2) / bc1);
    c.kind = kind;
    c.zero_grad = zero_grad;
    const bool vec = D % 4 == 0 && (uintptr_t)w % 16 == 0 && (uintptr_t)g % 16 == 0 && (!m || (uintptr_t)m % 16 == 0) &&
                     (!v || (uintptr_t)v % 16 == 0);
    cudaStream_t st = rbx_cast_stream(stream);
    if (vec) k_optim_rows<4><<<rows_grid(max_rows, D / 4), kThreads, 0, st>>>(w, g, m, v, rows, n_rows_dev, max_rows, D, clip, c);
    else k_optim_rows<1><<<rows_grid(max_rows, D), kThreads, 0, st>>>(w, g, m, v, rows, n_rows_dev, max_rows, D, clip, c);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

}  // extern "C"
