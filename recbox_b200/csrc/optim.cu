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

}  // namespace

extern "C" {

int rbx_sqnorm(const float* g, int64_t n, double* out, rbx_stream_t stream) {
    const char* who = "rbx_sqnorm";
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
    RBX_REQUIRE(sqnorm && coef, "%s: null pointer", who);
    k_clip_coef<<<1, 1, 0, rbx_cast_stream(stream)>>>(sqnorm, max_norm, coef, norm_out);
    RBX_LAUNCH_CHECK(who);
    return RBX_OK;
}

int rbx_adam_dense(float* w, const float* g, float* m, float* v, int64_t n, const float* clip, float lr, float beta1,
                   float beta2, float eps, int step, rbx_stream_t stream) {
    const char* who = "rbx_adam_dense";
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

}  // extern "C"
