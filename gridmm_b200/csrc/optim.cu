// The optimizer half of the pretraining step (BASELINE config 5): after the gradient all-reduce over NVLink
// (torch.distributed / NCCL, gridmm_b200/train.py) every rank applies the same update to its replica.  Reference:
//   clip_grad_norm_ over all parameters            pretrain_src/train_r2r.py:281-285
//   AdamW (decoupled weight decay, bias-corrected) pretrain_src/optim/adamw.py:57-104
//   two parameter groups (weight decay 0.01 / 0)   pretrain_src/optim/misc.py:12-22
// Parameters, gradients and both moments live in FLAT fp32 buffers (one contiguous range per group), so the whole update is two
// launches over ~198 M elements -- HBM-bound: 16 B read + 12 B written per element.
#include "common.cuh"
#include "host_util.h"

namespace gmm {

// out[0] += sum g[i]^2 (fp32 partial sums per thread, warp shuffle, one atomicAdd per CTA: order-independent up to fp32 rounding)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* out) {
    float s = 0.f;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
        const float4 v = g4[i];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    for (long long i = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * 256) s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = red[threadIdx.x];
        v += __shfl_xor_sync(0xffu, v, 4); v += __shfl_xor_sync(0xffu, v, 2); v += __shfl_xor_sync(0xffu, v, 1);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

struct AdamWParams {
    float* p; const float* g; float* m; float* v;
    long long n;
    float lr, beta1, beta2, eps, weight_decay, step_size;     // step_size = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    float grad_scale;            // multiplies every gradient first (1 / world size, 1 / loss scale)
    const float* sumsq;          // device scalar: sum of squares of the SCALED... see below; or null (no clipping)
    float max_norm;              // clip_grad_norm_: g *= min(1, max_norm / (||g|| + 1e-6)), ||g|| taken after grad_scale
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWParams& a, float gs) {
    g *= gs;
    m = a.beta1 * m + (1.0f - a.beta1) * g;
    v = a.beta2 * v + (1.0f - a.beta2) * g * g;
    p -= a.step_size * (m / (sqrtf(v) + a.eps));
    if (a.weight_decay > 0.0f) p -= a.lr * a.weight_decay * p;       // decoupled, on the updated value (adamw.py:101-102)
}

__global__ void __launch_bounds__(256) adamw_kernel(AdamWParams a) {
    float gs = a.grad_scale;
    if (a.sumsq) {
        const float norm = sqrtf(*a.sumsq) * a.grad_scale;            // the sum of squares was taken over the unscaled gradients
        const float coef = a.max_norm / (norm + 1e-6f);
        if (coef < 1.0f) gs *= coef;
    }
    const long long n4 = a.n >> 2;
    float4* p4 = reinterpret_cast<float4*>(a.p); const float4* g4 = reinterpret_cast<const float4*>(a.g);
    float4* m4 = reinterpret_cast<float4*>(a.m); float4* v4 = reinterpret_cast<float4*>(a.v);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = g4[i];
        adamw_one(p.x, g.x, m.x, v.x, a, gs); adamw_one(p.y, g.y, m.y, v.y, a, gs);
        adamw_one(p.z, g.z, m.z, v.z, a, gs); adamw_one(p.w, g.w, m.w, v.w, a, gs);
        p4[i] = p; m4[i] = m; v4[i] = v;
    }
    for (long long i = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; i < a.n; i += static_cast<long long>(gridDim.x) * 256)
        adamw_one(a.p[i], a.g[i], a.m[i], a.v[i], a, gs);
}

}  // namespace gmm

// out[0] += sum_i g[i]^2 (the caller zeroes `out` before the first range); g must be 16-byte aligned
extern "C" int gridmm_grad_sumsq(const float* g, long long n, float* out, cudaStream_t stream) {
    using namespace gmm;
    if (n <= 0) return 0;
    if (!g || !out) return GRIDMM_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(g) & 15) return GRIDMM_ERR_SHAPE;
    const int sms = gridmm_sm_count();
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = static_cast<long long>(sms > 0 ? sms : 148) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(g, n, out);
    gridmm_count_launch(1);
    return static_cast<int>(cudaGetLastError());
}

// One AdamW step (pretrain_src/optim/adamw.py:57-104) over a flat fp32 range: p, m, v updated in place from g * grad_scale, clipped by
// the global norm when sumsq != NULL (device scalar = sum of squares of ALL unscaled gradients, gridmm_grad_sumsq; max_norm as in
// torch.nn.utils.clip_grad_norm_).  step >= 1 is the optimizer step count (bias correction).
extern "C" int gridmm_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                                 float eps, float weight_decay, int step, float grad_scale, const float* sumsq, float max_norm,
                                 cudaStream_t stream) {
    using namespace gmm;
    if (n <= 0) return 0;
    if (!p || !g || !m || !v || step < 1) return GRIDMM_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
        return GRIDMM_ERR_SHAPE;
    AdamWParams a;
    a.p = p; a.g = g; a.m = m; a.v = v; a.n = n; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step), bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    a.step_size = static_cast<float>(static_cast<double>(lr) * sqrt(bc2) / bc1);
    a.grad_scale = grad_scale; a.sumsq = sumsq; a.max_norm = max_norm;
    const int sms = gridmm_sm_count();
    long long blocks = (n / 4 + 255) / 256;
    const long long cap = static_cast<long long>(sms > 0 ? sms : 148) * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adamw_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(a);
    gridmm_count_launch(1);
    return static_cast<int>(cudaGetLastError());
}
