// Fused AdamW step over a flat fp32 arena (hash tables + MLP weights), one streaming pass:
// 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter -- HBM bound.
// Replaces torch.optim.AdamW as configured by the reference (configs/neuralangelo-colmap_sparse.yaml:134-139,
// systems/utils.py:314-325): decoupled weight decay (default 0.01), betas (0.9, 0.99), eps 1e-15.
#include <math.h>
#include <algorithm>

#include "ia_common.cuh"

namespace {
__global__ void __launch_bounds__(256)
adamw_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m, float4 *__restrict__ v, int64_t n4,
             float *p_tail, const float *g_tail, float *m_tail, float *v_tail, int tail, float lr, float b1, float b2,
             float eps, float wd, float bc1, float bc2_sqrt, float gscale)
{
    auto upd = [&](float &pp, float gg, float &mm, float &vv) {
        gg *= gscale;
        pp *= 1.f - lr * wd;
        mm = b1 * mm + (1.f - b1) * gg;
        vv = b2 * vv + (1.f - b2) * gg * gg;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        pp -= (lr / bc1) * (mm / denom);
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        p[i] = pp; m[i] = mm; v[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) {
        const int t = threadIdx.x;
        upd(p_tail[t], g_tail[t], m_tail[t], v_tail[t]);
    }
}
}  // namespace

extern "C" int32_t ia_adamw_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                                 void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (param && grad && exp_avg && exp_avg_sq)), "adamw: NULL pointer");
    IA_REQUIRE(step >= 1, "adamw: step must be >= 1");
    IA_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0,
               "adamw: arenas must be 16-byte aligned");
    if (n == 0) return IA_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    const int64_t n4 = n / 4;
    const int tail = (int)(n - 4 * n4);
    const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ia_ceil_div(n4, 256), (int64_t)ia_sm_count() * 16));
    adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4 *>(param), reinterpret_cast<const float4 *>(grad), reinterpret_cast<float4 *>(exp_avg),
        reinterpret_cast<float4 *>(exp_avg_sq), n4, param + 4 * n4, grad + 4 * n4, exp_avg + 4 * n4, exp_avg_sq + 4 * n4, tail,
        lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale);
    IA_LAUNCH_OK("adamw_kernel");
    return IA_OK;
}
