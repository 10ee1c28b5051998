// Loss terms of the NeuS training step in two launches (forward sums + weighted total, backward of every term).
// Replaces the tensor expressions of reference systems/neus.py:132-160 (rgb MSE / L1 over the valid rays, eikonal,
// opaque / mask binary cross entropy on the clamped opacity, sparsity, curvature) and systems/criterions.py:155-159:
// as torch operators they are ~60 element-wise / reduction launches forward and as many in backward, per step, on tensors
// of 8 k rays / 1.5 M samples -- launch-bound glue (1.0 ms of GPU idle + kernel time per 24 ms step, and the larger part
// of a step at the 256-ray batches training starts from).
//
// Forward: one grid-stride pass; thread i takes ray i and sample i, per-CTA tree reduction, one double atomicAdd per CTA
// and term; the last CTA to finish (ticket counter) turns the sums into the terms and the lambda-weighted total.
// Backward: element-wise, reads the upstream gradient of the total from device memory.
#include <math.h>

#include "ia_common.cuh"

namespace {

constexpr int LS_THREADS = 256;
constexpr int LS_TERMS = 8;      // 0 rgb_mse 1 rgb_l1 2 eikonal 3 mask 4 opaque 5 sparsity 6 curvature 7 n_valid

struct LossPtrs {
    const float *comp_rgb, *rgb_gt;
    const uint8_t *valid;
    const float *opacity, *fg_mask;
    const float *sdf_grad, *sdf, *laplace;
};

__device__ __forceinline__ float clamp_opacity(float o) { return fminf(fmaxf(o, 1.0e-3f), 1.0f - 1.0e-3f); }

__global__ void __launch_bounds__(LS_THREADS)
losses_fwd_kernel(const ia_loss_args A, const LossPtrs P, double *__restrict__ sums, unsigned int *__restrict__ ticket,
                  float *__restrict__ terms, float *__restrict__ loss)
{
    float acc[LS_TERMS];
#pragma unroll
    for (int k = 0; k < LS_TERMS; ++k) acc[k] = 0.f;
    const int64_t stride = (int64_t)gridDim.x * LS_THREADS;
    const int64_t n = A.n_rays > A.n_samples ? A.n_rays : A.n_samples;
    for (int64_t i = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x; i < n; i += stride) {
        if (i < A.n_rays) {
            if (P.valid[i]) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float d = P.comp_rgb[3 * i + c] - P.rgb_gt[3 * i + c];
                    acc[0] = fmaf(d, d, acc[0]);
                    acc[1] += fabsf(d);
                }
                acc[7] += 1.f;
            }
            const float o = clamp_opacity(P.opacity[i]);
            const float lo = logf(o), l1o = logf(1.f - o);
            acc[4] -= o * lo + (1.f - o) * l1o;
            if (P.fg_mask) {
                const float t = P.fg_mask[i];
                acc[3] -= t * lo + (1.f - t) * l1o;
            }
        }
        if (i < A.n_samples) {
            const float gx = P.sdf_grad[3 * i], gy = P.sdf_grad[3 * i + 1], gz = P.sdf_grad[3 * i + 2];
            const float e = sqrtf(gx * gx + gy * gy + gz * gz) - 1.f;
            acc[2] = fmaf(e, e, acc[2]);
            acc[5] += expf(-A.sparsity_scale * fabsf(P.sdf[i]));
            if (P.laplace) acc[6] += fabsf(P.laplace[i]);
        }
    }
    __shared__ float red[LS_TERMS][LS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < LS_TERMS; ++k) {
        const float s = ia_warp_sum(acc[k]);
        if (lane == 0) red[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < LS_TERMS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < LS_THREADS / 32; ++w) s += (double)red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, s);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence();
            volatile double *vs = sums;
            const double n3 = 3.0 * vs[7], R = (double)A.n_rays, S = (double)A.n_samples;
            // means over empty sets are 0/0 = nan, as the reference's .mean() / F.mse_loss on empty tensors
            const float t_mse = (float)(vs[0] / n3), t_l1 = (float)(vs[1] / n3), t_eik = (float)(vs[2] / S);
            const float t_mask = P.fg_mask ? (float)(vs[3] / R) : 0.f, t_opq = (float)(vs[4] / R), t_sp = (float)(vs[5] / S);
            const bool curv = P.laplace != nullptr && A.lambda_curvature > 0.f;
            const float t_curv = curv ? (float)(vs[6] / S) : 0.f;
            terms[0] = t_mse; terms[1] = t_l1; terms[2] = t_eik; terms[3] = t_mask;
            terms[4] = t_opq; terms[5] = t_sp; terms[6] = t_curv; terms[7] = (float)vs[7];
            // the order of the reference's `loss +=` lines (systems/neus.py:134-160)
            float l = t_mse * A.lambda_rgb_mse;
            l += t_l1 * A.lambda_rgb_l1;
            l += t_eik * A.lambda_eikonal;
            if (P.fg_mask) l += t_mask * A.lambda_mask;
            l += t_opq * A.lambda_opaque;
            l += t_sp * A.lambda_sparsity;
            if (curv) l += t_curv * A.lambda_curvature;
            *loss = l;
        }
    }
}

__global__ void __launch_bounds__(LS_THREADS)
losses_bwd_kernel(const ia_loss_args A, const LossPtrs P, const float *__restrict__ terms, const float *__restrict__ dloss,
                  float *__restrict__ d_comp_rgb, float *__restrict__ d_opacity, float *__restrict__ d_sdf_grad,
                  float *__restrict__ d_sdf, float *__restrict__ d_laplace)
{
    const float dl = __ldg(dloss);
    const float n3 = 3.f * __ldg(terms + 7);
    const float inv_r = 1.f / (float)A.n_rays, inv_s = 1.f / (float)A.n_samples;
    const float c_mse = dl * A.lambda_rgb_mse * 2.f / n3, c_l1 = dl * A.lambda_rgb_l1 / n3;
    const int64_t stride = (int64_t)gridDim.x * LS_THREADS;
    const int64_t n = A.n_rays > A.n_samples ? A.n_rays : A.n_samples;
    for (int64_t i = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x; i < n; i += stride) {
        if (i < A.n_rays) {
            if (d_comp_rgb) {
                const bool v = P.valid[i] != 0;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float d = P.comp_rgb[3 * i + c] - P.rgb_gt[3 * i + c];
                    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                    d_comp_rgb[3 * i + c] = v ? c_mse * d + c_l1 * sg : 0.f;
                }
            }
            if (d_opacity) {
                const float raw = P.opacity[i];
                float g = 0.f;
                if (raw >= 1.0e-3f && raw <= 1.0f - 1.0e-3f) {      // torch.clamp passes the gradient inside [min, max]
                    // BCE(o, o): both arguments are the opacity (systems/neus.py:149), so d/do = log(1 - o) - log(o)
                    g = dl * A.lambda_opaque * inv_r * (logf(1.f - raw) - logf(raw));
                    if (P.fg_mask) {
                        const float t = P.fg_mask[i];
                        g -= dl * A.lambda_mask * inv_r * (t / raw - (1.f - t) / (1.f - raw));
                    }
                }
                d_opacity[i] = g;
            }
        }
        if (i < A.n_samples) {
            if (d_sdf_grad) {
                const float gx = P.sdf_grad[3 * i], gy = P.sdf_grad[3 * i + 1], gz = P.sdf_grad[3 * i + 2];
                const float nrm = sqrtf(gx * gx + gy * gy + gz * gz);
                const float k = nrm > 0.f ? dl * A.lambda_eikonal * inv_s * 2.f * (nrm - 1.f) / nrm : 0.f;
                d_sdf_grad[3 * i] = k * gx;
                d_sdf_grad[3 * i + 1] = k * gy;
                d_sdf_grad[3 * i + 2] = k * gz;
            }
            if (d_sdf) {
                const float s = P.sdf[i];
                const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
                d_sdf[i] = -dl * A.lambda_sparsity * inv_s * A.sparsity_scale * sg * expf(-A.sparsity_scale * fabsf(s));
            }
            if (d_laplace) {
                const float v = P.laplace[i];
                const float sg = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
                d_laplace[i] = A.lambda_curvature > 0.f ? dl * A.lambda_curvature * inv_s * sg : 0.f;
            }
        }
    }
}

int check_args(const ia_loss_args *a, const LossPtrs &P, const char *who)
{
    IA_REQUIRE(a != nullptr, "%s: args is NULL", who);
    IA_REQUIRE(a->n_rays >= 0 && a->n_samples >= 0, "%s: negative sizes", who);
    IA_REQUIRE(a->n_rays == 0 || (P.comp_rgb && P.rgb_gt && P.valid && P.opacity), "%s: NULL ray tensor", who);
    IA_REQUIRE(a->n_samples == 0 || (P.sdf_grad && P.sdf), "%s: NULL sample tensor", who);
    return IA_OK;
}

unsigned grid_for(int64_t n)
{
    const int64_t want = ia_ceil_div(n > 0 ? n : 1, LS_THREADS);
    const int64_t cap = (int64_t)ia_sm_count() * 8;
    return (unsigned)(want < cap ? want : cap);
}

}  // namespace

extern "C" int64_t ia_neus_losses_workspace_bytes(void) { return (int64_t)(LS_TERMS * sizeof(double) + 16); }

extern "C" int32_t ia_neus_losses_fwd(const ia_loss_args *args, const float *comp_rgb, const float *rgb_gt, const uint8_t *valid,
                                      const float *opacity, const float *fg_mask, const float *sdf_grad, const float *sdf,
                                      const float *laplace, void *workspace, float *terms, float *loss, void *stream)
{
    const LossPtrs P = {comp_rgb, rgb_gt, valid, opacity, fg_mask, sdf_grad, sdf, laplace};
    int rc = check_args(args, P, "neus_losses_fwd");
    if (rc) return rc;
    IA_REQUIRE(workspace && terms && loss, "neus_losses_fwd: NULL output");
    cudaStream_t s = (cudaStream_t)stream;
    IA_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)ia_neus_losses_workspace_bytes(), s));
    double *sums = reinterpret_cast<double *>(workspace);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(sums + LS_TERMS);
    const int64_t n = args->n_rays > args->n_samples ? args->n_rays : args->n_samples;
    losses_fwd_kernel<<<grid_for(n), LS_THREADS, 0, s>>>(*args, P, sums, ticket, terms, loss);
    IA_LAUNCH_OK("losses_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_neus_losses_bwd(const ia_loss_args *args, const float *comp_rgb, const float *rgb_gt, const uint8_t *valid,
                                      const float *opacity, const float *fg_mask, const float *sdf_grad, const float *sdf,
                                      const float *laplace, const float *terms, const float *dloss, float *d_comp_rgb,
                                      float *d_opacity, float *d_sdf_grad, float *d_sdf, float *d_laplace, void *stream)
{
    const LossPtrs P = {comp_rgb, rgb_gt, valid, opacity, fg_mask, sdf_grad, sdf, laplace};
    int rc = check_args(args, P, "neus_losses_bwd");
    if (rc) return rc;
    IA_REQUIRE(terms && dloss, "neus_losses_bwd: NULL terms / dloss");
    IA_REQUIRE(d_laplace == nullptr || laplace != nullptr, "neus_losses_bwd: d_laplace without laplace");
    const int64_t n = args->n_rays > args->n_samples ? args->n_rays : args->n_samples;
    if (n == 0) return IA_OK;
    losses_bwd_kernel<<<grid_for(n), LS_THREADS, 0, (cudaStream_t)stream>>>(*args, P, terms, dloss, d_comp_rgb, d_opacity, d_sdf_grad,
                                                                            d_sdf, d_laplace);
    IA_LAUNCH_OK("losses_bwd_kernel");
    return IA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The sparse-point terms (reference systems/neus.py:173-186): with the SDF and its gradient evaluated at the SfM points,
//   sdf_l1     = (F.l1_loss(sdf, 0) * weights).mean()  = mean|sdf| * mean(weights)          (a scalar times the weights: C-11)
//   normal_cos = (1 - sum(normalize(grad) * normalize(normal_gt), -1)).mean()
//   total      = lambda_sdf_l1 * sdf_l1 + lambda_normal * normal_cos
// one launch each way instead of ~14 + ~14.  out = (sdf_l1, normal_cos, total); workspace as ia_neus_losses_fwd.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void unit3(const float *__restrict__ p, float eps, float (&u)[3], float &nrm)
{
    const float a = p[0], b = p[1], c = p[2];
    nrm = sqrtf(a * a + b * b + c * c);
    const float den = fmaxf(nrm, eps);
    u[0] = a / den; u[1] = b / den; u[2] = c / den;
}

__global__ void __launch_bounds__(LS_THREADS)
point_losses_fwd_kernel(const float *__restrict__ sdf, const float *__restrict__ grad, const float *__restrict__ normal_gt,
                        const float *__restrict__ weights, int64_t n, float lambda_sdf, float lambda_normal,
                        double *__restrict__ sums, unsigned int *__restrict__ ticket, float *__restrict__ out)
{
    float acc[3] = {0.f, 0.f, 0.f};            // sum |sdf|, sum weights, sum (1 - cos)
    const int64_t stride = (int64_t)gridDim.x * LS_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x; i < n; i += stride) {
        acc[0] += fabsf(sdf[i]);
        acc[1] += weights[i];
        float a[3], b[3], na, nb;
        unit3(grad + 3 * i, 1e-12f, a, na);
        unit3(normal_gt + 3 * i, 1e-12f, b, nb);
        acc[2] += 1.0f - (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]);
    }
    __shared__ float red[3][LS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float s = ia_warp_sum(acc[k]);
        if (lane == 0) red[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < LS_THREADS / 32; ++w) s += (double)red[threadIdx.x][w];
        atomicAdd(sums + threadIdx.x, s);
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence();
            volatile double *vs = sums;
            const double P = (double)n;
            const float l1 = (float)((vs[0] / P) * (vs[1] / P)), nc = (float)(vs[2] / P);
            out[0] = l1;
            out[1] = nc;
            out[2] = l1 * lambda_sdf + nc * lambda_normal;
            out[3] = (float)(vs[1] / P);            // mean(weights), for the backward
        }
    }
}

__global__ void __launch_bounds__(LS_THREADS)
point_losses_bwd_kernel(const float *__restrict__ sdf, const float *__restrict__ grad, const float *__restrict__ normal_gt, int64_t n,
                        float lambda_sdf, float lambda_normal, const float *__restrict__ out, const float *__restrict__ dloss,
                        float *__restrict__ d_sdf, float *__restrict__ d_grad)
{
    const int64_t i = (int64_t)blockIdx.x * LS_THREADS + threadIdx.x;
    if (i >= n) return;
    const float dl = __ldg(dloss), inv_p = 1.f / (float)n;
    if (d_sdf) {
        const float s = sdf[i];
        const float sg = s > 0.f ? 1.f : (s < 0.f ? -1.f : 0.f);
        d_sdf[i] = dl * lambda_sdf * __ldg(out + 3) * inv_p * sg;
    }
    if (d_grad) {
        float a[3], b[3], na, nb;
        unit3(grad + 3 * i, 1e-12f, a, na);
        unit3(normal_gt + 3 * i, 1e-12f, b, nb);
        const float k = -dl * lambda_normal * inv_p;        // d total / d cos
        if (na > 1e-12f) {
            const float dot = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
#pragma unroll
            for (int c = 0; c < 3; ++c) d_grad[3 * i + c] = k * (b[c] - a[c] * dot) / na;
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) d_grad[3 * i + c] = k * b[c] / 1e-12f;
        }
    }
}

}  // namespace

extern "C" int32_t ia_point_losses_fwd(const float *sdf, const float *grad, const float *normal_gt, const float *weights, int64_t n,
                                       float lambda_sdf_l1, float lambda_normal, void *workspace, float *out4, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (sdf && grad && normal_gt && weights)) && workspace && out4, "point_losses_fwd: NULL pointer");
    cudaStream_t s = (cudaStream_t)stream;
    IA_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)ia_neus_losses_workspace_bytes(), s));
    double *sums = reinterpret_cast<double *>(workspace);
    unsigned int *ticket = reinterpret_cast<unsigned int *>(sums + LS_TERMS);
    point_losses_fwd_kernel<<<grid_for(n), LS_THREADS, 0, s>>>(sdf, grad, normal_gt, weights, n, lambda_sdf_l1, lambda_normal, sums,
                                                                ticket, out4);
    IA_LAUNCH_OK("point_losses_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_point_losses_bwd(const float *sdf, const float *grad, const float *normal_gt, int64_t n, float lambda_sdf_l1,
                                       float lambda_normal, const float *out4, const float *dloss, float *d_sdf, float *d_grad,
                                       void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (sdf && grad && normal_gt)) && out4 && dloss, "point_losses_bwd: NULL pointer");
    if (n == 0) return IA_OK;
    point_losses_bwd_kernel<<<(unsigned)ia_ceil_div(n, LS_THREADS), LS_THREADS, 0, (cudaStream_t)stream>>>(
        sdf, grad, normal_gt, n, lambda_sdf_l1, lambda_normal, out4, dloss, d_sdf, d_grad);
    IA_LAUNCH_OK("point_losses_bwd_kernel");
    return IA_OK;
}
