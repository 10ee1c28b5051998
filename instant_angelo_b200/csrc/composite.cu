// Segmented per-ray alpha/transmittance scan + weighted reductions, forward and backward.
// Fuses reference NeuSModel.get_alpha (models/neus.py:117-139), nerfacc.render_weight_from_alpha /
// render_weight_from_density (models/neus.py:181, 234) and the four / three
// nerfacc.accumulate_along_rays calls (models/neus.py:182-184, 235-239) into ONE pass per direction.
//
// Mapping: one warp per ray.  Samples of a ray are contiguous (packed_info = (offset, count)), so the
// warp streams them 32 at a time with fully coalesced loads, runs the exclusive product scan of
// (1 - alpha) with shuffles and carries the running transmittance between chunks.  Backward walks the
// chunks in reverse with a shuffle suffix-sum of g_k * w_k.  HBM-streaming kernel: every per-sample
// array is read once and written once.
#include <math.h>

#include "ia_common.cuh"

namespace {

constexpr int CP_WARPS = 8;  // warps (rays) per CTA

struct CompArgs {
    int mode;
    int64_t n_rays;
    const int32_t *packed_info;
    const float *alpha_in;
    const float *sdf, *normal, *dirs, *dists, *inv_s;
    float cos_anneal;
    const float *sigma, *t_starts, *t_ends;
    const float *t_mid, *rgb, *nrm;
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

struct NeusTerms {
    float c, prev, next, P, N, raw;
};

// get_alpha of models/neus.py:117-139 for one sample
__device__ __forceinline__ NeusTerms neus_alpha(const CompArgs &A, int64_t s, float inv_s)
{
    NeusTerms t;
    const float nx = A.normal[3 * s], ny = A.normal[3 * s + 1], nz = A.normal[3 * s + 2];
    const float dx = A.dirs[3 * s], dy = A.dirs[3 * s + 1], dz = A.dirs[3 * s + 2];
    t.c = dx * nx + dy * ny + dz * nz;
    const float a = A.cos_anneal;
    const float iter_cos = -(fmaxf(-t.c * 0.5f + 0.5f, 0.f) * (1.0f - a) + fmaxf(-t.c, 0.f) * a);
    const float e = iter_cos * A.dists[s] * 0.5f;
    const float sd = A.sdf[s];
    t.next = sd + e;
    t.prev = sd - e;
    t.P = sigmoidf_(t.prev * inv_s);
    t.N = sigmoidf_(t.next * inv_s);
    t.raw = (t.P - t.N + 1e-5f) / (t.P + 1e-5f);
    return t;
}

__device__ __forceinline__ float sample_alpha(const CompArgs &A, int64_t s, float inv_s)
{
    if (A.mode == IA_ALPHA_NEUS) {
        const NeusTerms t = neus_alpha(A, s, inv_s);
        return fminf(fmaxf(t.raw, 0.f), 1.f);
    }
    if (A.mode == IA_ALPHA_DENSITY) return 1.0f - expf(-A.sigma[s] * (A.t_ends[s] - A.t_starts[s]));
    return A.alpha_in[s];
}

__global__ void __launch_bounds__(CP_WARPS * 32)
composite_fwd_kernel(const CompArgs A, float *__restrict__ alpha_out, float *__restrict__ trans_out,
                     float *__restrict__ weights, float *__restrict__ opacity, float *__restrict__ depth,
                     float *__restrict__ comp_rgb, float *__restrict__ comp_nrm)
{
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * CP_WARPS + (threadIdx.x >> 5);
    if (r >= A.n_rays) return;
    const int64_t off = A.packed_info[2 * r];
    const int cnt = A.packed_info[2 * r + 1];
    const float inv_s = A.mode == IA_ALPHA_NEUS ? __ldg(A.inv_s) : 0.f;
    float T = 1.0f;
    float acc_o = 0.f, acc_d = 0.f, acc_c[3] = {0.f, 0.f, 0.f}, acc_n[3] = {0.f, 0.f, 0.f};
    for (int b = 0; b < cnt; b += 32) {
        const int j = b + lane;
        const bool valid = j < cnt;
        const int64_t s = off + j;
        const float a = valid ? sample_alpha(A, s, inv_s) : 0.f;
        float incl = 1.0f - a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= t;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float Ti = T * excl;
        const float w = Ti * a;
        T = T * __shfl_sync(0xffffffffu, incl, 31);
        if (valid) {
            if (alpha_out) alpha_out[s] = a;
            if (trans_out) trans_out[s] = Ti;
            if (weights) weights[s] = w;
            acc_o += w;
            if (A.t_mid) acc_d += w * A.t_mid[s];
            if (A.rgb) {
                acc_c[0] += w * A.rgb[3 * s];
                acc_c[1] += w * A.rgb[3 * s + 1];
                acc_c[2] += w * A.rgb[3 * s + 2];
            }
            if (A.nrm) {
                acc_n[0] += w * A.nrm[3 * s];
                acc_n[1] += w * A.nrm[3 * s + 1];
                acc_n[2] += w * A.nrm[3 * s + 2];
            }
        }
    }
    acc_o = ia_warp_sum(acc_o);
    acc_d = ia_warp_sum(acc_d);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        acc_c[k] = ia_warp_sum(acc_c[k]);
        acc_n[k] = ia_warp_sum(acc_n[k]);
    }
    if (lane == 0) {
        if (opacity) opacity[r] = acc_o;
        if (depth) depth[r] = acc_d;
        if (comp_rgb) { comp_rgb[3 * r] = acc_c[0]; comp_rgb[3 * r + 1] = acc_c[1]; comp_rgb[3 * r + 2] = acc_c[2]; }
        if (comp_nrm) { comp_nrm[3 * r] = acc_n[0]; comp_nrm[3 * r + 1] = acc_n[1]; comp_nrm[3 * r + 2] = acc_n[2]; }
    }
}

__global__ void __launch_bounds__(CP_WARPS * 32)
composite_bwd_kernel(const CompArgs A, const float *__restrict__ alpha, const float *__restrict__ trans,
                     const float *__restrict__ g_weights, const float *__restrict__ g_opacity,
                     const float *__restrict__ g_depth, const float *__restrict__ g_rgb,
                     const float *__restrict__ g_nrm, float *__restrict__ d_alpha_in, float *__restrict__ d_sdf,
                     float *__restrict__ d_normal, float *__restrict__ d_inv_s, float *__restrict__ d_sigma,
                     float *__restrict__ d_rgb, float *__restrict__ d_nrm)
{
    __shared__ float s_dinv[CP_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * CP_WARPS + warp;
    float dinv_acc = 0.f;
    if (r < A.n_rays) {
        const int64_t off = A.packed_info[2 * r];
        const int cnt = A.packed_info[2 * r + 1];
        const float inv_s = A.mode == IA_ALPHA_NEUS ? __ldg(A.inv_s) : 0.f;
        const float go = g_opacity ? g_opacity[r] : 0.f;
        const float gd = (g_depth && A.t_mid) ? g_depth[r] : 0.f;
        float gc[3] = {0.f, 0.f, 0.f}, gn[3] = {0.f, 0.f, 0.f};
        if (g_rgb && A.rgb) { gc[0] = g_rgb[3 * r]; gc[1] = g_rgb[3 * r + 1]; gc[2] = g_rgb[3 * r + 2]; }
        if (g_nrm && A.nrm) { gn[0] = g_nrm[3 * r]; gn[1] = g_nrm[3 * r + 1]; gn[2] = g_nrm[3 * r + 2]; }
        float suffix = 0.f;  // sum of g_k * w_k over samples after the current chunk
        const int last = cnt > 0 ? ((cnt - 1) / 32) * 32 : -1;
        for (int b = last; b >= 0; b -= 32) {
            const int j = b + lane;
            const bool valid = j < cnt;
            const int64_t s = off + j;
            float a = 0.f, Ti = 0.f, g = 0.f;
            if (valid) {
                a = alpha[s];
                Ti = trans[s];
                g = go;
                if (g_weights) g += g_weights[s];
                if (A.t_mid) g += gd * A.t_mid[s];
                if (A.rgb) g += gc[0] * A.rgb[3 * s] + gc[1] * A.rgb[3 * s + 1] + gc[2] * A.rgb[3 * s + 2];
                if (A.nrm) g += gn[0] * A.nrm[3 * s] + gn[1] * A.nrm[3 * s + 1] + gn[2] * A.nrm[3 * s + 2];
            }
            const float w = Ti * a;
            const float gw = g * w;
            float rs = gw;  // reverse inclusive scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float t = __shfl_down_sync(0xffffffffu, rs, o);
                if (lane + o < 32) rs += t;
            }
            const float after = suffix + (rs - gw);
            suffix += __shfl_sync(0xffffffffu, rs, 0);
            if (!valid) continue;
            const float dalpha = g * Ti - after / fmaxf(1.0f - a, 1e-10f);
            if (d_rgb) { d_rgb[3 * s] = w * gc[0]; d_rgb[3 * s + 1] = w * gc[1]; d_rgb[3 * s + 2] = w * gc[2]; }
            if (d_nrm) { d_nrm[3 * s] = w * gn[0]; d_nrm[3 * s + 1] = w * gn[1]; d_nrm[3 * s + 2] = w * gn[2]; }
            if (A.mode == IA_ALPHA_GIVEN) {
                if (d_alpha_in) d_alpha_in[s] = dalpha;
            } else if (A.mode == IA_ALPHA_DENSITY) {
                // alpha = 1 - exp(-sigma*dt)  =>  d alpha / d sigma = dt * (1 - alpha)
                if (d_sigma) d_sigma[s] = dalpha * (A.t_ends[s] - A.t_starts[s]) * (1.0f - a);
            } else {
                const NeusTerms t = neus_alpha(A, s, inv_s);
                float dsdf = 0.f, dc = 0.f;
                if (t.raw >= 0.f && t.raw <= 1.f) {
                    const float den = t.P + 1e-5f;
                    const float dP = dalpha * (t.N / (den * den));
                    const float dN = -dalpha / den;
                    const float dpa = dP * t.P * (1.0f - t.P);  // grad wrt (prev*inv_s)
                    const float dna = dN * t.N * (1.0f - t.N);  // grad wrt (next*inv_s)
                    dinv_acc += dpa * t.prev + dna * t.next;
                    const float dprev = dpa * inv_s, dnext = dna * inv_s;
                    dsdf = dprev + dnext;
                    const float de = dnext - dprev;
                    const float dic = de * A.dists[s] * 0.5f;
                    const float an = A.cos_anneal;
                    // d iter_cos / d cos, relu'(0) = 0 as in torch
                    const float dic_dc = ((-t.c * 0.5f + 0.5f) > 0.f ? 0.5f * (1.0f - an) : 0.f) + ((-t.c) > 0.f ? an : 0.f);
                    dc = dic * dic_dc;
                }
                if (d_sdf) d_sdf[s] = dsdf;
                if (d_normal) {
                    d_normal[3 * s] = dc * A.dirs[3 * s];
                    d_normal[3 * s + 1] = dc * A.dirs[3 * s + 1];
                    d_normal[3 * s + 2] = dc * A.dirs[3 * s + 2];
                }
            }
        }
    }
    if (A.mode == IA_ALPHA_NEUS && d_inv_s != nullptr) {
        dinv_acc = ia_warp_sum(dinv_acc);
        if (lane == 0) s_dinv[warp] = dinv_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < CP_WARPS; ++w) t += s_dinv[w];
            if (t != 0.f) atomicAdd(d_inv_s, t);
        }
    }
}

int to_args(const ia_composite_args *a, CompArgs *A)
{
    IA_REQUIRE(a != nullptr, "composite: args is NULL");
    IA_REQUIRE(a->n_rays >= 0 && a->n_samples >= 0, "composite: negative sizes");
    IA_REQUIRE(a->n_rays == 0 || a->packed_info != nullptr, "composite: packed_info is NULL");
    if (a->n_samples > 0) {
        if (a->mode == IA_ALPHA_GIVEN) IA_REQUIRE(a->alpha_in, "composite: alpha_in is NULL");
        else if (a->mode == IA_ALPHA_NEUS)
            IA_REQUIRE(a->sdf && a->normal && a->dirs && a->dists && a->inv_s, "composite: NeuS inputs missing");
        else if (a->mode == IA_ALPHA_DENSITY)
            IA_REQUIRE(a->sigma && a->t_starts && a->t_ends, "composite: density inputs missing");
        else
            IA_REQUIRE(false, "composite: unknown mode %d", a->mode);
    }
    A->mode = a->mode;
    A->n_rays = a->n_rays;
    A->packed_info = a->packed_info;
    A->alpha_in = a->alpha_in;
    A->sdf = a->sdf; A->normal = a->normal; A->dirs = a->dirs; A->dists = a->dists; A->inv_s = a->inv_s;
    A->cos_anneal = a->cos_anneal_ratio;
    A->sigma = a->sigma; A->t_starts = a->t_starts; A->t_ends = a->t_ends;
    A->t_mid = a->t_mid; A->rgb = a->rgb; A->nrm = a->nrm;
    return IA_OK;
}

}  // namespace

extern "C" int32_t ia_composite_fwd(const ia_composite_args *args, float *alpha, float *trans, float *weights,
                                    float *opacity, float *depth, float *comp_rgb, float *comp_normal, void *stream)
{
    CompArgs A;
    int rc = to_args(args, &A);
    if (rc) return rc;
    if (A.n_rays == 0) return IA_OK;
    composite_fwd_kernel<<<(unsigned)ia_ceil_div(A.n_rays, CP_WARPS), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        A, alpha, trans, weights, opacity, depth, comp_rgb, comp_normal);
    IA_LAUNCH_OK("composite_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_composite_bwd(const ia_composite_args *args, const float *alpha, const float *trans,
                                    const float *g_weights, const float *g_opacity, const float *g_depth,
                                    const float *g_comp_rgb, const float *g_comp_normal, float *d_alpha_in,
                                    float *d_sdf, float *d_normal, float *d_inv_s, float *d_sigma, float *d_rgb,
                                    float *d_nrm, void *stream)
{
    CompArgs A;
    int rc = to_args(args, &A);
    if (rc) return rc;
    IA_REQUIRE(args->n_samples == 0 || (alpha && trans), "composite_bwd: saved alpha/transmittance missing");
    if (A.n_rays == 0) return IA_OK;
    composite_bwd_kernel<<<(unsigned)ia_ceil_div(A.n_rays, CP_WARPS), CP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        A, alpha, trans, g_weights, g_opacity, g_depth, g_comp_rgb, g_comp_normal, d_alpha_in, d_sdf, d_normal, d_inv_s,
        d_sigma, d_rgb, d_nrm);
    IA_LAUNCH_OK("composite_bwd_kernel");
    return IA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-ray mix of the foreground and background renders (reference models/neus.py:186 and :272-276):
//   comp_rgb_bg = comp_rgb_bg_raw + background_color (1 - opacity_bg),  comp_rgb_full = comp_rgb + comp_rgb_bg (1 - opacity),
//   rays_valid = opacity > 0, rays_valid_bg = opacity_bg > 0, rays_valid_full = their OR
// -- ten element-wise launches on [R, 3] tensors and a dozen in backward, one each here.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void ray_mix_fwd_kernel(const float *__restrict__ rgb, const float *__restrict__ op, const float *__restrict__ rgb_bg_raw,
                                   const float *__restrict__ op_bg, const float *__restrict__ bg_color, int64_t n,
                                   float *__restrict__ rgb_bg, float *__restrict__ rgb_full, uint8_t *__restrict__ valid,
                                   uint8_t *__restrict__ valid_bg, uint8_t *__restrict__ valid_full)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float o = op[i], ob = op_bg[i];
    const float t = 1.0f - o, tb = 1.0f - ob;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float b = __fadd_rn(rgb_bg_raw[3 * i + c], __fmul_rn(bg_color[c], tb));        // rounded as the tensor expression
        rgb_bg[3 * i + c] = b;
        rgb_full[3 * i + c] = __fadd_rn(rgb[3 * i + c], __fmul_rn(b, t));
    }
    const bool v = o > 0.f, vb = ob > 0.f;
    valid[i] = v;
    valid_bg[i] = vb;
    valid_full[i] = v || vb;
}

__global__ void ray_mix_bwd_kernel(const float *__restrict__ op, const float *__restrict__ op_bg, const float *__restrict__ bg_color,
                                   const float *__restrict__ rgb_bg, const float *__restrict__ g_bg, const float *__restrict__ g_full,
                                   int64_t n, float *__restrict__ d_rgb, float *__restrict__ d_op, float *__restrict__ d_rgb_bg_raw,
                                   float *__restrict__ d_op_bg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float t = 1.0f - op[i];
    float dop = 0.f, dopb = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float gf = g_full ? g_full[3 * i + c] : 0.f;
        const float gb = (g_bg ? g_bg[3 * i + c] : 0.f) + gf * t;        // total gradient of comp_rgb_bg
        if (d_rgb) d_rgb[3 * i + c] = gf;
        if (d_rgb_bg_raw) d_rgb_bg_raw[3 * i + c] = gb;
        dop -= gf * rgb_bg[3 * i + c];
        dopb -= gb * bg_color[c];
    }
    if (d_op) d_op[i] = dop;
    if (d_op_bg) d_op_bg[i] = dopb;
}
}  // namespace

extern "C" int32_t ia_ray_mix_fwd(const float *comp_rgb, const float *opacity, const float *comp_rgb_bg_raw, const float *opacity_bg,
                                  const float *background_color, int64_t n_rays, float *comp_rgb_bg, float *comp_rgb_full,
                                  uint8_t *rays_valid, uint8_t *rays_valid_bg, uint8_t *rays_valid_full, void *stream)
{
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (comp_rgb && opacity && comp_rgb_bg_raw && opacity_bg && background_color && comp_rgb_bg &&
                                               comp_rgb_full && rays_valid && rays_valid_bg && rays_valid_full)),
               "ray_mix_fwd: NULL pointer with n_rays=%lld", (long long)n_rays);
    if (n_rays == 0) return IA_OK;
    ray_mix_fwd_kernel<<<(unsigned)ia_ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        comp_rgb, opacity, comp_rgb_bg_raw, opacity_bg, background_color, n_rays, comp_rgb_bg, comp_rgb_full, rays_valid, rays_valid_bg,
        rays_valid_full);
    IA_LAUNCH_OK("ray_mix_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_ray_mix_bwd(const float *opacity, const float *opacity_bg, const float *background_color, const float *comp_rgb_bg,
                                  const float *g_comp_rgb_bg, const float *g_comp_rgb_full, int64_t n_rays, float *d_comp_rgb,
                                  float *d_opacity, float *d_comp_rgb_bg_raw, float *d_opacity_bg, void *stream)
{
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (opacity && opacity_bg && background_color && comp_rgb_bg)),
               "ray_mix_bwd: NULL pointer with n_rays=%lld", (long long)n_rays);
    if (n_rays == 0) return IA_OK;
    ray_mix_bwd_kernel<<<(unsigned)ia_ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        opacity, opacity_bg, background_color, comp_rgb_bg, g_comp_rgb_bg, g_comp_rgb_full, n_rays, d_comp_rgb, d_opacity,
        d_comp_rgb_bg_raw, d_opacity_bg);
    IA_LAUNCH_OK("ray_mix_bwd_kernel");
    return IA_OK;
}
