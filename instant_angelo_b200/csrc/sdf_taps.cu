// Fused elementwise stages of VolumeSDF's finite-difference gradient and curvature (reference
// models/geometry.py:219-275), forward and backward.  They replace ~60 eager PyTorch launches per step over
// [S,6,3]-sized tensors with 8 streaming kernels (each reads/writes every element exactly once):
//   fd_taps    : taps01[s,k,:] = ((clamp(base[s,:] + eps*e_k, -r, r)) + r) / (2r)             (geometry.py:221-232, 253-265)
//   fd_grad    : grad[s,i]     = 0.5 * (sdf[s,2i] - sdf[s,2i+1]) / eps                          (geometry.py:234, 267)
//   curv_shift : normals = normalize(grad); shifted = pts01 + cross(normals, normalize(rnd))*eps (geometry.py:238-246)
//   curv_angle : laplace = acos(clamp(normals . normalize(g_shift), +-(1-1e-6))) / pi            (geometry.py:269-275)
// Quirks of the reference are kept: taps are clamped in world space, the tangent is not normalised, the shifted point
// is built in normalised coordinates and then treated as a world-space point.
#include <math.h>

#include "ia_common.cuh"

namespace {

constexpr float NORM_EPS = 1e-12f;   // F.normalize default eps

__global__ void __launch_bounds__(256)
fd_taps_fwd_kernel(const float *__restrict__ base, int64_t n, float eps, float r, float *__restrict__ taps01)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float p[3] = {base[3 * s], base[3 * s + 1], base[3 * s + 2]};
    const float two_r = r - (-r);
    float *o = taps01 + 18 * s;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = p[d];
            if (d == (k >> 1)) v = v + ((k & 1) ? -eps : eps);
            v = fminf(fmaxf(v, -r), r);
            o[3 * k + d] = (v - (-r)) / two_r;
        }
    }
}

__global__ void __launch_bounds__(256)
fd_taps_bwd_kernel(const float *__restrict__ base, int64_t n, float eps, float r, const float *__restrict__ dtaps01,
                   float *__restrict__ dbase)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float p[3] = {base[3 * s], base[3 * s + 1], base[3 * s + 2]};
    const float two_r = r - (-r);
    const float *g = dtaps01 + 18 * s;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float v = p[d];
            if (d == (k >> 1)) v = v + ((k & 1) ? -eps : eps);
            if (v >= -r && v <= r) acc[d] += g[3 * k + d] / two_r;   // clamp passes the gradient on [min, max]
        }
    }
    dbase[3 * s] = acc[0];
    dbase[3 * s + 1] = acc[1];
    dbase[3 * s + 2] = acc[2];
}

__global__ void __launch_bounds__(256)
fd_grad_fwd_kernel(const float *__restrict__ sdf6, int64_t n, float eps, float *__restrict__ grad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (sample, axis)
    if (i >= 3 * n) return;
    const float2 pm = reinterpret_cast<const float2 *>(sdf6)[i];
    grad[i] = 0.5f * (pm.x - pm.y) / eps;
}

__global__ void __launch_bounds__(256)
fd_grad_bwd_kernel(const float *__restrict__ dgrad, int64_t n, float eps, float *__restrict__ dsdf6)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * n) return;
    const float g = dgrad[i] / eps * 0.5f;
    reinterpret_cast<float2 *>(dsdf6)[i] = make_float2(g, -g);
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 ld3(const float *p, int64_t s) { return {p[3 * s], p[3 * s + 1], p[3 * s + 2]}; }
__device__ __forceinline__ void st3(float *p, int64_t s, V3 v) { p[3 * s] = v.x; p[3 * s + 1] = v.y; p[3 * s + 2] = v.z; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ V3 scale3(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 normalize3(V3 a, float *len)
{
    const float l = fmaxf(sqrtf(dot3(a, a)), NORM_EPS);
    *len = l;
    return {a.x / l, a.y / l, a.z / l};
}
// gradient of n = a / max(|a|, eps) w.r.t. a, given dn (the clamp branch |a| < eps is treated as a constant divisor)
__device__ __forceinline__ V3 normalize3_bwd(V3 n, float len, V3 dn, bool tiny)
{
    if (tiny) return scale3(dn, 1.0f / len);
    const float nd = dot3(n, dn);
    return {(dn.x - n.x * nd) / len, (dn.y - n.y * nd) / len, (dn.z - n.z * nd) / len};
}

__global__ void __launch_bounds__(256)
curv_shift_fwd_kernel(const float *__restrict__ grad, const float *__restrict__ rnd, const float *__restrict__ pts01, int64_t n,
                      float eps, float *__restrict__ normals, float *__restrict__ shifted)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float lg, lr;
    const V3 nrm = normalize3(ld3(grad, s), &lg);
    const V3 u = normalize3(ld3(rnd, s), &lr);
    const V3 t = cross3(nrm, u);
    const V3 p = ld3(pts01, s);
    st3(normals, s, nrm);
    st3(shifted, s, {p.x + t.x * eps, p.y + t.y * eps, p.z + t.z * eps});
}

__global__ void __launch_bounds__(256)
curv_shift_bwd_kernel(const float *__restrict__ grad, const float *__restrict__ rnd, int64_t n, float eps,
                      const float *__restrict__ dnormals, const float *__restrict__ dshifted, float *__restrict__ dgrad)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const V3 g = ld3(grad, s);
    float lg, lr;
    const V3 nrm = normalize3(g, &lg);
    const V3 u = normalize3(ld3(rnd, s), &lr);
    V3 dn = dnormals ? ld3(dnormals, s) : V3{0.f, 0.f, 0.f};
    if (dshifted) {
        const V3 dt = scale3(ld3(dshifted, s), eps);     // shifted = p + t*eps, t = n x u  =>  dn += u x dt
        const V3 c = cross3(u, dt);
        dn.x += c.x; dn.y += c.y; dn.z += c.z;
    }
    st3(dgrad, s, normalize3_bwd(nrm, lg, dn, sqrtf(dot3(g, g)) < NORM_EPS));
}

__global__ void __launch_bounds__(256)
curv_angle_fwd_kernel(const float *__restrict__ normals, const float *__restrict__ gshift, int64_t n, float *__restrict__ laplace)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float l;
    const V3 ns = normalize3(ld3(gshift, s), &l);
    const float d = dot3(ld3(normals, s), ns);
    const float c = fminf(fmaxf(d, -1.0f + 1e-6f), 1.0f - 1e-6f);
    laplace[s] = acosf(c) / 3.14159265358979323846f;
}

__global__ void __launch_bounds__(256)
curv_angle_bwd_kernel(const float *__restrict__ normals, const float *__restrict__ gshift, int64_t n,
                      const float *__restrict__ dlaplace, float *__restrict__ dnormals, float *__restrict__ dgshift)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const V3 gs = ld3(gshift, s);
    float l;
    const V3 ns = normalize3(gs, &l);
    const V3 nrm = ld3(normals, s);
    const float d = dot3(nrm, ns);
    const float lo = -1.0f + 1e-6f, hi = 1.0f - 1e-6f;
    float dd = 0.f;
    if (d >= lo && d <= hi) dd = -dlaplace[s] / (3.14159265358979323846f * sqrtf(1.0f - d * d));
    st3(dnormals, s, scale3(ns, dd));
    st3(dgshift, s, normalize3_bwd(ns, l, scale3(nrm, dd), sqrtf(dot3(gs, gs)) < NORM_EPS));
}

inline unsigned blocks_for(int64_t n) { return (unsigned)ia_ceil_div(n, 256); }

}  // namespace

extern "C" int32_t ia_fd_taps_fwd(const float *base, int64_t n, float eps, float radius, float *taps01, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (base && taps01)), "fd_taps_fwd: NULL pointer");
    IA_REQUIRE(radius > 0.f, "fd_taps_fwd: radius must be > 0");
    if (n == 0) return IA_OK;
    fd_taps_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(base, n, eps, radius, taps01);
    IA_LAUNCH_OK("fd_taps_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_fd_taps_bwd(const float *base, int64_t n, float eps, float radius, const float *dtaps01, float *dbase,
                                  void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (base && dtaps01 && dbase)), "fd_taps_bwd: NULL pointer");
    if (n == 0) return IA_OK;
    fd_taps_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(base, n, eps, radius, dtaps01, dbase);
    IA_LAUNCH_OK("fd_taps_bwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_fd_grad_fwd(const float *sdf6, int64_t n, float eps, float *grad, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (sdf6 && grad)), "fd_grad_fwd: NULL pointer");
    IA_REQUIRE(eps > 0.f, "fd_grad_fwd: eps must be > 0");
    if (n == 0) return IA_OK;
    fd_grad_fwd_kernel<<<blocks_for(3 * n), 256, 0, (cudaStream_t)stream>>>(sdf6, n, eps, grad);
    IA_LAUNCH_OK("fd_grad_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_fd_grad_bwd(const float *dgrad, int64_t n, float eps, float *dsdf6, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (dgrad && dsdf6)), "fd_grad_bwd: NULL pointer");
    if (n == 0) return IA_OK;
    fd_grad_bwd_kernel<<<blocks_for(3 * n), 256, 0, (cudaStream_t)stream>>>(dgrad, n, eps, dsdf6);
    IA_LAUNCH_OK("fd_grad_bwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_curv_shift_fwd(const float *grad, const float *rnd, const float *pts01, int64_t n, float eps,
                                     float *normals, float *shifted, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (grad && rnd && pts01 && normals && shifted)), "curv_shift_fwd: NULL pointer");
    if (n == 0) return IA_OK;
    curv_shift_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(grad, rnd, pts01, n, eps, normals, shifted);
    IA_LAUNCH_OK("curv_shift_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_curv_shift_bwd(const float *grad, const float *rnd, int64_t n, float eps, const float *dnormals,
                                     const float *dshifted, float *dgrad, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (grad && rnd && dgrad)), "curv_shift_bwd: NULL pointer");
    if (n == 0) return IA_OK;
    curv_shift_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(grad, rnd, n, eps, dnormals, dshifted, dgrad);
    IA_LAUNCH_OK("curv_shift_bwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_curv_angle_fwd(const float *normals, const float *gshift, int64_t n, float *laplace, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (normals && gshift && laplace)), "curv_angle_fwd: NULL pointer");
    if (n == 0) return IA_OK;
    curv_angle_fwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(normals, gshift, n, laplace);
    IA_LAUNCH_OK("curv_angle_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_curv_angle_bwd(const float *normals, const float *gshift, int64_t n, const float *dlaplace,
                                     float *dnormals, float *dgshift, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (normals && gshift && dlaplace && dnormals && dgshift)), "curv_angle_bwd: NULL pointer");
    if (n == 0) return IA_OK;
    curv_angle_bwd_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(normals, gshift, n, dlaplace, dnormals, dgshift);
    IA_LAUNCH_OK("curv_angle_bwd_kernel");
    return IA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sample points of the marched intervals and the normalisation of the SDF gradient: the tensor expressions between the
// marcher and the networks (reference models/neus.py:153-157, 218-223: t_origins / t_dirs gathers, midpoints, positions,
// dists; :229 F.normalize), one launch each instead of ~9 / ~12 (forward + backward) element-wise launches per call.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// positions = rays_o[ri] + rays_d[ri] * ((t0 + t1) / 2), rounded operation by operation as the tensor expression is (no
// fused multiply-add): the background marcher prunes samples on a density evaluated at these positions, and the sample
// set is a bit-exact contract
__global__ void ray_samples_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const int32_t *__restrict__ ri,
                                   const float *__restrict__ t0, const float *__restrict__ t1, int64_t n, float *__restrict__ pos,
                                   float *__restrict__ dirs, float *__restrict__ mid, float *__restrict__ dist)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = ri[i];
    const float a = t0[i], b = t1[i];
    const float m = __fmul_rn(__fadd_rn(a, b), 0.5f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float d = __ldg(rays_d + 3 * r + c);
        pos[3 * i + c] = __fadd_rn(__ldg(rays_o + 3 * r + c), __fmul_rn(d, m));
        if (dirs) dirs[3 * i + c] = d;
    }
    if (mid) mid[i] = m;
    if (dist) dist[i] = __fsub_rn(b, a);
}

__global__ void normalize3_fwd_kernel(const float *__restrict__ x, int64_t n, float eps, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
    const float den = fmaxf(sqrtf(a * a + b * b + c * c), eps);      // F.normalize: x / max(|x|, eps)
    out[3 * i] = a / den;
    out[3 * i + 1] = b / den;
    out[3 * i + 2] = c / den;
}

// y = x / max(|x|, eps):  |x| > eps: dx = (g - y (y . g)) / |x| ;  otherwise (the clamp passes no gradient to the norm) dx = g / eps
__global__ void normalize3_bwd_kernel(const float *__restrict__ x, const float *__restrict__ g, int64_t n, float eps,
                                      float *__restrict__ dx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
    const float ga = g[3 * i], gb = g[3 * i + 1], gc = g[3 * i + 2];
    const float nrm = sqrtf(a * a + b * b + c * c);
    if (nrm > eps) {
        const float inv = 1.f / nrm;
        const float ya = a * inv, yb = b * inv, yc = c * inv;
        const float dot = ya * ga + yb * gb + yc * gc;
        dx[3 * i] = (ga - ya * dot) * inv;
        dx[3 * i + 1] = (gb - yb * dot) * inv;
        dx[3 * i + 2] = (gc - yc * dot) * inv;
    } else {
        const float inv = 1.f / eps;
        dx[3 * i] = ga * inv;
        dx[3 * i + 1] = gb * inv;
        dx[3 * i + 2] = gc * inv;
    }
}

}  // namespace

extern "C" int32_t ia_ray_samples(const float *rays_o, const float *rays_d, const int32_t *ray_indices, const float *t_starts,
                                  const float *t_ends, int64_t n, float *positions, float *t_dirs, float *midpoints, float *dists,
                                  void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (rays_o && rays_d && ray_indices && t_starts && t_ends && positions)),
               "ray_samples: NULL pointer with n=%lld", (long long)n);
    if (n == 0) return IA_OK;
    ray_samples_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, ray_indices, t_starts, t_ends, n,
                                                                                        positions, t_dirs, midpoints, dists);
    IA_LAUNCH_OK("ray_samples_kernel");
    return IA_OK;
}

extern "C" int32_t ia_normalize3_fwd(const float *x, int64_t n, float eps, float *out, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (x && out)) && eps > 0.f, "normalize3_fwd: bad arguments (n=%lld)", (long long)n);
    if (n == 0) return IA_OK;
    normalize3_fwd_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, eps, out);
    IA_LAUNCH_OK("normalize3_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_normalize3_bwd(const float *x, const float *dout, int64_t n, float eps, float *dx, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (x && dout && dx)) && eps > 0.f, "normalize3_bwd: bad arguments (n=%lld)", (long long)n);
    if (n == 0) return IA_OK;
    normalize3_bwd_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, dout, n, eps, dx);
    IA_LAUNCH_OK("normalize3_bwd_kernel");
    return IA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// contract_to_unisphere (reference models/geometry.py:19-31) for sample positions that carry no gradient: the AABB
// normalisation (x + r) / 2r, or the un-bounded-sphere contraction of the background model -- as tensor expressions 3 and
// 11 element-wise launches per call, on up to 8 M rows when the background occupancy grid refreshes.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
__global__ void contract_kernel(const float *__restrict__ x, int64_t n, float r, int type, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float two_r = r - (-r);
    float v[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) v[d] = (x[3 * i + d] - (-r)) / two_r;             // scale_anything(x, (-r, r), (0, 1))
    if (type == IA_UN_BOUNDED_SPHERE) {
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = v[d] * 2.f - 1.f;
        const float mag = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (mag > 1.f) {
            const float k = 2.f - 1.f / mag;
#pragma unroll
            for (int d = 0; d < 3; ++d) v[d] = k * (v[d] / mag);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = v[d] / 4.f + 0.5f;
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) out[3 * i + d] = v[d];
}
}  // namespace

extern "C" int32_t ia_contract(const float *x, int64_t n, float radius, int32_t contraction_type, float *out, void *stream)
{
    IA_REQUIRE(n >= 0 && (n == 0 || (x && out)), "contract: NULL pointer with n=%lld", (long long)n);
    IA_REQUIRE(radius > 0.f, "contract: radius must be > 0");
    IA_REQUIRE(contraction_type == IA_AABB || contraction_type == IA_UN_BOUNDED_SPHERE, "contract: unsupported contraction type %d",
               contraction_type);
    if (n == 0) return IA_OK;
    contract_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, radius, contraction_type, out);
    IA_LAUNCH_OK("contract_kernel");
    return IA_OK;
}
