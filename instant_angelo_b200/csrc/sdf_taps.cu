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
