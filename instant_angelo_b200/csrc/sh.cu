// Real spherical harmonics of directions (degree <= 4), forward and input-gradient.
// Replaces tcnn.Encoding(otype=SphericalHarmonics) (reference models/network_utils.py:90-91, called from
// models/texture.py:25, 52, 129, 134); basis per SURVEY.md Appendix A.2.  Pure streaming: 12 B in,
// 4*degree^2 B out per direction.
#include "ia_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
sh_fwd_kernel(const float *__restrict__ d01, int64_t n, int degree, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = d01[3 * i] * 2.f - 1.f, y = d01[3 * i + 1] * 2.f - 1.f, z = d01[3 * i + 2] * 2.f - 1.f;
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    float *o = out + i * (degree * degree);
    o[0] = 0.28209479177387814f;
    if (degree <= 1) return;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    if (degree <= 2) return;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    if (degree <= 3) return;
    o[9] = 0.59004358992664352f * y * (-3.f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.f - 5.f * z2);
    o[12] = 0.3731763325901154f * z * (5.f * z2 - 3.f);
    o[13] = 0.45704579946446572f * x * (1.f - 5.f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.f * y2);
}

__global__ void __launch_bounds__(256)
sh_bwd_kernel(const float *__restrict__ d01, int64_t n, int degree, const float *__restrict__ dout,
              float *__restrict__ dd01)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = d01[3 * i] * 2.f - 1.f, y = d01[3 * i + 1] * 2.f - 1.f, z = d01[3 * i + 2] * 2.f - 1.f;
    const float x2 = x * x, y2 = y * y, z2 = z * z;
    const float *g = dout + i * (degree * degree);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (degree > 1) {
        gy += -0.48860251190291987f * g[1];
        gz += 0.48860251190291987f * g[2];
        gx += -0.48860251190291987f * g[3];
    }
    if (degree > 2) {
        gx += 1.0925484305920792f * y * g[4];
        gy += 1.0925484305920792f * x * g[4];
        gy += -1.0925484305920792f * z * g[5];
        gz += -1.0925484305920792f * y * g[5];
        gz += 2.f * 0.94617469575755997f * z * g[6];
        gx += -1.0925484305920792f * z * g[7];
        gz += -1.0925484305920792f * x * g[7];
        gx += 2.f * 0.54627421529603959f * x * g[8];
        gy += -2.f * 0.54627421529603959f * y * g[8];
    }
    if (degree > 3) {
        const float c9 = 0.59004358992664352f, c10 = 2.8906114426405538f, c11 = 0.45704579946446572f;
        const float c12 = 0.3731763325901154f, c14 = 1.4453057213202769f;
        // 9: c9*y*(-3x^2+y^2)
        gx += c9 * (-6.f * x * y) * g[9];
        gy += c9 * (-3.f * x2 + 3.f * y2) * g[9];
        // 10: c10*x*y*z
        gx += c10 * y * z * g[10];
        gy += c10 * x * z * g[10];
        gz += c10 * x * y * g[10];
        // 11: c11*y*(1-5z^2)
        gy += c11 * (1.f - 5.f * z2) * g[11];
        gz += c11 * y * (-10.f * z) * g[11];
        // 12: c12*z*(5z^2-3)
        gz += c12 * (15.f * z2 - 3.f) * g[12];
        // 13: c11*x*(1-5z^2)
        gx += c11 * (1.f - 5.f * z2) * g[13];
        gz += c11 * x * (-10.f * z) * g[13];
        // 14: c14*z*(x^2-y^2)
        gx += c14 * 2.f * x * z * g[14];
        gy += -c14 * 2.f * y * z * g[14];
        gz += c14 * (x2 - y2) * g[14];
        // 15: c9*x*(-x^2+3y^2)
        gx += c9 * (-3.f * x2 + 3.f * y2) * g[15];
        gy += c9 * 6.f * x * y * g[15];
    }
    // d(2*d01-1)/d d01 = 2
    dd01[3 * i] = 2.f * gx;
    dd01[3 * i + 1] = 2.f * gy;
    dd01[3 * i + 2] = 2.f * gz;
}

}  // namespace

extern "C" int32_t ia_sh_fwd(const float *d01, int64_t n, int32_t degree, float *out, void *stream)
{
    IA_REQUIRE(degree >= 1 && degree <= 4, "sh_fwd: degree %d not in [1,4]", degree);
    IA_REQUIRE(n >= 0 && (n == 0 || (d01 && out)), "sh_fwd: NULL pointer");
    if (n == 0) return IA_OK;
    sh_fwd_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d01, n, degree, out);
    IA_LAUNCH_OK("sh_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_sh_bwd(const float *d01, int64_t n, int32_t degree, const float *dout, float *dd01, void *stream)
{
    IA_REQUIRE(degree >= 1 && degree <= 4, "sh_bwd: degree %d not in [1,4]", degree);
    IA_REQUIRE(n >= 0 && (n == 0 || (d01 && dout && dd01)), "sh_bwd: NULL pointer");
    if (n == 0) return IA_OK;
    sh_bwd_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d01, n, degree, dout, dd01);
    IA_LAUNCH_OK("sh_bwd_kernel");
    return IA_OK;
}
