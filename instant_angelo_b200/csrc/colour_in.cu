// Input row of the colour head when the SDF network's wide output layer is FOLDED into the colour network's first layer.
//
// Reference: feature = cat[out, points*2-1] (models/geometry.py:206-207), network_inp = cat[feature, dirs_embd, normals]
// (models/texture.py:26-27), colour layer 0: z = Wc0 network_inp + bc0.  With out = Wl h + bl (h = last hidden layer of the SDF
// network) the first 65 columns of Wc0 compose with Wl:  z = (Wc0[:, :65] Wl) h + Wc0[:, 65:] [pts | enc | normal] + (bc0 + Wc0[:, :65] bl),
// so the colour network can read h directly and the [N, 65] geometry output is never formed.  Only the four columns that are used
// outside the colour network are: sdf = out[:, 0] and the dual-colour diffuse term out[:, 1:4] (models/texture.py:58).
//
//   forward : tin[N, ld] = [h (64) | pts01*2-1 (3) | enc (n_enc) | normal (3) | zero padding],  out4[N, 4] = h W4^T + b4
//   backward: dh = dtin[:, :64] + dout4 W4;  dpts01 = 2 dtin[:, 64:67];  denc, dnormal = column blocks;  dW4 += dout4^T h;  db4 += sum dout4
// HBM streaming: 64-row tiles staged through shared memory, every global access a whole line; dW4 accumulates in one register per
// thread across the tiles of a persistent CTA.
#include <algorithm>

#include "ia_common.cuh"

namespace {

constexpr int CI_THREADS = 256;
constexpr int CI_W = 64;
constexpr int CI_T = 64;          // rows per tile
constexpr int CI_HS = 68;         // smem row stride of the h tile: 16-byte aligned rows, (4 r + k) % 32 distinct for 8 rows of a warp

// extras block of a row: columns [64, ld) of tin = [pts01*2-1 | enc | normal | zeros]
__device__ __forceinline__ float extra_value(int c, int64_t row, const float *__restrict__ pts01, const float *__restrict__ enc, int n_enc,
                                             const float *__restrict__ normal)
{
    if (c < 3) return fmaf(__ldg(pts01 + row * 3 + c), 2.f, -1.f);
    if (c < 3 + n_enc) return __ldg(enc + row * n_enc + (c - 3));
    if (c < 6 + n_enc) return __ldg(normal + row * 3 + (c - 3 - n_enc));
    return 0.f;
}

__global__ void __launch_bounds__(CI_THREADS)
colour_in_fwd_kernel(const float *__restrict__ h, int64_t n, const float *__restrict__ W4, const float *__restrict__ b4,
                     const float *__restrict__ pts01, const float *__restrict__ enc, int n_enc, const float *__restrict__ normal,
                     float *__restrict__ tin, int ld, float *__restrict__ sdf, float *__restrict__ rgb_raw)
{
    __shared__ __align__(16) float hs[CI_T * CI_HS];
    __shared__ float w4s[4 * 65];
    const int tid = threadIdx.x;
    w4s[(tid >> 6) * 65 + (tid & 63)] = __ldg(W4 + tid);
    const int r_dot = tid >> 2, o_dot = tid & 3;
    const float bias = __ldg(b4 + o_dot);
    const int n_x = ld - CI_W;
    const int64_t n_tiles = (n + CI_T - 1) / CI_T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * CI_T;
        __syncthreads();                 // previous tile's dot products have read hs (and w4s is written)
#pragma unroll
        for (int i = 0; i < CI_T * 16 / CI_THREADS; ++i) {
            const int idx = tid + i * CI_THREADS, r = idx >> 4, q = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (base + r < n) {
                v = __ldg(reinterpret_cast<const float4 *>(h + (base + r) * CI_W) + q);
                *reinterpret_cast<float4 *>(tin + (base + r) * ld + 4 * q) = v;
            }
            *reinterpret_cast<float4 *>(&hs[r * CI_HS + 4 * q]) = v;
        }
        for (int idx = tid; idx < CI_T * n_x; idx += CI_THREADS) {
            const int r = idx / n_x, c = idx - r * n_x;
            if (base + r < n) tin[(base + r) * ld + CI_W + c] = extra_value(c, base + r, pts01, enc, n_enc, normal);
        }
        __syncthreads();
        float acc = bias;
#pragma unroll 16
        for (int k = 0; k < CI_W; ++k) acc = fmaf(hs[r_dot * CI_HS + k], w4s[o_dot * 65 + k], acc);
        const int64_t row = base + r_dot;
        if (row < n) {
            if (o_dot == 0) sdf[row] = acc;
            else rgb_raw[row * 3 + (o_dot - 1)] = acc;
        }
    }
}

__global__ void __launch_bounds__(CI_THREADS)
colour_in_bwd_kernel(const float *__restrict__ h, int64_t n, const float *__restrict__ W4, const float *__restrict__ dtin, int ld,
                     int n_enc, const float *__restrict__ dsdf, const float *__restrict__ drgb, float *__restrict__ dh,
                     float *__restrict__ dW4, float *__restrict__ db4, float *__restrict__ dpts01, float *__restrict__ denc,
                     float *__restrict__ dnormal)
{
    __shared__ __align__(16) float hs[CI_T * CI_HS];
    __shared__ __align__(16) float w4s[4 * CI_W];
    __shared__ float d4s[CI_T * 4];
    const int tid = threadIdx.x;
    w4s[tid] = __ldg(W4 + tid);
    const int o_w = tid >> 6, k_w = tid & 63;      // this thread's dW4 entry
    float gw = 0.f, gb = 0.f;
    const int n_extra = 3 + n_enc + 3;
    const int64_t n_tiles = (n + CI_T - 1) / CI_T;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * CI_T;
        __syncthreads();                 // previous tile's accumulation has read hs / d4s
        {
            const int r = tid >> 2, o = tid & 3;
            const int64_t row = base + r;
            float v = 0.f;
            if (row < n) v = o == 0 ? (dsdf ? __ldg(dsdf + row) : 0.f) : (drgb ? __ldg(drgb + row * 3 + (o - 1)) : 0.f);
            d4s[r * 4 + o] = v;
        }
        if (dW4) {
#pragma unroll
            for (int i = 0; i < CI_T * 16 / CI_THREADS; ++i) {
                const int idx = tid + i * CI_THREADS, r = idx >> 4, q = idx & 15;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + r < n) v = __ldg(reinterpret_cast<const float4 *>(h + (base + r) * CI_W) + q);
                *reinterpret_cast<float4 *>(&hs[r * CI_HS + 4 * q]) = v;
            }
        }
        __syncthreads();
        if (dh) {
#pragma unroll
            for (int i = 0; i < CI_T * 16 / CI_THREADS; ++i) {
                const int idx = tid + i * CI_THREADS, r = idx >> 4, q = idx & 15;
                if (base + r < n) {
                    float4 g = __ldg(reinterpret_cast<const float4 *>(dtin + (base + r) * ld) + q);
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        const float d = d4s[r * 4 + o];
                        const float4 w = *reinterpret_cast<const float4 *>(&w4s[o * CI_W + 4 * q]);
                        g.x = fmaf(d, w.x, g.x); g.y = fmaf(d, w.y, g.y); g.z = fmaf(d, w.z, g.z); g.w = fmaf(d, w.w, g.w);
                    }
                    *reinterpret_cast<float4 *>(dh + (base + r) * CI_W + 4 * q) = g;
                }
            }
        }
        for (int idx = tid; idx < CI_T * n_extra; idx += CI_THREADS) {
            const int r = idx / n_extra, c = idx - r * n_extra;
            const int64_t row = base + r;
            if (row < n) {
                const float v = __ldg(dtin + row * ld + CI_W + c);
                if (c < 3) { if (dpts01) dpts01[row * 3 + c] = 2.f * v; }
                else if (c < 3 + n_enc) { if (denc) denc[row * n_enc + (c - 3)] = v; }
                else if (dnormal) dnormal[row * 3 + (c - 3 - n_enc)] = v;
            }
        }
        if (dW4) {
#pragma unroll 16
            for (int r = 0; r < CI_T; ++r) {
                const float d = d4s[r * 4 + o_w];
                gw = fmaf(d, hs[r * CI_HS + k_w], gw);
                gb += d;
            }
        }
    }
    if (dW4) {
        atomicAdd(dW4 + o_w * CI_W + k_w, gw);
        if (db4 && k_w == 0) atomicAdd(db4 + o_w, gb);
    }
}

int check_args(int64_t n, int32_t n_enc, int64_t ld)
{
    IA_REQUIRE(n >= 0, "ia_colour_in: n < 0");
    IA_REQUIRE(n_enc >= 0 && n_enc <= 64, "ia_colour_in: n_enc %d out of range", n_enc);
    IA_REQUIRE(ld >= CI_W + 3 + n_enc + 3 && (ld & 3) == 0, "ia_colour_in: row stride %lld must be a multiple of 4 and >= 64 + 3 + n_enc + 3",
               (long long)ld);
    return IA_OK;
}

}  // namespace

extern "C" int32_t ia_colour_in_fwd(const float *h, int64_t n, const float *W4, const float *b4, const float *pts01,
                                    const float *enc, int32_t n_enc, const float *normal, float *tin, int64_t ld, float *sdf,
                                    float *rgb_raw, void *stream)
{
    const int rc = check_args(n, n_enc, ld);
    if (rc != IA_OK) return rc;
    if (n == 0) return IA_OK;
    IA_REQUIRE(h && W4 && b4 && pts01 && normal && tin && sdf && rgb_raw && (enc || n_enc == 0), "ia_colour_in_fwd: NULL pointer");
    const unsigned blocks = (unsigned)std::min<int64_t>(ia_ceil_div(n, CI_T), (int64_t)ia_sm_count() * 6);
    colour_in_fwd_kernel<<<blocks, CI_THREADS, 0, (cudaStream_t)stream>>>(h, n, W4, b4, pts01, enc, n_enc, normal, tin, (int)ld, sdf, rgb_raw);
    IA_LAUNCH_OK("colour_in_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_colour_in_bwd(const float *h, int64_t n, const float *W4, const float *dtin, int64_t ld, int32_t n_enc,
                                    const float *dsdf, const float *drgb, float *dh, float *dW4, float *db4, float *dpts01,
                                    float *denc, float *dnormal, void *stream)
{
    const int rc = check_args(n, n_enc, ld);
    if (rc != IA_OK) return rc;
    if (n == 0) return IA_OK;
    IA_REQUIRE(h && W4 && dtin, "ia_colour_in_bwd: NULL pointer");
    const unsigned blocks = (unsigned)std::min<int64_t>(ia_ceil_div(n, CI_T), (int64_t)ia_sm_count() * 6);
    colour_in_bwd_kernel<<<blocks, CI_THREADS, 0, (cudaStream_t)stream>>>(h, n, W4, dtin, (int)ld, n_enc, dsdf, drgb, dh, dW4, db4, dpts01,
                                                                         denc, dnormal);
    IA_LAUNCH_OK("colour_in_bwd_kernel");
    return IA_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// The fold itself, in parameter space: the colour network's flat parameter vector with the geometry output layer composed
// into its first layer (texture.forward_fused_head):
//   W_eff[o, :] = [ sum_k Wc0[o,k] Wl[k,:]  (64 columns) | Wc0[o, n_feat:] | 0-pad ],  b_eff[o] = bc0[o] + sum_k Wc0[o,k] bl[k],
// followed by the untouched rest of the colour network.  As tensor expressions (broadcast product + sum + cat, and their
// autograd) this is 8 launches forward and 28 in backward for a 64 x 65 x 64 product of parameters.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int FH_W = 64;

__global__ void fold_head_fwd_kernel(const float *__restrict__ flat, const float *__restrict__ wl, const float *__restrict__ bl,
                                     int n_in, int n_feat, int ld, int n_rest, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_w = FH_W * ld;
    if (i < n_w) {
        const int o = i / ld, c = i - o * ld;
        const float *w = flat + o * n_in;
        float v = 0.f;
        if (c < FH_W) {
            for (int k = 0; k < n_feat; ++k) v = fmaf(w[k], wl[k * FH_W + c], v);
        } else if (c - FH_W < n_in - n_feat) {
            v = w[n_feat + (c - FH_W)];
        }
        out[i] = v;
    } else if (i < n_w + FH_W) {
        const int o = i - n_w;
        const float *w = flat + o * n_in;
        float v = flat[FH_W * n_in + o];
        for (int k = 0; k < n_feat; ++k) v = fmaf(w[k], bl[k], v);
        out[i] = v;
    } else if (i < n_w + FH_W + n_rest) {
        out[i] = flat[FH_W * n_in + FH_W + (i - n_w - FH_W)];
    }
}

// g = d(out).  dflat is ADDED to (it may be the accumulator behind the colour network's flat vector), dwl / dbl are written.
__global__ void fold_head_bwd_kernel(const float *__restrict__ g, const float *__restrict__ flat, const float *__restrict__ wl,
                                     const float *__restrict__ bl, int n_in, int n_feat, int ld, int n_rest,
                                     float *__restrict__ dflat, float *__restrict__ dwl, float *__restrict__ dbl)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_w0 = FH_W * n_in, n_w = FH_W * ld;
    const float *gb = g + n_w;
    if (i < n_w0) {                                          // d Wc0[o, k]
        const int o = i / n_in, k = i - o * n_in;
        float v;
        if (k < n_feat) {
            v = gb[o] * bl[k];
            const float *gm = g + o * ld, *w = wl + k * FH_W;
            for (int c = 0; c < FH_W; ++c) v = fmaf(gm[c], w[c], v);
        } else {
            v = g[o * ld + FH_W + (k - n_feat)];
        }
        if (dflat) dflat[i] += v;
    } else if (i < n_w0 + FH_W) {                            // d bc0
        if (dflat) dflat[i] += gb[i - n_w0];
    } else if (i < n_w0 + FH_W + n_rest) {                   // rest of the colour network
        if (dflat) dflat[i] += g[n_w + FH_W + (i - n_w0 - FH_W)];
    } else if (i < n_w0 + FH_W + n_rest + n_feat * FH_W) {   // d Wl[k, c]
        const int j = i - (n_w0 + FH_W + n_rest), k = j / FH_W, c = j - k * FH_W;
        float v = 0.f;
        for (int o = 0; o < FH_W; ++o) v = fmaf(g[o * ld + c], flat[o * n_in + k], v);
        if (dwl) dwl[j] = v;
    } else if (i < n_w0 + FH_W + n_rest + n_feat * FH_W + n_feat) {   // d bl[k]
        const int k = i - (n_w0 + FH_W + n_rest + n_feat * FH_W);
        float v = 0.f;
        for (int o = 0; o < FH_W; ++o) v = fmaf(gb[o], flat[o * n_in + k], v);
        if (dbl) dbl[k] = v;
    }
}

int fold_check(const float *flat, const float *wl, const float *bl, int n_in, int n_feat, int ld, int64_t n_rest, const char *who)
{
    IA_REQUIRE(flat && wl && bl, "%s: NULL pointer", who);
    IA_REQUIRE(n_feat >= 1 && n_feat <= n_in && ld >= FH_W + (n_in - n_feat) && n_rest >= 0 && n_rest < (1 << 24) && ld < 4096,
               "%s: bad shape (n_in %d, n_feat %d, ld %d)", who, n_in, n_feat, ld);
    return IA_OK;
}
}  // namespace

extern "C" int32_t ia_fold_head_fwd(const float *flat, const float *w_last, const float *b_last, int32_t n_in, int32_t n_feat,
                                    int32_t ld, int64_t n_rest, float *flat_eff, void *stream)
{
    int rc = fold_check(flat, w_last, b_last, n_in, n_feat, ld, n_rest, "fold_head_fwd");
    if (rc) return rc;
    IA_REQUIRE(flat_eff != nullptr, "fold_head_fwd: flat_eff is NULL");
    const int total = FH_W * ld + FH_W + (int)n_rest;
    fold_head_fwd_kernel<<<(unsigned)ia_ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(flat, w_last, b_last, n_in, n_feat, ld,
                                                                                               (int)n_rest, flat_eff);
    IA_LAUNCH_OK("fold_head_fwd_kernel");
    return IA_OK;
}

extern "C" int32_t ia_fold_head_bwd(const float *dflat_eff, const float *flat, const float *w_last, const float *b_last, int32_t n_in,
                                    int32_t n_feat, int32_t ld, int64_t n_rest, float *dflat, float *dw_last, float *db_last,
                                    void *stream)
{
    int rc = fold_check(flat, w_last, b_last, n_in, n_feat, ld, n_rest, "fold_head_bwd");
    if (rc) return rc;
    IA_REQUIRE(dflat_eff != nullptr, "fold_head_bwd: dflat_eff is NULL");
    const int total = FH_W * n_in + FH_W + (int)n_rest + n_feat * FH_W + n_feat;
    fold_head_bwd_kernel<<<(unsigned)ia_ceil_div(total, 128), 128, 0, (cudaStream_t)stream>>>(dflat_eff, flat, w_last, b_last, n_in, n_feat,
                                                                                               ld, (int)n_rest, dflat, dw_last, db_last);
    IA_LAUNCH_OK("fold_head_bwd_kernel");
    return IA_OK;
}
