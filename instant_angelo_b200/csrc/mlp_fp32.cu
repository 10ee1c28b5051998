// Width-64 fully fused MLP, fp32 FFMA path (bit-faithful parity mode), forward and backward.
// Replaces VanillaMLP.forward (reference models/network_utils.py:96-113: nn.Linear stack, fp32, autocast
// disabled, Softplus(beta=100) for sphere-init nets else ReLU) for geometry 35->64->64->65, colour
// 87->64->64->3, UniSDF heads, background nets.
//
// One persistent CTA per SM keeps every weight matrix of the network resident in shared memory and
// streams 128-row tiles: input tile -> smem (feature-major, XOR-swizzled), each layer is a
// 128 x 64 x K register-tiled GEMM (8x4 micro-tile, LDS.128 operands) whose epilogue (bias + activation)
// writes the next layer's operand back to smem, so hidden activations never touch HBM.  Backward
// recomputes the hidden activations in-kernel, chains dZ through the layers in place, and accumulates
// dW / db in registers across all tiles of the CTA (one atomic flush per CTA at the end).
// The tensor-core (tcgen05) variant lives in mlp_tc.cu; this file is the precision reference for it.
#include <math.h>
#include <algorithm>

#include "ia_common.cuh"

namespace {

constexpr int MT = 128;        // rows per tile
constexpr int NT = 256;        // threads per CTA: 16 (ty) x 16 (tx)
constexpr int W = 64;          // hidden width
constexpr float BETA = 100.f;  // Softplus beta

struct MlpDims {
    int n_in0, n_in1, din, K0p;  // K0p = din rounded up to 4
    float s0, o0;
    int nh;        // hidden layers (1|2)
    int n_out;     // network outputs
    int nou;       // outputs used
    int act;       // IA_ACT_RELU | IA_ACT_SOFTPLUS100
    // parameter offsets (floats) in the flat layout
    int pW0, pb0, pW1, pb1, pWl, pbl;
};

__device__ __forceinline__ int kidx(int k, int m) { return k * MT + ((((m >> 2) ^ ((k >> 2) & 7))) << 2) + (m & 3); }

__device__ __forceinline__ float act_fwd(float z, int act)
{
    if (act == IA_ACT_SOFTPLUS100) return (z * BETA > 20.f) ? z : log1pf(expf(z * BETA)) / BETA;
    return fmaxf(z, 0.f);
}

// derivative expressed through the activation OUTPUT h (no pre-activations are kept)
__device__ __forceinline__ float act_bwd_from_out(float h, int act)
{
    if (act == IA_ACT_SOFTPLUS100) return -expm1f(-BETA * h);  // sigmoid(beta*z) = 1 - exp(-beta*h)
    return h > 0.f ? 1.f : 0.f;
}

// acc[8][4] += A(k, 8ty..8ty+7) * B[k][n0..n0+3]
__device__ __forceinline__ void gemm_8x4(const float *__restrict__ As, int K, const float *__restrict__ Bs, int ldb,
                                         int n0, int ty, float (&acc)[8][4])
{
    const int c0 = 2 * ty, c1 = 2 * ty + 1;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const int sw = (k >> 2) & 7;
        const float4 a0 = *reinterpret_cast<const float4 *>(As + k * MT + ((c0 ^ sw) << 2));
        const float4 a1 = *reinterpret_cast<const float4 *>(As + k * MT + ((c1 ^ sw) << 2));
        const float4 b = *reinterpret_cast<const float4 *>(Bs + k * ldb + n0);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// acc[4][4] += sum_m P(o0+a, m) * Q(i0+b, m);  bsum[a] += sum_m P(o0+a, m)
__device__ __forceinline__ void gemm_dw(const float *__restrict__ P, int o0, const float *__restrict__ Q, int i0,
                                        float (&acc)[4][4], float (&bsum)[4])
{
    const int swo = (o0 >> 2) & 7, swi = (i0 >> 2) & 7;
#pragma unroll 2
    for (int mc = 0; mc < MT / 4; ++mc) {
        float4 p[4], q[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) p[a] = *reinterpret_cast<const float4 *>(P + (o0 + a) * MT + ((mc ^ swo) << 2));
#pragma unroll
        for (int b = 0; b < 4; ++b) q[b] = *reinterpret_cast<const float4 *>(Q + (i0 + b) * MT + ((mc ^ swi) << 2));
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            bsum[a] += (p[a].x + p[a].y) + (p[a].z + p[a].w);
#pragma unroll
            for (int b = 0; b < 4; ++b)
                acc[a][b] += p[a].x * q[b].x + p[a].y * q[b].y + p[a].z * q[b].z + p[a].w * q[b].w;
        }
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4])
{
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// (bias + activation) epilogue: acc -> Hs (feature-major, swizzled), feature index n0..n0+3, rows 8ty..8ty+7
__device__ __forceinline__ void store_hidden(float *__restrict__ Hs, const float (&acc)[8][4], const float *__restrict__ bias,
                                             int n0, int ty, int act)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = n0 + j;
        const int sw = (k >> 2) & 7;
        const float b = bias[k];
        float4 lo, hi;
        lo.x = act_fwd(acc[0][j] + b, act); lo.y = act_fwd(acc[1][j] + b, act);
        lo.z = act_fwd(acc[2][j] + b, act); lo.w = act_fwd(acc[3][j] + b, act);
        hi.x = act_fwd(acc[4][j] + b, act); hi.y = act_fwd(acc[5][j] + b, act);
        hi.z = act_fwd(acc[6][j] + b, act); hi.w = act_fwd(acc[7][j] + b, act);
        *reinterpret_cast<float4 *>(Hs + k * MT + (((2 * ty) ^ sw) << 2)) = lo;
        *reinterpret_cast<float4 *>(Hs + k * MT + (((2 * ty + 1) ^ sw) << 2)) = hi;
    }
}

// dZ = dH (*) act'(H), written over H in place
__device__ __forceinline__ void store_dz_inplace(float *__restrict__ Hs, const float (&acc)[8][4], int n0, int ty, int act)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = n0 + j;
        const int sw = (k >> 2) & 7;
        float4 *plo = reinterpret_cast<float4 *>(Hs + k * MT + (((2 * ty) ^ sw) << 2));
        float4 *phi = reinterpret_cast<float4 *>(Hs + k * MT + (((2 * ty + 1) ^ sw) << 2));
        float4 lo = *plo, hi = *phi;
        lo.x = acc[0][j] * act_bwd_from_out(lo.x, act); lo.y = acc[1][j] * act_bwd_from_out(lo.y, act);
        lo.z = acc[2][j] * act_bwd_from_out(lo.z, act); lo.w = acc[3][j] * act_bwd_from_out(lo.w, act);
        hi.x = acc[4][j] * act_bwd_from_out(hi.x, act); hi.y = acc[5][j] * act_bwd_from_out(hi.y, act);
        hi.z = acc[6][j] * act_bwd_from_out(hi.z, act); hi.w = acc[7][j] * act_bwd_from_out(hi.w, act);
        *plo = lo;
        *phi = hi;
    }
}

// global [rows, din] (in0 affine prefix + in1) -> Xs feature-major; rows >= n_valid and features >= din are zero
__device__ __forceinline__ void load_input_tile(float *__restrict__ Xs, const MlpDims &D, const float *__restrict__ in0,
                                                const float *__restrict__ in1, int64_t row0, int n_valid)
{
    const int total = MT * D.K0p;
    for (int i = threadIdx.x; i < total; i += NT) {
        const int m = i / D.K0p, k = i - m * D.K0p;
        float v = 0.f;
        if (m < n_valid && k < D.din) {
            if (k < D.n_in0) v = fmaf(__ldg(in0 + (row0 + m) * D.n_in0 + k), D.s0, D.o0);
            else v = __ldg(in1 + (row0 + m) * D.n_in1 + (k - D.n_in0));
        }
        Xs[kidx(k, m)] = v;
    }
}

// W[out][in] (global, row-major) -> Bs[k=in][n=out] with row stride ldb, zero padded to (Kp x ldb)
__device__ __forceinline__ void load_wt(float *__restrict__ Bs, const float *__restrict__ Wg, int n_out, int n_in, int Kp,
                                        int ldb)
{
    for (int i = threadIdx.x; i < Kp * ldb; i += NT) Bs[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < n_out * n_in; i += NT) {
        const int n = i / n_in, k = i - n * n_in;
        Bs[k * ldb + n] = __ldg(Wg + i);
    }
}

// W[out][in] -> Bs[k=out][n=in] with row stride ldb (zero padded to rows_p x ldb)
__device__ __forceinline__ void load_w(float *__restrict__ Bs, const float *__restrict__ Wg, int n_out, int n_in, int rows_p,
                                       int ldb)
{
    for (int i = threadIdx.x; i < rows_p * ldb; i += NT) {
        const int o = i / ldb, c = i - o * ldb;
        Bs[i] = (o < n_out && c < n_in) ? __ldg(Wg + o * n_in + c) : 0.f;
    }
}

__device__ __forceinline__ int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
mlp_fwd_kernel(const MlpDims D, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
               const float *__restrict__ params, float *__restrict__ out, int64_t ld_out)
{
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const bool gemv = D.nou <= 4;
    const int n_oslab = gemv ? 0 : (D.nou + 63) / 64;
    float *Xs = smem;
    float *H1s = Xs + D.K0p * MT;
    float *H2s = H1s + W * MT;
    float *Bf0 = H2s + (D.nh == 2 ? W * MT : 0);
    float *Bf1 = Bf0 + D.K0p * W;
    float *Bfl = Bf1 + (D.nh == 2 ? W * W : 0);             // slab path: [64][64*n_oslab]; gemv path: [nou][64]
    float *b0 = Bfl + (gemv ? D.nou * W : W * 64 * n_oslab);
    float *b1 = b0 + W;
    float *bl = b1 + W;

    load_wt(Bf0, params + D.pW0, W, D.din, D.K0p, W);
    if (D.nh == 2) load_wt(Bf1, params + D.pW1, W, W, W, W);
    if (gemv) {
        for (int i = tid; i < D.nou * W; i += NT) Bfl[i] = __ldg(params + D.pWl + i);
    } else {
        load_wt(Bfl, params + D.pWl, D.nou, W, W, 64 * n_oslab);
    }
    for (int i = tid; i < W; i += NT) {
        b0[i] = __ldg(params + D.pb0 + i);
        b1[i] = D.nh == 2 ? __ldg(params + D.pb1 + i) : 0.f;
    }
    for (int i = tid; i < D.nou; i += NT) bl[i] = __ldg(params + D.pbl + i);
    __syncthreads();

    const int64_t n_tiles = (n + MT - 1) / MT;
    float acc[8][4];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * MT;
        const int n_valid = (int)min((int64_t)MT, n - row0);
        load_input_tile(Xs, D, in0, in1, row0, n_valid);
        __syncthreads();
        zero_acc(acc);
        gemm_8x4(Xs, D.din, Bf0, W, 4 * tx, ty, acc);
        store_hidden(H1s, acc, b0, 4 * tx, ty, D.act);
        __syncthreads();
        const float *HL = H1s;
        if (D.nh == 2) {
            zero_acc(acc);
            gemm_8x4(H1s, W, Bf1, W, 4 * tx, ty, acc);
            store_hidden(H2s, acc, b1, 4 * tx, ty, D.act);
            __syncthreads();
            HL = H2s;
        }
        if (gemv) {
            if (tid < MT && tid < n_valid) {
                for (int c = 0; c < D.nou; ++c) {
                    float s = bl[c];
#pragma unroll 8
                    for (int k = 0; k < W; ++k) s = fmaf(HL[kidx(k, tid)], Bfl[c * W + k], s);
                    out[(row0 + tid) * ld_out + c] = s;
                }
            }
        } else {
            for (int sl = 0; sl < n_oslab; ++sl) {
                zero_acc(acc);
                const int n0 = 64 * sl + 4 * tx;
                gemm_8x4(HL, W, Bfl, 64 * n_oslab, n0, ty, acc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = 8 * ty + i;
                    if (m < n_valid) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n0 + j < D.nou) out[(row0 + m) * ld_out + n0 + j] = acc[i][j] + bl[n0 + j];
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward (recompute + chain rule, in place)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 1)
mlp_bwd_kernel(const MlpDims D, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
               const float *__restrict__ params, const float *__restrict__ dout, int64_t ld_dout,
               float *__restrict__ din0, float *__restrict__ din1, float *__restrict__ dparams)
{
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const bool want_dx = (din0 != nullptr) || (din1 != nullptr);
    const int Gp = round_up(D.nou, 4);          // rows of the dY operand
    const int n_islab = (D.K0p + 63) / 64;      // column slabs of dX / dW0
    const int n_oslab = (Gp + 63) / 64;         // row slabs of dWl
    const int ldb0 = 64 * n_islab;
    float *Xs = smem;
    float *H1s = Xs + D.K0p * MT;
    float *H2s = H1s + W * MT;
    float *Gs = H2s + (D.nh == 2 ? W * MT : 0);
    float *Bf0 = Gs + Gp * MT;
    float *Bf1 = Bf0 + D.K0p * W;
    float *Bbl = Bf1 + (D.nh == 2 ? W * W : 0);   // [nou rows (padded to Gp)][64]
    float *Bb1 = Bbl + Gp * W;                    // [64][64]
    float *Bb0 = Bb1 + (D.nh == 2 ? W * W : 0);   // [64][ldb0]
    float *b0 = Bb0 + (want_dx ? W * ldb0 : 0);
    float *b1 = b0 + W;

    load_wt(Bf0, params + D.pW0, W, D.din, D.K0p, W);
    if (D.nh == 2) load_wt(Bf1, params + D.pW1, W, W, W, W);
    load_w(Bbl, params + D.pWl, D.nou, W, Gp, W);
    if (D.nh == 2) load_w(Bb1, params + D.pW1, W, W, W, W);
    if (want_dx) load_w(Bb0, params + D.pW0, W, D.din, W, ldb0);
    for (int i = tid; i < W; i += NT) {
        b0[i] = __ldg(params + D.pb0 + i);
        b1[i] = D.nh == 2 ? __ldg(params + D.pb1 + i) : 0.f;
    }
    __syncthreads();

    // persistent register accumulators for the parameter gradients (this thread's 4x4 blocks)
    float gW0[2][4][4], gW1[4][4], gWl[2][4][4];
    float gb0[4], gb1[4], gbl[2][4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        gb0[a] = gb1[a] = gbl[0][a] = gbl[1][a] = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) gW0[0][a][b] = gW0[1][a][b] = gW1[a][b] = gWl[0][a][b] = gWl[1][a][b] = 0.f;
    }

    const int64_t n_tiles = (n + MT - 1) / MT;
    float acc[8][4];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * MT;
        const int n_valid = (int)min((int64_t)MT, n - row0);
        load_input_tile(Xs, D, in0, in1, row0, n_valid);
        // dY tile, feature-major; rows >= nou and invalid points are zero
        for (int i = tid; i < MT * Gp; i += NT) {
            const int m = i / Gp, c = i - m * Gp;
            float v = 0.f;
            if (m < n_valid && c < D.nou) v = __ldg(dout + (row0 + m) * ld_dout + c);
            Gs[kidx(c, m)] = v;
        }
        __syncthreads();
        // ---- recompute hidden activations
        zero_acc(acc);
        gemm_8x4(Xs, D.din, Bf0, W, 4 * tx, ty, acc);
        store_hidden(H1s, acc, b0, 4 * tx, ty, D.act);
        __syncthreads();
        float *HL = H1s;
        if (D.nh == 2) {
            zero_acc(acc);
            gemm_8x4(H1s, W, Bf1, W, 4 * tx, ty, acc);
            store_hidden(H2s, acc, b1, 4 * tx, ty, D.act);
            __syncthreads();
            HL = H2s;
        }
        // ---- output layer: dWl += dY^T HL, dbl += sum dY, dHL = dY Wl
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int o0 = 64 * sl + 4 * ty;
            if (sl < n_oslab && o0 < Gp) gemm_dw(Gs, o0, HL, 4 * tx, gWl[sl], gbl[sl]);
        }
        zero_acc(acc);
        gemm_8x4(Gs, D.nou, Bbl, W, 4 * tx, ty, acc);
        __syncthreads();  // every dWl read of HL is done before it is overwritten
        store_dz_inplace(HL, acc, 4 * tx, ty, D.act);
        __syncthreads();
        if (D.nh == 2) {
            // ---- hidden layer 2: dW1 += dZ2^T H1, dH1 = dZ2 W1
            gemm_dw(H2s, 4 * ty, H1s, 4 * tx, gW1, gb1);
            zero_acc(acc);
            gemm_8x4(H2s, W, Bb1, W, 4 * tx, ty, acc);
            __syncthreads();
            store_dz_inplace(H1s, acc, 4 * tx, ty, D.act);
            __syncthreads();
        }
        // ---- first layer: dW0 += dZ1^T X, dX = dZ1 W0
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int i0 = 64 * sl + 4 * tx;
            if (sl < n_islab && i0 < D.K0p) {
                if (sl == 0) {
                    gemm_dw(H1s, 4 * ty, Xs, i0, gW0[0], gb0);
                } else {
                    float unused[4] = {0.f, 0.f, 0.f, 0.f};  // the bias sum belongs to slab 0 only
                    gemm_dw(H1s, 4 * ty, Xs, i0, gW0[1], unused);
                }
            }
        }
        if (want_dx) {
            for (int sl = 0; sl < n_islab; ++sl) {
                const int n0 = 64 * sl + 4 * tx;
                zero_acc(acc);
                gemm_8x4(H1s, W, Bb0, ldb0, n0, ty, acc);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = 8 * ty + i;
                    if (m >= n_valid) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = n0 + j;
                        if (c >= D.din) continue;
                        if (c < D.n_in0) {
                            if (din0) din0[(row0 + m) * D.n_in0 + c] = acc[i][j] * D.s0;
                        } else if (din1) {
                            din1[(row0 + m) * D.n_in1 + (c - D.n_in0)] = acc[i][j];
                        }
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- flush parameter gradients
    if (dparams != nullptr) {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int i0 = 64 * sl + 4 * tx;
            if (sl < n_islab && i0 < D.K0p) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (i0 + b < D.din) atomicAdd(dparams + D.pW0 + (4 * ty + a) * D.din + i0 + b, gW0[sl][a][b]);
            }
        }
        if (tx == 0) {
#pragma unroll
            for (int a = 0; a < 4; ++a) atomicAdd(dparams + D.pb0 + 4 * ty + a, gb0[a]);
        }
        if (D.nh == 2) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) atomicAdd(dparams + D.pW1 + (4 * ty + a) * W + 4 * tx + b, gW1[a][b]);
            if (tx == 0) {
#pragma unroll
                for (int a = 0; a < 4; ++a) atomicAdd(dparams + D.pb1 + 4 * ty + a, gb1[a]);
            }
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const int o0 = 64 * sl + 4 * ty;
            if (sl < n_oslab && o0 < Gp) {
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    if (o0 + a >= D.nou) continue;
#pragma unroll
                    for (int b = 0; b < 4; ++b) atomicAdd(dparams + D.pWl + (o0 + a) * W + 4 * tx + b, gWl[sl][a][b]);
                    if (tx == 0) atomicAdd(dparams + D.pbl + o0 + a, gbl[sl][a]);
                }
            }
        }
    }
}

int make_dims(const ia_mlp_desc *d, int32_t n_out_used, MlpDims *D)
{
    IA_REQUIRE(d != nullptr, "mlp: desc is NULL");
    IA_REQUIRE(d->width == W, "mlp: width must be 64 (got %d)", d->width);
    IA_REQUIRE(d->n_hidden_layers == 1 || d->n_hidden_layers == 2, "mlp: n_hidden_layers must be 1 or 2 (got %d)", d->n_hidden_layers);
    IA_REQUIRE(d->n_in0 >= 0 && d->n_in0 <= 8 && d->n_in1 >= 0, "mlp: bad input split");
    const int din = d->n_in0 + d->n_in1;
    IA_REQUIRE(din >= 1 && din <= 128, "mlp: input width %d not in [1,128]", din);
    IA_REQUIRE(d->n_out >= 1 && d->n_out <= 128, "mlp: n_out %d not in [1,128]", d->n_out);
    IA_REQUIRE(n_out_used >= 1 && n_out_used <= d->n_out, "mlp: n_out_used %d not in [1,%d]", n_out_used, d->n_out);
    IA_REQUIRE(d->hidden_act == IA_ACT_RELU || d->hidden_act == IA_ACT_SOFTPLUS100, "mlp: unsupported hidden activation %d", d->hidden_act);
    if (d->out_act != IA_ACT_NONE) {
        ia_set_error("mlp: fused output activation %d not supported (apply it on the caller side)", d->out_act);
        return IA_ERR_UNSUPPORTED;
    }
    D->n_in0 = d->n_in0; D->n_in1 = d->n_in1; D->din = din; D->K0p = (din + 3) / 4 * 4;
    D->s0 = d->in0_scale; D->o0 = d->in0_offset;
    D->nh = d->n_hidden_layers; D->n_out = d->n_out; D->nou = n_out_used; D->act = d->hidden_act;
    int p = 0;
    D->pW0 = p; p += W * din;
    D->pb0 = p; p += W;
    D->pW1 = p; D->pb1 = p;
    if (D->nh == 2) { D->pW1 = p; p += W * W; D->pb1 = p; p += W; }
    D->pWl = p; p += d->n_out * W;
    D->pbl = p;
    return IA_OK;
}

}  // namespace

extern "C" int64_t ia_mlp_param_count(const ia_mlp_desc *d)
{
    if (!d) return -1;
    const int64_t din = d->n_in0 + d->n_in1;
    int64_t p = (int64_t)d->width * din + d->width;
    if (d->n_hidden_layers == 2) p += (int64_t)d->width * d->width + d->width;
    p += (int64_t)d->n_out * d->width + d->n_out;
    return p;
}

int ia_mlp_fwd_fp32(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                    int32_t n_out_used, float *out, int64_t ld_out, void *stream)
{
    MlpDims D;
    int rc = make_dims(desc, n_out_used, &D);
    if (rc) return rc;
    IA_REQUIRE(n >= 0 && (n == 0 || (params && out)), "mlp_fwd: NULL pointer");
    IA_REQUIRE(n == 0 || ((D.n_in0 == 0 || in0) && (D.n_in1 == 0 || in1)), "mlp_fwd: missing input pointer");
    IA_REQUIRE(ld_out >= n_out_used, "mlp_fwd: ld_out < n_out_used");
    if (n == 0) return IA_OK;
    const bool gemv = D.nou <= 4;
    const int n_oslab = gemv ? 0 : (D.nou + 63) / 64;
    size_t fl = (size_t)D.K0p * MT + W * MT + (D.nh == 2 ? W * MT : 0) + (size_t)D.K0p * W + (D.nh == 2 ? W * W : 0) +
                (gemv ? D.nou * W : W * 64 * n_oslab) + 2 * W + D.nou;
    const size_t bytes = fl * sizeof(float);
    IA_REQUIRE(bytes <= 227 * 1024, "mlp_fwd: configuration needs %zu B of shared memory", bytes);
    IA_CUDA_OK(cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int64_t n_tiles = ia_ceil_div(n, MT);
    const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ia_sm_count());
    mlp_fwd_kernel<<<blocks, NT, bytes, (cudaStream_t)stream>>>(D, in0, in1, n, params, out, ld_out);
    IA_LAUNCH_OK("mlp_fwd_kernel");
    return IA_OK;
}

int ia_mlp_bwd_fp32(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                    const float *dout, int32_t n_out_used, int64_t ld_dout, float *din0, float *din1, float *dparams,
                    void *stream)
{
    MlpDims D;
    int rc = make_dims(desc, n_out_used, &D);
    if (rc) return rc;
    IA_REQUIRE(n >= 0 && (n == 0 || (params && dout)), "mlp_bwd: NULL pointer");
    IA_REQUIRE(n == 0 || ((D.n_in0 == 0 || in0) && (D.n_in1 == 0 || in1)), "mlp_bwd: missing input pointer");
    IA_REQUIRE(ld_dout >= n_out_used, "mlp_bwd: ld_dout < n_out_used");
    if (n == 0) return IA_OK;
    const bool want_dx = din0 || din1;
    const int Gp = (D.nou + 3) / 4 * 4;
    const int n_islab = (D.K0p + 63) / 64;
    size_t fl = (size_t)D.K0p * MT + W * MT + (D.nh == 2 ? W * MT : 0) + (size_t)Gp * MT + (size_t)D.K0p * W +
                (D.nh == 2 ? W * W : 0) + (size_t)Gp * W + (D.nh == 2 ? W * W : 0) + (want_dx ? W * 64 * n_islab : 0) + 2 * W;
    const size_t bytes = fl * sizeof(float);
    IA_REQUIRE(bytes <= 227 * 1024, "mlp_bwd: configuration needs %zu B of shared memory", bytes);
    IA_CUDA_OK(cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const int64_t n_tiles = ia_ceil_div(n, MT);
    const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ia_sm_count());
    mlp_bwd_kernel<<<blocks, NT, bytes, (cudaStream_t)stream>>>(D, in0, in1, n, params, dout, ld_dout, din0, din1, dparams);
    IA_LAUNCH_OK("mlp_bwd_kernel");
    return IA_OK;
}
