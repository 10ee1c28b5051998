// Wide output layer of the width-64 networks: out[n, n_out] = h[n, 64] * W[n_out, 64]^T + b   (n_out <= 72 on the
// tiled kernels, <= 128 on the generic ones), forward and backward, fp32 FFMA.  Used for the 65-feature centre evaluation
// of the SDF network (reference models/geometry.py:206: the full `feature` output), whose hidden layers run on the
// tensor-core kernel (mlp_tc.cu, feature mode).  Skinny-N / skinny-K GEMMs like these are where a general GEMM library
// does worst (cuBLAS picks a large-K SIMT kernel for dW = dY^T h: 1.5 ms at 1.5 M rows).
//
// Tiled kernels (n_out <= 72): one CTA owns 128 rows at a time; every thread keeps a 4 x 9 (forward), 4 x 8 (input
// gradient) or 18 x 2 (weight gradient) block of accumulators in registers so that the inner loop is FFMA-bound (36 / 32
// / 36 FMAs against 7 / 6 / 7 shared-memory wavefronts); all global traffic is staged through shared memory so that rows
// with arbitrary leading dimensions (the colour head's 87-wide input row) are still read and written in full lines.
#include <algorithm>

#include "ia_common.cuh"

namespace {

constexpr int LW = 64;          // input width
constexpr int LROWS = 128;      // rows per tile
constexpr int LTHREADS = 256;
constexpr int LPAD = 65;        // padded row length of the activation tiles (conflict-free row-per-lane access)
constexpr int WPAD = 68;        // padded row length of the weight tile (16-byte aligned rows for float4 broadcasts)
constexpr int MAX_NOUT = 128;
constexpr int TILED_NOUT = 72;  // 8 column groups x 9 outputs
constexpr int WTS = 96;         // forward: W^T row = 8 groups x 12 floats (9 used), 16-byte aligned groups
constexpr int OLD = 97;         // output / dY staging tile leading dimension (odd: conflict-free row-per-lane access); up to 96 columns

// Optional tail of the colour head's input row (reference models/geometry.py:207 + models/texture.py:26-27): columns
// [n_out, n_out + 3 + n_enc + 3) of `out` = (pts01*2-1 | enc | normal); column 0 / columns 1..3 are also copied out.
struct HeadTail {
    const float *pts01, *enc, *normal;   // [n,3], [n,n_enc], [n,3]; pts01 == nullptr: no tail
    int n_enc;
    float *sdf, *rgb_raw;                // [n], [n,3] (rgb_raw may be nullptr)
};

// -------------------------------------------------------------------------------------------------------------------
// tiled forward: thread (tr = t & 31, tc = t >> 5) -> rows {tr, tr+32, tr+64, tr+96} x outputs {9 tc .. 9 tc + 8}
__global__ void __launch_bounds__(LTHREADS, 2)
linear64_fwd_tiled_kernel(const float *__restrict__ h, int64_t n, const float *__restrict__ Wg, const float *__restrict__ bg,
                          int n_out, float *__restrict__ out, int64_t ld_out, const HeadTail T)
{
    extern __shared__ __align__(16) float sm[];
    float *Wt = sm;                        // [64][WTS]: Wt[k][12 g + j] = W[9 g + j][k]
    float *bs = Wt + LW * WTS;             // [72]
    float *Hs = bs + TILED_NOUT;           // [LROWS][LPAD]; reused as the output tile [LROWS][OLD]
    const int tid = threadIdx.x;
    for (int i = tid; i < LW * WTS; i += LTHREADS) Wt[i] = 0.f;
    for (int i = tid; i < TILED_NOUT; i += LTHREADS) bs[i] = i < n_out ? __ldg(bg + i) : 0.f;
    __syncthreads();
    for (int i = tid; i < n_out * LW; i += LTHREADS) {
        const int o = i / LW, k = i - o * LW;
        Wt[k * WTS + 12 * (o / 9) + (o % 9)] = __ldg(Wg + i);
    }
    const int tr = tid & 31, tc = tid >> 5;
    const int64_t n_tiles = (n + LROWS - 1) / LROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * LROWS;
        __syncthreads();
        {
            float4 v[LROWS * (LW / 4) / LTHREADS];
#pragma unroll
            for (int u = 0; u < LROWS * (LW / 4) / LTHREADS; ++u) {
                const int i = tid + u * LTHREADS, rr = i >> 4, q = i & 15;
                v[u] = (row0 + rr < n) ? __ldg(reinterpret_cast<const float4 *>(h + (row0 + rr) * LW) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < LROWS * (LW / 4) / LTHREADS; ++u) {
                const int i = tid + u * LTHREADS, rr = i >> 4, q = i & 15;
                float *d = Hs + rr * LPAD + 4 * q;
                d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
            }
        }
        // tail columns of this tile: one element per lane for each of the warp's 16 rows, fetched before the FMA loop
        const int n_tail = T.pts01 != nullptr ? 6 + T.n_enc : 0;
        float tail[LROWS / (LTHREADS / 32)];
        if (n_tail > 0) {
            const int c = tid & 31;
#pragma unroll
            for (int u = 0; u < LROWS / (LTHREADS / 32); ++u) {
                const int64_t r = row0 + (tid >> 5) + u * (LTHREADS / 32);
                // Raw loads only, branch-free source select: the first use of a loaded value stalls the (in-order) warp, so
                // any arithmetic on tail[u] here would serialise the 16 rows on DRAM latency (measured: +0.45 ms per 1.5 M
                // rows).  The pts01*2-1 affine is applied when the values are written to the output tile.
                const float *src = c < 3 ? T.pts01 + 3 * r + c
                                         : (c < 3 + T.n_enc ? T.enc + r * T.n_enc + (c - 3) : T.normal + 3 * r + (c - 3 - T.n_enc));
                tail[u] = __ldg((r < n && c < n_tail) ? src : T.pts01);
            }
        }
        __syncthreads();
        float acc[4][9];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) acc[i][j] = bs[9 * tc + j];
#pragma unroll 4
        for (int k = 0; k < LW; ++k) {
            float hv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) hv[i] = Hs[(tr + 32 * i) * LPAD + k];
            const float4 *w4 = reinterpret_cast<const float4 *>(Wt + k * WTS + 12 * tc);
            const float4 wa = w4[0], wb = w4[1], wc = w4[2];
            const float w[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 9; ++j) acc[i][j] = fmaf(hv[i], w[j], acc[i][j]);
        }
        __syncthreads();                   // every thread is done reading Hs: reuse it as the output tile
        float *Os = Hs;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j)
                if (9 * tc + j < n_out) Os[(tr + 32 * i) * OLD + 9 * tc + j] = acc[i][j];
        if (n_tail > 0 && (tid & 31) < n_tail) {
#pragma unroll
            for (int u = 0; u < LROWS / (LTHREADS / 32); ++u)
                Os[((tid >> 5) + u * (LTHREADS / 32)) * OLD + n_out + (tid & 31)] = (tid & 31) < 3 ? tail[u] * 2.0f - 1.0f : tail[u];
        }
        __syncthreads();
        const int n_cols = n_out + n_tail;
        for (int rr = tid >> 5; rr < LROWS; rr += LTHREADS / 32) {
            const int64_t r = row0 + rr;
            if (r < n) {
                float *dst = out + r * ld_out;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int o = (tid & 31) + 32 * c;
                    if (o < n_cols) dst[o] = Os[rr * OLD + o];
                }
                if (n_tail > 0 && (tid & 31) < 4) {
                    const float v = Os[rr * OLD + (tid & 31)];
                    if ((tid & 31) == 0) T.sdf[r] = v;
                    else if (T.rgb_raw != nullptr) T.rgb_raw[3 * r + (tid & 31) - 1] = v;
                }
            }
        }
    }
}

// Stage a [LROWS][<= 96] tile of dY (+ the optional extra gradient of its first n_ex columns) into shared memory.
// One warp per row, three 32-lane column chunks; the row loop is unrolled so that 12 independent loads are in flight per
// thread (a loop with one load -> one shared store per trip serialises the whole tile on DRAM latency).
// Column o lands at Ds[rr * ld + 20 * (o / 18) + o % 18] when GROUPED (weight-gradient layout), else at Ds[rr * ld + o];
// columns in [n_out, n_cols) are written as zeros.
template <int NTHREADS, bool GROUPED>
__device__ __forceinline__ void stage_dy(float *Ds, int ld, const float *__restrict__ dy, int64_t ld_dy,
                                         const float *__restrict__ dex, int n_ex, int64_t row0, int64_t n, int n_out, int n_cols,
                                         int tid)
{
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = NTHREADS / 32;
#pragma unroll 4
    for (int rr = warp; rr < LROWS; rr += NW) {
        const bool live = row0 + rr < n;
        const float *src = dy + (row0 + rr) * ld_dy;
        float v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int o = lane + 32 * c;
            v[c] = (live && o < n_out) ? __ldg(src + o) : 0.f;
        }
        if (live && lane < n_ex) v[0] += __ldg(dex + (row0 + rr) * n_ex + lane);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int o = lane + 32 * c;
            if (o < n_cols) Ds[rr * ld + (GROUPED ? 20 * (o / 18) + (o % 18) : o)] = v[c];
        }
    }
}

// tiled input gradient: thread (tr, tc) -> rows {tr + 32 i} x k in [8 tc, 8 tc + 8)
__global__ void __launch_bounds__(LTHREADS, 2)
linear64_bwd_input_tiled_kernel(const float *__restrict__ dy, int64_t ld_dy, const float *__restrict__ dex, int n_ex, int64_t n,
                                const float *__restrict__ Wg, int n_out, float *__restrict__ dh, int n_tail_enc,
                                float *__restrict__ dpts01, float *__restrict__ denc, float *__restrict__ dnormal)
{
    extern __shared__ __align__(16) float sm[];
    float *Ws = sm;                        // [n_out][WPAD]
    float *Ds = Ws + TILED_NOUT * WPAD;    // [LROWS][OLD]; reused as the dh tile [LROWS][LPAD]
    const int tid = threadIdx.x;
    for (int i = tid; i < n_out * LW; i += LTHREADS) Ws[(i / LW) * WPAD + (i % LW)] = __ldg(Wg + i);
    const int tr = tid & 31, tc = tid >> 5;
    const int64_t n_tiles = (n + LROWS - 1) / LROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * LROWS;
        __syncthreads();
        const int n_tail = n_tail_enc >= 0 ? 6 + n_tail_enc : 0;     // tail columns [n_out, n_out + n_tail) of the same rows
        stage_dy<LTHREADS, false>(Ds, OLD, dy, ld_dy, dex, n_ex, row0, n, n_out + n_tail, n_out + n_tail, tid);
        __syncthreads();
        if (n_tail > 0) {
            const int c = tid & 31;
            for (int rr = tid >> 5; rr < LROWS; rr += LTHREADS / 32) {
                const int64_t r = row0 + rr;
                if (r < n && c < n_tail) {
                    const float g = Ds[rr * OLD + n_out + c];
                    if (c < 3) {
                        if (dpts01 != nullptr) dpts01[3 * r + c] = 2.0f * g;
                    } else if (c < 3 + n_tail_enc) {
                        if (denc != nullptr) denc[r * n_tail_enc + (c - 3)] = g;
                    } else if (dnormal != nullptr) {
                        dnormal[3 * r + (c - 3 - n_tail_enc)] = g;
                    }
                }
            }
        }
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int o = 0; o < n_out; ++o) {
            float g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] = Ds[(tr + 32 * i) * OLD + o];
            const float4 *w4 = reinterpret_cast<const float4 *>(Ws + o * WPAD + 8 * tc);
            const float4 wa = w4[0], wb = w4[1];
            const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(g[i], w[j], acc[i][j]);
        }
        __syncthreads();
        float *Hs = Ds;                    // [LROWS][LPAD]
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) Hs[(tr + 32 * i) * LPAD + 8 * tc + j] = acc[i][j];
        __syncthreads();
        for (int i = tid; i < LROWS * (LW / 4); i += LTHREADS) {
            const int rr = i >> 4, q = i & 15;
            if (row0 + rr < n) {
                const float *s = Hs + rr * LPAD + 4 * q;
                reinterpret_cast<float4 *>(dh + (row0 + rr) * LW)[q] = make_float4(s[0], s[1], s[2], s[3]);
            }
        }
    }
}

// tiled weight gradient: 128 threads; thread (tk = t & 31, to = t >> 5) -> outputs {18 to .. 18 to + 17} x k in {2 tk, 2 tk + 1};
// accumulators persist over all tiles of the CTA, one atomic flush at the end.
constexpr int WG_THREADS = 128;
constexpr int DLD = 80;                    // dY tile row: 4 groups of 20 floats (18 used), 16-byte aligned groups

__global__ void __launch_bounds__(WG_THREADS, 3)
linear64_bwd_weight_tiled_kernel(const float *__restrict__ h, const float *__restrict__ dy, int64_t ld_dy,
                                 const float *__restrict__ dex, int n_ex, int64_t n, int n_out, float *__restrict__ dW,
                                 float *__restrict__ db)
{
    extern __shared__ __align__(16) float sm[];
    float *Hs = sm;                        // [LROWS][LW]
    float *Ds = Hs + LROWS * LW;           // [LROWS][DLD], columns >= n_out are zero
    const int tid = threadIdx.x, tk = tid & 31, to = tid >> 5;
    float acc[18][2], bsum[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) acc[j][0] = acc[j][1] = bsum[j] = 0.f;
    const int64_t n_tiles = (n + LROWS - 1) / LROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * LROWS;
        __syncthreads();
#pragma unroll
        for (int u0 = 0; u0 < LROWS * (LW / 4) / WG_THREADS; u0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = tid + (u0 + u) * WG_THREADS, rr = i >> 4, q = i & 15;
                v[u] = (row0 + rr < n) ? __ldg(reinterpret_cast<const float4 *>(h + (row0 + rr) * LW) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = tid + (u0 + u) * WG_THREADS, rr = i >> 4, q = i & 15;
                reinterpret_cast<float4 *>(Hs + rr * LW)[q] = v[u];
            }
        }
        stage_dy<WG_THREADS, true>(Ds, DLD, dy, ld_dy, dex, n_ex, row0, n, n_out, TILED_NOUT, tid);
        __syncthreads();
#pragma unroll 2
        for (int rr = 0; rr < LROWS; ++rr) {
            const float2 hv = *reinterpret_cast<const float2 *>(Hs + rr * LW + 2 * tk);
            const float4 *d4 = reinterpret_cast<const float4 *>(Ds + rr * DLD + 20 * to);   // warp-uniform: broadcast loads
            const float4 ga = d4[0], gb = d4[1], gc = d4[2], gd = d4[3];
            const float2 ge = *reinterpret_cast<const float2 *>(d4 + 4);
            const float g[18] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w, gc.x, gc.y, gc.z, gc.w,
                                 gd.x, gd.y, gd.z, gd.w, ge.x, ge.y};
#pragma unroll
            for (int j = 0; j < 18; ++j) {
                acc[j][0] = fmaf(g[j], hv.x, acc[j][0]);
                acc[j][1] = fmaf(g[j], hv.y, acc[j][1]);
            }
            if (tk == 0) {
#pragma unroll
                for (int j = 0; j < 18; ++j) bsum[j] += g[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 18; ++j) {
        const int o = 18 * to + j;
        if (o < n_out) {
            atomicAdd(dW + o * LW + 2 * tk, acc[j][0]);
            atomicAdd(dW + o * LW + 2 * tk + 1, acc[j][1]);
            if (tk == 0 && db != nullptr) atomicAdd(db + o, bsum[j]);
        }
    }
}

// -------------------------------------------------------------------------------------------------------------------
// generic kernels (72 < n_out <= 128)
// out[r][o] = b[o] + sum_k h[r][k] W[o][k].  Thread (r = t & 127, half = t >> 7) produces outputs o = half, half+2, ...
__global__ void __launch_bounds__(LTHREADS)
linear64_fwd_kernel(const float *__restrict__ h, int64_t n, const float *__restrict__ Wg, const float *__restrict__ bg, int n_out,
                    float *__restrict__ out, int64_t ld_out)
{
    extern __shared__ __align__(16) float sm[];
    float *Ws = sm;                       // [n_out][WPAD]
    float *bs = Ws + n_out * WPAD;        // [n_out]
    float *Hs = bs + ((n_out + 3) & ~3);  // [LROWS][LPAD]
    const int tid = threadIdx.x;
    for (int i = tid; i < n_out * LW; i += LTHREADS) Ws[(i / LW) * WPAD + (i % LW)] = __ldg(Wg + i);
    for (int i = tid; i < n_out; i += LTHREADS) bs[i] = __ldg(bg + i);
    const int r = tid & (LROWS - 1), half = tid >> 7;
    const int64_t n_tiles = (n + LROWS - 1) / LROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * LROWS;
        __syncthreads();
        for (int i = tid; i < LROWS * LW; i += LTHREADS) {
            const int rr = i / LW, k = i - rr * LW;
            Hs[rr * LPAD + k] = (row0 + rr < n) ? __ldg(h + (row0 + rr) * LW + k) : 0.f;
        }
        __syncthreads();
        float hv[LW];
#pragma unroll
        for (int k = 0; k < LW; ++k) hv[k] = Hs[r * LPAD + k];
        if (row0 + r < n) {
            for (int o = half; o < n_out; o += 2) {
                float acc = bs[o];
                const float4 *w4 = reinterpret_cast<const float4 *>(Ws + o * WPAD);
#pragma unroll
                for (int q = 0; q < LW / 4; ++q) {
                    const float4 w = w4[q];
                    acc = fmaf(hv[4 * q], w.x, acc);
                    acc = fmaf(hv[4 * q + 1], w.y, acc);
                    acc = fmaf(hv[4 * q + 2], w.z, acc);
                    acc = fmaf(hv[4 * q + 3], w.w, acc);
                }
                out[(row0 + r) * ld_out + o] = acc;
            }
        }
    }
}

// dh[r][k] = sum_o dy[r][o] W[o][k]   (thread (r, half) produces k in [32*half, 32*half+32))
__global__ void __launch_bounds__(LTHREADS)
linear64_bwd_input_kernel(const float *__restrict__ dy, int64_t ld_dy, const float *__restrict__ dex, int n_ex, int64_t n,
                          const float *__restrict__ Wg, int n_out, float *__restrict__ dh)
{
    extern __shared__ __align__(16) float sm[];
    float *Ws = sm;                         // [n_out][WPAD]
    float *Ds = Ws + n_out * WPAD;          // [LROWS][n_out + 1]
    const int ldd = n_out + 1;
    const int tid = threadIdx.x;
    for (int i = tid; i < n_out * LW; i += LTHREADS) Ws[(i / LW) * WPAD + (i % LW)] = __ldg(Wg + i);
    const int r = tid & (LROWS - 1), half = tid >> 7;
    const int64_t n_tiles = (n + LROWS - 1) / LROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * LROWS;
        __syncthreads();
        for (int i = tid; i < LROWS * n_out; i += LTHREADS) {
            const int rr = i / n_out, o = i - rr * n_out;
            float g = 0.f;
            if (row0 + rr < n) {
                g = __ldg(dy + (row0 + rr) * ld_dy + o);
                if (o < n_ex) g += __ldg(dex + (row0 + rr) * n_ex + o);
            }
            Ds[rr * ldd + o] = g;
        }
        __syncthreads();
        float acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.f;
        for (int o = 0; o < n_out; ++o) {
            const float g = Ds[r * ldd + o];
            const float4 *w4 = reinterpret_cast<const float4 *>(Ws + o * WPAD + 32 * half);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 w = w4[q];
                acc[4 * q] = fmaf(g, w.x, acc[4 * q]);
                acc[4 * q + 1] = fmaf(g, w.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(g, w.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(g, w.w, acc[4 * q + 3]);
            }
        }
        if (row0 + r < n) {
            float4 *dst = reinterpret_cast<float4 *>(dh + (row0 + r) * LW + 32 * half);
#pragma unroll
            for (int q = 0; q < 8; ++q) dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
    }
}

// dW[o][k] += sum_r dy[r][o] h[r][k];  db[o] += sum_r dy[r][o].  Thread (ty = t >> 4, tx = t & 15) owns outputs
// o in {ty, ty+16, ...} x k in [4tx, 4tx+4); accumulators persist over all tiles of the CTA, one atomic flush at the end.
template <int NOBLK>     // ceil(n_out / 16): compile-time so that only the live output blocks cost instructions
__global__ void __launch_bounds__(LTHREADS)
linear64_bwd_weight_kernel(const float *__restrict__ h, const float *__restrict__ dy, int64_t ld_dy, const float *__restrict__ dex,
                           int n_ex, int64_t n, int n_out, float *__restrict__ dW, float *__restrict__ db)
{
    extern __shared__ __align__(16) float sm[];
    float *Hs = sm;                         // [LROWS][WPAD]  (16-byte aligned rows)
    float *Ds = Hs + LROWS * WPAD;          // [LROWS][16*NOBLK + 1], columns >= n_out are zero
    const int ldd = 16 * NOBLK + 1;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[NOBLK][4], bsum[NOBLK];
#pragma unroll
    for (int a = 0; a < NOBLK; ++a) {
        bsum[a] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
    }
    const int64_t n_tiles = (n + LROWS - 1) / LROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * LROWS;
        __syncthreads();
        for (int i = tid; i < LROWS * LW; i += LTHREADS) {
            const int rr = i / LW, k = i - rr * LW;
            Hs[rr * WPAD + k] = (row0 + rr < n) ? __ldg(h + (row0 + rr) * LW + k) : 0.f;
        }
        for (int i = tid; i < LROWS * 16 * NOBLK; i += LTHREADS) {
            const int rr = i / (16 * NOBLK), o = i - rr * (16 * NOBLK);
            float g = 0.f;
            if (row0 + rr < n && o < n_out) {
                g = __ldg(dy + (row0 + rr) * ld_dy + o);
                if (o < n_ex) g += __ldg(dex + (row0 + rr) * n_ex + o);
            }
            Ds[rr * ldd + o] = g;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < LROWS; ++rr) {
            const float4 hv = *reinterpret_cast<const float4 *>(Hs + rr * WPAD + 4 * tx);
#pragma unroll
            for (int a = 0; a < NOBLK; ++a) {
                const float g = Ds[rr * ldd + ty + 16 * a];
                acc[a][0] = fmaf(g, hv.x, acc[a][0]);
                acc[a][1] = fmaf(g, hv.y, acc[a][1]);
                acc[a][2] = fmaf(g, hv.z, acc[a][2]);
                acc[a][3] = fmaf(g, hv.w, acc[a][3]);
                if (tx == 0) bsum[a] += g;
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NOBLK; ++a) {
        const int o = ty + 16 * a;
        if (o < n_out) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(dW + o * LW + 4 * tx + j, acc[a][j]);
            if (tx == 0 && db != nullptr) atomicAdd(db + o, bsum[a]);
        }
    }
}

}  // namespace

static int linear64_fwd_impl(const float *h, int64_t n, const float *W, const float *b, int32_t n_out, float *out, int64_t ld_out,
                             const HeadTail &T, void *stream);

extern "C" int32_t ia_linear64_fwd(const float *h, int64_t n, const float *W, const float *b, int32_t n_out, float *out,
                                   int64_t ld_out, void *stream)
{
    HeadTail T{};
    return linear64_fwd_impl(h, n, W, b, n_out, out, ld_out, T, stream);
}

extern "C" int32_t ia_sdf_head_fwd(const float *h, int64_t n, const float *W, const float *b, int32_t n_feat, const float *pts01,
                                   const float *enc, int32_t n_enc, const float *normal, float *tin, int64_t ld_tin, float *sdf,
                                   float *rgb_raw, void *stream)
{
    IA_REQUIRE(n_feat >= 4 && n_feat <= TILED_NOUT, "sdf_head_fwd: n_feat %d not in [4,%d]", n_feat, TILED_NOUT);
    IA_REQUIRE(n_enc >= 0 && 6 + n_enc <= 32 && n_feat + 6 + n_enc <= 96, "sdf_head_fwd: n_enc %d too wide", n_enc);
    IA_REQUIRE(n == 0 || (pts01 && normal && sdf && (n_enc == 0 || enc)), "sdf_head_fwd: NULL pointer");
    IA_REQUIRE(ld_tin >= n_feat + 6 + n_enc, "sdf_head_fwd: ld_tin too small");
    HeadTail T{pts01, enc, normal, n_enc, sdf, rgb_raw};
    return linear64_fwd_impl(h, n, W, b, n_feat, tin, ld_tin, T, stream);
}

static int linear64_fwd_impl(const float *h, int64_t n, const float *W, const float *b, int32_t n_out, float *out, int64_t ld_out,
                             const HeadTail &T, void *stream)
{
    IA_REQUIRE(n_out >= 1 && n_out <= MAX_NOUT, "linear64_fwd: n_out %d not in [1,%d]", n_out, MAX_NOUT);
    IA_REQUIRE(n >= 0 && (n == 0 || (h && W && b && out)), "linear64_fwd: NULL pointer");
    IA_REQUIRE(ld_out >= n_out, "linear64_fwd: ld_out < n_out");
    IA_REQUIRE(((uintptr_t)h & 15) == 0, "linear64_fwd: h must be 16-byte aligned");
    if (n == 0) return IA_OK;
    const unsigned blocks = (unsigned)std::min<int64_t>(ia_ceil_div(n, LROWS), (int64_t)ia_sm_count() * 2);
    if (n_out <= TILED_NOUT) {
        const size_t bytes = sizeof(float) * ((size_t)LW * WTS + TILED_NOUT + (size_t)LROWS * OLD);
        IA_CUDA_OK(cudaFuncSetAttribute(linear64_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        linear64_fwd_tiled_kernel<<<blocks, LTHREADS, bytes, (cudaStream_t)stream>>>(h, n, W, b, n_out, out, ld_out, T);
        IA_LAUNCH_OK("linear64_fwd_tiled_kernel");
        return IA_OK;
    }
    const size_t bytes = sizeof(float) * ((size_t)n_out * WPAD + ((n_out + 3) & ~3) + (size_t)LROWS * LPAD);
    IA_CUDA_OK(cudaFuncSetAttribute(linear64_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    linear64_fwd_kernel<<<blocks, LTHREADS, bytes, (cudaStream_t)stream>>>(h, n, W, b, n_out, out, ld_out);
    IA_LAUNCH_OK("linear64_fwd_kernel");
    return IA_OK;
}

static int linear64_bwd_impl(const float *h, int64_t n, const float *W, const float *dout, int64_t ld_dout, int32_t n_out,
                             const float *dextra, int32_t n_extra, float *dh, float *dW, float *db, int n_tail_enc, float *dpts01,
                             float *denc, float *dnormal, void *stream);

extern "C" int32_t ia_linear64_bwd(const float *h, int64_t n, const float *W, const float *dout, int64_t ld_dout, int32_t n_out,
                                   const float *dextra, int32_t n_extra, float *dh, float *dW, float *db, void *stream)
{
    return linear64_bwd_impl(h, n, W, dout, ld_dout, n_out, dextra, n_extra, dh, dW, db, -1, nullptr, nullptr, nullptr, stream);
}

extern "C" int32_t ia_sdf_head_bwd(const float *h, int64_t n, const float *W, const float *dtin, int64_t ld_tin, int32_t n_feat,
                                   int32_t n_enc, const float *dextra, int32_t n_extra, float *dh, float *dW, float *db,
                                   float *dpts01, float *denc, float *dnormal, void *stream)
{
    IA_REQUIRE(n_feat >= 4 && n_feat <= TILED_NOUT, "sdf_head_bwd: n_feat %d not in [4,%d]", n_feat, TILED_NOUT);
    IA_REQUIRE(n_enc >= 0 && 6 + n_enc <= 32 && n_feat + 6 + n_enc <= 96, "sdf_head_bwd: n_enc %d too wide", n_enc);
    IA_REQUIRE(ld_tin >= n_feat + 6 + n_enc, "sdf_head_bwd: ld_tin too small");
    if (n == 0) return IA_OK;
    IA_REQUIRE(dh != nullptr, "sdf_head_bwd: dh is NULL (the tail gradients are produced by the input-gradient kernel)");
    return linear64_bwd_impl(h, n, W, dtin, ld_tin, n_feat, dextra, n_extra, dh, dW, db, n_enc, dpts01, denc, dnormal, stream);
}

static int linear64_bwd_impl(const float *h, int64_t n, const float *W, const float *dout, int64_t ld_dout, int32_t n_out,
                             const float *dextra, int32_t n_extra, float *dh, float *dW, float *db, int n_tail_enc, float *dpts01,
                             float *denc, float *dnormal, void *stream)
{
    IA_REQUIRE(n_extra >= 0 && n_extra <= n_out && (n_extra == 0 || dextra != nullptr), "linear64_bwd: bad dextra / n_extra=%d", n_extra);
    IA_REQUIRE(n_out >= 1 && n_out <= MAX_NOUT, "linear64_bwd: n_out %d not in [1,%d]", n_out, MAX_NOUT);
    IA_REQUIRE(n >= 0 && (n == 0 || (h && W && dout)), "linear64_bwd: NULL pointer");
    IA_REQUIRE(ld_dout >= n_out, "linear64_bwd: ld_dout < n_out");
    IA_REQUIRE(dh == nullptr || ((uintptr_t)dh & 15) == 0, "linear64_bwd: dh must be 16-byte aligned");
    IA_REQUIRE(((uintptr_t)h & 15) == 0, "linear64_bwd: h must be 16-byte aligned");
    if (n == 0) return IA_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)std::min<int64_t>(ia_ceil_div(n, LROWS), (int64_t)ia_sm_count() * 2);
    const bool tiled = n_out <= TILED_NOUT;
    if (dh != nullptr) {
        if (tiled) {
            const size_t bytes = sizeof(float) * ((size_t)TILED_NOUT * WPAD + (size_t)LROWS * OLD);
            IA_CUDA_OK(cudaFuncSetAttribute(linear64_bwd_input_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            linear64_bwd_input_tiled_kernel<<<blocks, LTHREADS, bytes, s>>>(dout, ld_dout, dextra, n_extra, n, W, n_out, dh, n_tail_enc, dpts01,
                                                                            denc, dnormal);
            IA_LAUNCH_OK("linear64_bwd_input_tiled_kernel");
        } else {
            const size_t bytes = sizeof(float) * ((size_t)n_out * WPAD + (size_t)LROWS * (n_out + 1));
            IA_CUDA_OK(cudaFuncSetAttribute(linear64_bwd_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            linear64_bwd_input_kernel<<<blocks, LTHREADS, bytes, s>>>(dout, ld_dout, dextra, n_extra, n, W, n_out, dh);
            IA_LAUNCH_OK("linear64_bwd_input_kernel");
        }
    }
    if (dW != nullptr) {
        if (tiled) {
            const size_t bytes = sizeof(float) * ((size_t)LROWS * LW + (size_t)LROWS * DLD);
            IA_CUDA_OK(cudaFuncSetAttribute(linear64_bwd_weight_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            const unsigned wblocks = (unsigned)std::min<int64_t>(ia_ceil_div(n, LROWS), (int64_t)ia_sm_count() * 3);
            linear64_bwd_weight_tiled_kernel<<<wblocks, WG_THREADS, bytes, s>>>(h, dout, ld_dout, dextra, n_extra, n, n_out, dW, db);
            IA_LAUNCH_OK("linear64_bwd_weight_tiled_kernel");
            return IA_OK;
        }
        const int noblk = (n_out + 15) / 16;
        const size_t bytes = sizeof(float) * ((size_t)LROWS * WPAD + (size_t)LROWS * (16 * noblk + 1));
        const unsigned wblocks = (unsigned)std::min<int64_t>(ia_ceil_div(n, LROWS), (int64_t)ia_sm_count() * (bytes <= 72 * 1024 ? 3 : 2));
#define IA_L64W(NB)                                                                                                              \
    case NB:                                                                                                                     \
        IA_CUDA_OK(cudaFuncSetAttribute(linear64_bwd_weight_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
        linear64_bwd_weight_kernel<NB><<<wblocks, LTHREADS, bytes, s>>>(h, dout, ld_dout, dextra, n_extra, n, n_out, dW, db);   \
        break;
        switch (noblk) {
            IA_L64W(5) IA_L64W(6) IA_L64W(7) IA_L64W(8)
        }
#undef IA_L64W
        IA_LAUNCH_OK("linear64_bwd_weight_kernel");
    }
    return IA_OK;
}
