// grad_type 'analytic' on the tensor cores: the SDF network together with d sdf / d(input), and the adjoint of that pair
// (a second-order adjoint of the network).  Reference: models/geometry.py:206 (network(encoding(points))) and :214-218
// (torch.autograd.grad(sdf, points, create_graph=True)): training differentiates THROUGH the normal.
//
// Network (3 xyz + 32 hash features -> 64 -> 64 -> n_out, Softplus(beta = 100); column 0 of the output layer is the SDF):
//     z1 = W0 x + b0,  h1 = sp(z1),  s1 = sp'(z1) = sigmoid(beta z1)
//     z2 = W1 h1 + b1, h2 = sp(z2),  s2 = sigmoid(beta z2)            y = wl . h2 + bl          (wl = row 0 of the last layer)
//     a2 = wl * s2,    u1 = W1^T a2, a1 = u1 * s1,   g = W0^T a1 = dy/dx
// ia_mlp_fwd_grad returns h2 (the wide output layer is applied by linear64.cu, as for the first-order centre evaluation) and g.
// ia_mlp_fwd_grad_bwd takes the cotangents (dh2, dg) and returns d(input), d(parameters):
//     A1 = W0 dg  (= the tangent of z1 along dg),  U1 = A1 * s1,  A2 = W1 U1,
//     dwl += A2 * s2,   dz2 = dh2 * s2 + A2 * wl * beta s2 (1 - s2),   dh1 = W1^T dz2,
//     dz1 = dh1 * s1 + A1 * u1 * beta s1 (1 - s1),   dx = W0^T dz1,
//     dW1 += a2 U1^T + dz2 [h1 | 1]^T,   dW0 += a1 dg^T + dz1 [x | 1]^T
// i.e. seven row GEMMs and four dW GEMMs per 128-row tile, all as 3xF16-split tcgen05.mma with fp32 TMEM accumulators
// (mlp_tc_device.cuh).  One CTA = 512 threads = one tile at a time (thread (row r, column group cg) owns 16 accumulator
// columns), persistent over tiles; z1, z2, A1 and u1 stay in TMEM for the whole tile (they are re-read by later phases instead
// of being spilled to shared memory), the dW accumulators persist in TMEM across all tiles of the CTA.
// Gradient-like operands carry one launch-wide power-of-two scale S (fp16 normal range), derived from abs-max pre-passes
// over dh2 / dg and norm bounds of the weights; it is removed exactly when results leave the kernel.
#include "mlp_tc_device.cuh"

int ia_tc_absmax_slot(const float *v, int64_t n, int cols, int64_t ld, cudaStream_t stream, float **slot_out);   // mlp_tc.cu

namespace {

constexpr int T2 = 512;
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
// TMEM columns
constexpr uint32_t C_Z1 = 0, C_Z2 = 64, C_A1 = 128, C_U1 = 192, C_WK = 256, C_DX = 320, C_DW1 = 368, C_DW0 = 440;

struct Plan2 {
    uint32_t x_hi, x_lo, g_hi, g_lo, h_hi, h_lo, a_hi, a_lo, b_hi, b_lo, c_hi, c_lo, operands_end;
    uint32_t w0_hi, w0_lo, w1_hi, w1_lo, wl, b0, b1, dwl, red, mbar, tmem, ops, total;
};

__host__ __device__ inline Plan2 make_plan2(bool bwd)
{
    Plan2 p;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
    p.x_hi = take(6 * 2048); p.x_lo = take(6 * 2048);
    p.h_hi = take(9 * 2048); p.h_lo = take(9 * 2048);
    p.a_hi = take(8 * 2048); p.a_lo = take(8 * 2048);
    p.g_hi = take(bwd ? 6 * 2048 : 0); p.g_lo = take(bwd ? 6 * 2048 : 0);
    p.b_hi = take(bwd ? 8 * 2048 : 0); p.b_lo = take(bwd ? 8 * 2048 : 0);
    p.c_hi = take(bwd ? 8 * 2048 : 0); p.c_lo = take(bwd ? 8 * 2048 : 0);
    p.operands_end = o;
    p.w0_hi = take(6 * 1024); p.w0_lo = take(6 * 1024);
    p.w1_hi = take(8 * 1024); p.w1_lo = take(8 * 1024);
    p.wl = take(W * 4); p.b0 = take(W * 4); p.b1 = take(W * 4);
    p.dwl = take(W * 4);
    p.red = take(4 * 64 * 4);
    p.mbar = take(64);
    p.tmem = take(16);
    p.ops = take(16 * sizeof(Operand));
    p.total = o;
    return p;
}

__device__ __forceinline__ float rcp_ftz(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Softplus(beta) pieces from the pre-activation z: e = exp(-|beta z|), r = 1 / (1 + e)
//   sigmoid = z >= 0 ? r : e r ;  sigmoid (1 - sigmoid) = e r^2 ;  softplus = max(z, 0) + log(1 + e) / beta
struct Sp { float e, r, t; };
__device__ __forceinline__ Sp sp_parts(float z)
{
    Sp p;
    p.t = z * (BETA * LOG2E);
    p.e = ex2_ftz(-fabsf(p.t));
    p.r = rcp_ftz(1.0f + p.e);
    return p;
}
__device__ __forceinline__ float sp_sigmoid(const Sp &p) { return p.t >= 0.f ? p.r : p.e * p.r; }
__device__ __forceinline__ float sp_curv(const Sp &p) { return BETA * p.e * p.r * p.r; }          // beta s (1 - s)

struct Common2 {
    char *smem;
    uint32_t sbase, tmem, lane_addr;
    int tid, lane, q, cg, r, c0;
    Plan2 P;
};

// stage weights / biases, mbarriers, TMEM, operand descriptors; zero every operand buffer (also the dW accumulators when bwd)
__device__ __forceinline__ void setup2(Common2 &c, const TcDims &D, const float *__restrict__ params, bool bwd, uint32_t tmem_cols)
{
    const Plan2 &P = c.P;
    char *smem = c.smem;
    const int tid = c.tid;
    for (uint32_t i = (uint32_t)tid * 16u; i < P.operands_end; i += T2 * 16u) *reinterpret_cast<uint4 *>(smem + i) = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < W * 6; i += T2) {          // W0 [64][35] -> split K-major B operand, 48 columns, internal order
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = 8 * c8 + j;
            a[j] = col < D.din ? __ldg(params + D.pW0 + o * D.din + global_col(col, D.n_in0, D.n_in1)) : 0.f;
        }
        store_split8(smem + P.w0_hi, smem + P.w0_lo, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
    for (int i = tid; i < W * 8; i += T2) {
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __ldg(params + D.pW1 + o * W + 8 * c8 + j);
        store_split8(smem + P.w1_hi, smem + P.w1_lo, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
    float *wl_s = reinterpret_cast<float *>(smem + P.wl), *b0_s = reinterpret_cast<float *>(smem + P.b0);
    float *b1_s = reinterpret_cast<float *>(smem + P.b1), *dwl = reinterpret_cast<float *>(smem + P.dwl);
    if (tid < W) {
        wl_s[tid] = __ldg(params + D.pWl + tid);
        b0_s[tid] = __ldg(params + D.pb0 + tid);
        b1_s[tid] = __ldg(params + D.pb1 + tid);
        dwl[tid] = 0.f;
    }
    if (tid < 8) mbar_init(c.sbase + P.mbar + 8u * (uint32_t)tid, 1);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(c.sbase + P.tmem), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 0) {
        Operand *ops = reinterpret_cast<Operand *>(smem + P.ops);
        const uint32_t s = c.sbase;
        ops[0] = act_as_A_kmajor(s + P.x_hi, s + P.x_lo);
        ops[1] = act_as_A_kmajor(s + P.h_hi, s + P.h_lo);
        ops[2] = act_as_A_kmajor(s + P.a_hi, s + P.a_lo);
        ops[3] = act_as_mnmajor(s + P.x_hi, s + P.x_lo);
        ops[4] = act_as_mnmajor(s + P.h_hi, s + P.h_lo);
        ops[5] = act_as_mnmajor(s + P.a_hi, s + P.a_lo);
        ops[6] = w_as_B_kmajor(s + P.w0_hi, s + P.w0_lo);
        ops[7] = w_as_B_kmajor(s + P.w1_hi, s + P.w1_lo);
        ops[8] = w_as_B_mnmajor(s + P.w0_hi, s + P.w0_lo);
        ops[9] = w_as_B_mnmajor(s + P.w1_hi, s + P.w1_lo);
        if (bwd) {
            ops[10] = act_as_A_kmajor(s + P.g_hi, s + P.g_lo);
            ops[11] = act_as_mnmajor(s + P.g_hi, s + P.g_lo);
            ops[12] = act_as_mnmajor(s + P.b_hi, s + P.b_lo);
            ops[13] = act_as_A_kmajor(s + P.c_hi, s + P.c_lo);
            ops[14] = act_as_mnmajor(s + P.c_hi, s + P.c_lo);
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *reinterpret_cast<volatile uint32_t *>(smem + P.tmem);
    c.lane_addr = ((uint32_t)c.q * 32u) << 16;
}

// all threads: operand writes visible to the async proxy, CTA barrier, one lane of warp 0 issues
template <typename F>
__device__ __forceinline__ void sync_issue(const Common2 &c, F issue)
{
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (c.tid < 32) {
        if (elect_one()) {
            tc_fence_after();
            issue();
        }
        __syncwarp();
    }
}

struct Bar {
    uint32_t addr, phase;
    __device__ __forceinline__ void wait() { mbar_wait(addr, phase); phase ^= 1u; tc_fence_after(); }
};

// this thread's slice of the input row: feature chunk cg (8 floats) and, for cg == 0, xyz
struct XRegs { float f[8], xyz[3]; };
__device__ __forceinline__ void load_x2(XRegs &R, const float *__restrict__ in0, const float *__restrict__ in1, int64_t row, bool valid, int cg)
{
    const float4 *src = reinterpret_cast<const float4 *>(in1 + row * 32 + 8 * cg);
    const float4 a = valid ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f), b = valid ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
    R.f[0] = a.x; R.f[1] = a.y; R.f[2] = a.z; R.f[3] = a.w; R.f[4] = b.x; R.f[5] = b.y; R.f[6] = b.z; R.f[7] = b.w;
    if (cg == 0) {
#pragma unroll
        for (int j = 0; j < 3; ++j) R.xyz[j] = valid ? __ldg(in0 + row * 3 + j) : 0.f;
    }
}
__device__ __forceinline__ void store_x2(const Common2 &c, const TcDims &D, const XRegs &R, bool valid)
{
    char *hi = c.smem + c.P.x_hi, *lo = c.smem + c.P.x_lo;
    store_split8(hi, lo, (uint32_t)c.cg * 2048u + (uint32_t)c.r * 16u, R.f);
    if (c.cg == 0) {
        const float a[8] = {valid ? fmaf(R.xyz[0], D.s0, D.o0) : 0.f, valid ? fmaf(R.xyz[1], D.s0, D.o0) : 0.f,
                            valid ? fmaf(R.xyz[2], D.s0, D.o0) : 0.f, valid ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f};
        store_split8(hi, lo, 4u * 2048u + (uint32_t)c.r * 16u, a);
    }
}

__device__ __forceinline__ void store16(const Common2 &c, uint32_t hi_off, uint32_t lo_off, const float (&v)[16])
{
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = v[8 * hf + j];
        store_split8(c.smem + hi_off, c.smem + lo_off, (uint32_t)(2 * c.cg + hf) * 2048u + (uint32_t)c.r * 16u, a);
    }
}

// s1 = sigmoid(beta z1) (and beta s1 (1 - s1)) for this thread's 16 columns, from the pre-activations that stay in TMEM for the
// whole tile.  (1 - exp(-beta h1) from the H1 operand buffer would save one MUFU per element, but it cancels for small s1:
// absolute error 6e-8, i.e. 6e-6 relative at s1 = 0.01, which the second-order chain multiplies up.)
template <bool CURV>
__device__ __forceinline__ void s1_from_z1(uint32_t taddr, const float *__restrict__ b0c, float (&s1)[16], float (&curv)[CURV ? 16 : 1])
{
    float z[16];
    tmem_ld16(taddr, z);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const Sp p = sp_parts(z[j] + b0c[j]);
        s1[j] = sp_sigmoid(p);
        if (CURV) curv[j] = sp_curv(p);
    }
}

// Norm bounds of the weights, every thread gets them: r0 / r1 = max row L1 norm of W0 / W1, c1 = max column L1 norm of W1,
// wlmax = max |wl|.  `red` = 256 floats of shared memory; contains a CTA barrier.
struct Norms { float r0, r1, c1, wlmax; };
__device__ __forceinline__ Norms weight_norms(const Common2 &c, const TcDims &D, const float *__restrict__ params, float *red)
{
    if (c.tid < 64) {
        float s = 0.f;
        for (int k = 0; k < D.din; ++k) s += fabsf(__ldg(params + D.pW0 + c.tid * D.din + k));
        red[c.tid] = s;
    } else if (c.tid < 128) {
        const int j = c.tid - 64;
        float s = 0.f;
        for (int k = 0; k < W; ++k) s += fabsf(__ldg(params + D.pW1 + j * W + k));
        red[c.tid] = s;
    } else if (c.tid < 192) {
        const int i = c.tid - 128;
        float s = 0.f;
        for (int k = 0; k < W; ++k) s += fabsf(__ldg(params + D.pW1 + k * W + i));
        red[c.tid] = s;
    } else if (c.tid < 256) {
        red[c.tid] = fabsf(__ldg(params + D.pWl + (c.tid - 192)));
    }
    __syncthreads();
    Norms nm{0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < 64; ++k) {
        nm.r0 = fmaxf(nm.r0, red[k]); nm.r1 = fmaxf(nm.r1, red[64 + k]); nm.c1 = fmaxf(nm.c1, red[128 + k]);
        nm.wlmax = fmaxf(nm.wlmax, red[192 + k]);
    }
    return nm;
}

// power-of-two scale that puts the bound x > 0 into [2^12, 2^13): fp16 operands written with it cannot overflow (max 65504) and
// keep the (hi, lo) pair's ~22 bits for everything within 2^-12 of the bound (fp16 subnormals end at 2^-24)
__device__ __forceinline__ void pow2_scale13(float x, float &scale, float &inv_scale)
{
    pow2_scale(x, scale, inv_scale);
    scale *= 128.f;
    inv_scale *= (1.f / 128.f);
}

// ------------------------------------------------------------------------------------------------------------
// forward: h2 [n,64] and g = d y / d(input) (g1 [n,32] w.r.t. in1, g0 [n,3] w.r.t. the raw in0)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T2, 1)
mlp_tc_fwd_grad_kernel(const TcDims Din, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
                       const float *__restrict__ params, float *__restrict__ h_out, float *__restrict__ g0, float *__restrict__ g1)
{
    const TcDims D = specialise<1>(Din);
    extern __shared__ __align__(1024) char smem[];
    Common2 c;
    c.smem = smem; c.sbase = smem_u32(smem); c.P = make_plan2(false);
    c.tid = threadIdx.x; c.lane = c.tid & 31;
    const int w = c.tid >> 5;
    c.q = w & 3; c.cg = w >> 2; c.r = 32 * c.q + c.lane; c.c0 = 16 * c.cg;
    setup2(c, D, params, false, 256u);
    const Plan2 &P = c.P;
    const Operand *ops = reinterpret_cast<const Operand *>(smem + P.ops);
    const Operand &AX = ops[0], &AH = ops[1], &AA = ops[2], &BW0 = ops[6], &BW1 = ops[7], &BW0T = ops[8], &BW1T = ops[9];
    const float *wl_s = reinterpret_cast<const float *>(smem + P.wl), *b0_s = reinterpret_cast<const float *>(smem + P.b0);
    const float *b1_s = reinterpret_cast<const float *>(smem + P.b1);
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0), idesc_dh = make_idesc(128, W, 0, 1), idesc_dx = make_idesc(128, 48, 0, 1);
    Bar mb{c.sbase + P.mbar, 0u};
    const uint32_t tm = c.tmem + c.lane_addr;
    // operand scales of the gradient chain: |a2| <= wlmax, |u1|, |a1| <= c1 wlmax
    float sa2, inv_sa2, sa1, inv_sa1;
    {
        const Norms nm = weight_norms(c, D, params, reinterpret_cast<float *>(smem + P.red));
        pow2_scale13(fmaxf(nm.wlmax, 1e-30f), sa2, inv_sa2);
        pow2_scale13(fmaxf(nm.c1 * nm.wlmax, 1e-30f), sa1, inv_sa1);
    }

    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    XRegs X;
    int64_t tile = blockIdx.x;
    if (tile < n_tiles) load_x2(X, in0, in1, tile * ROWS + c.r, tile * ROWS + c.r < n, c.cg);
    for (; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS + c.r;
        const bool valid = row < n;
        store_x2(c, D, X, valid);
        sync_issue(c, [&]() { issue_gemm(c.tmem + C_Z1, AX, BW0, idesc_fwd, 48 / 16); umma_commit(mb.addr); });
        {
            const int64_t nrow = (tile + gridDim.x) * ROWS + c.r;
            const bool nvalid = tile + gridDim.x < n_tiles && nrow < n;
            load_x2(X, in0, in1, nvalid ? nrow : 0, nvalid, c.cg);
        }
        mb.wait();
        {   // h1 -> H
            float v[16];
            tmem_ld16(tm + C_Z1 + (uint32_t)c.c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_fwd<IA_ACT_SOFTPLUS100>(v[j] + b0_s[c.c0 + j]);
            store16(c, P.h_hi, P.h_lo, v);
        }
        sync_issue(c, [&]() { issue_gemm(c.tmem + C_Z2, AH, BW1, idesc_fwd, W / 16); umma_commit(mb.addr); });
        mb.wait();
        {   // h2 -> global, a2 = wl * s2 -> A
            float v[16], a2[16];
            tmem_ld16(tm + C_Z2 + (uint32_t)c.c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float z = v[j] + b1_s[c.c0 + j];
                const Sp p = sp_parts(z);
                a2[j] = wl_s[c.c0 + j] * sa2 * sp_sigmoid(p);
                v[j] = fmaxf(z, 0.f) + lg2_ftz(1.0f + p.e) * (LN2 / BETA);
            }
            if (valid && h_out != nullptr) {
                float4 *dst = reinterpret_cast<float4 *>(h_out + row * W + c.c0);
#pragma unroll
                for (int k = 0; k < 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            }
            store16(c, P.a_hi, P.a_lo, a2);
        }
        sync_issue(c, [&]() { issue_gemm(c.tmem + C_U1, AA, BW1T, idesc_dh, W / 16); umma_commit(mb.addr); });
        mb.wait();
        {   // a1 = u1 * s1 -> H (its own h1 entries are consumed here; the L1 GEMM that read H has completed)
            float v[16], s1[16], unused[1];
            s1_from_z1<false>(tm + C_Z1 + (uint32_t)c.c0, b0_s + c.c0, s1, unused);
            tmem_ld16(tm + C_U1 + (uint32_t)c.c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= s1[j] * (inv_sa2 * sa1);
            store16(c, P.h_hi, P.h_lo, v);
        }
        sync_issue(c, [&]() { issue_gemm(c.tmem + C_DX, AH, BW0T, idesc_dx, W / 16); umma_commit(mb.addr); });
        mb.wait();
        if (c.cg < 3) {   // g: internal columns [0,32) -> g1, [32,35) -> g0 (chain rule through in0 * s0 + o0)
            float v[16];
            tmem_ld16(tm + C_DX + (uint32_t)c.c0, v);
            if (valid) {
                if (c.cg < 2) {
                    if (g1 != nullptr) {
                        float4 *dst = reinterpret_cast<float4 *>(g1 + row * 32 + c.c0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            dst[k] = make_float4(v[4 * k] * inv_sa1, v[4 * k + 1] * inv_sa1, v[4 * k + 2] * inv_sa1, v[4 * k + 3] * inv_sa1);
                    }
                } else if (g0 != nullptr) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) g0[row * 3 + j] = v[j] * (D.s0 * inv_sa1);
                }
            }
        }
        tc_fence_before();      // the next tile's first GEMM overwrites C_Z1 only; C_DX is rewritten three barriers later
    }
    tc_fence_before();
    __syncthreads();
    if (c.tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(c.tmem), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// adjoint of (h2, g) w.r.t. inputs and parameters
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(T2, 1)
mlp_tc_fwd_grad_bwd_kernel(const TcDims Din, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
                           const float *__restrict__ params, const float *__restrict__ dh, const float *__restrict__ dg0,
                           const float *__restrict__ dg1, float *__restrict__ din0, float *__restrict__ din1,
                           float *__restrict__ dparams, const float *__restrict__ hmax_p, const float *__restrict__ gmax0_p,
                           const float *__restrict__ gmax1_p)
{
    const TcDims D = specialise<1>(Din);
    extern __shared__ __align__(1024) char smem[];
    Common2 c;
    c.smem = smem; c.sbase = smem_u32(smem); c.P = make_plan2(true);
    c.tid = threadIdx.x; c.lane = c.tid & 31;
    const int w = c.tid >> 5;
    c.q = w & 3; c.cg = w >> 2; c.r = 32 * c.q + c.lane; c.c0 = 16 * c.cg;
    setup2(c, D, params, true, 512u);
    const Plan2 &P = c.P;
    const Operand *ops = reinterpret_cast<const Operand *>(smem + P.ops);
    const Operand &AX = ops[0], &AH = ops[1], &AA = ops[2], &XT = ops[3], &HT = ops[4], &AT = ops[5];
    const Operand &BW0 = ops[6], &BW1 = ops[7], &BW0T = ops[8], &BW1T = ops[9];
    const Operand &AG = ops[10], &GT = ops[11], &BT = ops[12], &AC = ops[13], &CT = ops[14];
    const float *wl_s = reinterpret_cast<const float *>(smem + P.wl), *b0_s = reinterpret_cast<const float *>(smem + P.b0);
    const float *b1_s = reinterpret_cast<const float *>(smem + P.b1);
    float *dwl = reinterpret_cast<float *>(smem + P.dwl), *red = reinterpret_cast<float *>(smem + P.red);
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0), idesc_dh = make_idesc(128, W, 0, 1), idesc_dx = make_idesc(128, 48, 0, 1);
    const uint32_t idesc_w64 = make_idesc(64, 64, 1, 1), idesc_w72 = make_idesc(64, 72, 1, 1), idesc_w48 = make_idesc(64, 48, 1, 1);
    Bar mbA{c.sbase + P.mbar, 0u}, mbB{c.sbase + P.mbar + 8u, 0u}, mbC{c.sbase + P.mbar + 16u, 0u}, mbD{c.sbase + P.mbar + 24u, 0u},
        mbG{c.sbase + P.mbar + 32u, 0u};
    const uint32_t tm = c.tmem + c.lane_addr;

    // ---- launch-wide gradient scale from norm bounds of the weights (every factor below is an upper bound: |s| <= 1,
    //      beta s (1 - s) <= beta / 4)
    const Norms nm = weight_norms(c, D, params, red);
    // zero the persistent dW accumulators: every operand buffer is zero at this point
    sync_issue(c, [&]() {
        issue_gemm(c.tmem + C_DW1, AT, HT, idesc_w72, ROWS / 16);
        issue_gemm(c.tmem + C_DW0, AT, XT, idesc_w48, ROWS / 16);
        umma_commit(mbA.addr);
    });
    float scale, inv_scale;
    {
        const float r0 = nm.r0, r1 = nm.r1, c1 = nm.c1, wlmax = nm.wlmax;
        const float hm = hmax_p ? __ldg(hmax_p) : 0.f;
        const float gm = fmaxf(gmax1_p ? __ldg(gmax1_p) : 0.f, (gmax0_p ? __ldg(gmax0_p) : 0.f) * fabsf(D.s0));
        const float b_a1 = r0 * gm, b_a2 = r1 * b_a1, b_z2 = hm + 0.25f * BETA * wlmax * b_a2, b_h1 = c1 * b_z2;
        const float b_z1 = b_h1 + 0.25f * BETA * b_a1 * c1 * wlmax;
        const float bound = fmaxf(fmaxf(fmaxf(gm, hm), fmaxf(b_a1, b_a2)), fmaxf(fmaxf(b_z2, b_h1), b_z1));
        pow2_scale13(fmaxf(bound, 1e-30f), scale, inv_scale);
    }
    mbA.wait();
    if (c.r < ROWS && c.cg == 0) {   // constant columns: H chunk 8 = (1, 0, ..., 0): bias gradient of layer 1
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_split8(smem + P.h_hi, smem + P.h_lo, 8u * 2048u + (uint32_t)c.r * 16u, one);
    }

    float gwl[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) gwl[j] = 0.f;
    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    XRegs X;
    float G[8], g0r[3];
    auto load_g = [&](int64_t row, bool valid) {
        if (dg1 != nullptr && valid) {
            const float4 *src = reinterpret_cast<const float4 *>(dg1 + row * 32 + 8 * c.cg);
            const float4 a = __ldg(src), b = __ldg(src + 1);
            G[0] = a.x; G[1] = a.y; G[2] = a.z; G[3] = a.w; G[4] = b.x; G[5] = b.y; G[6] = b.z; G[7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) G[j] = 0.f;
        }
        if (c.cg == 1) {
#pragma unroll
            for (int j = 0; j < 3; ++j) g0r[j] = (dg0 != nullptr && valid) ? __ldg(dg0 + row * 3 + j) : 0.f;
        }
    };
    int64_t tile = blockIdx.x;
    const bool had_tiles = tile < n_tiles;
    if (had_tiles) {
        const int64_t row = tile * ROWS + c.r;
        load_x2(X, in0, in1, row, row < n, c.cg);
        load_g(row, row < n);
    }
    for (; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS + c.r;
        const bool valid = row < n;
        // ---- P0: X, S dg -> smem;  z1 = X W0^T,  A1 = (S dg) W0^T
        store_x2(c, D, X, valid);
        {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = G[j] * scale;
            store_split8(smem + P.g_hi, smem + P.g_lo, (uint32_t)c.cg * 2048u + (uint32_t)c.r * 16u, a);
            if (c.cg == 1) {     // cotangent of g0 (raw in0) -> cotangent of the scaled input column
                const float b[8] = {g0r[0] * D.s0 * scale, g0r[1] * D.s0 * scale, g0r[2] * D.s0 * scale, 0.f, 0.f, 0.f, 0.f, 0.f};
                store_split8(smem + P.g_hi, smem + P.g_lo, 4u * 2048u + (uint32_t)c.r * 16u, b);
            }
        }
        sync_issue(c, [&]() {
            issue_gemm(c.tmem + C_Z1, AX, BW0, idesc_fwd, 48 / 16);
            umma_commit(mbA.addr);
            issue_gemm(c.tmem + C_A1, AG, BW0, idesc_fwd, 48 / 16);
            umma_commit(mbG.addr);
        });
        {
            const int64_t nrow = (tile + gridDim.x) * ROWS + c.r;
            const bool nvalid = tile + gridDim.x < n_tiles && nrow < n;
            load_x2(X, in0, in1, nvalid ? nrow : 0, nvalid, c.cg);
            load_g(nvalid ? nrow : 0, nvalid);
        }
        mbA.wait();
        // ---- P1: h1 -> H;  z2 = H W1^T
        {
            float v[16];
            tmem_ld16(tm + C_Z1 + (uint32_t)c.c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = act_fwd<IA_ACT_SOFTPLUS100>(v[j] + b0_s[c.c0 + j]);
            store16(c, P.h_hi, P.h_lo, v);
        }
        sync_issue(c, [&]() { issue_gemm(c.tmem + C_Z2, AH, BW1, idesc_fwd, W / 16); umma_commit(mbA.addr); });
        mbA.wait();
        // ---- P2: a2 = wl * s2 -> A;  u1 = A W1
        {
            float v[16];
            tmem_ld16(tm + C_Z2 + (uint32_t)c.c0, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = wl_s[c.c0 + j] * sp_sigmoid(sp_parts(v[j] + b1_s[c.c0 + j]));
            store16(c, P.a_hi, P.a_lo, v);
        }
        sync_issue(c, [&]() { issue_gemm(c.tmem + C_U1, AA, BW1T, idesc_dh, W / 16); umma_commit(mbA.addr); });
        mbA.wait();
        mbG.wait();
        // ---- P3: a1 = u1 * s1 -> B,  U1 = A1 * s1 -> C;  A2 = C W1^T;  dW1 += A^T C,  dW0 += B^T (S dg)
        {
            float u[16], a1[16], s1[16], unused[1];
            s1_from_z1<false>(tm + C_Z1 + (uint32_t)c.c0, b0_s + c.c0, s1, unused);
            tmem_ld16(tm + C_U1 + (uint32_t)c.c0, u);
            tmem_ld16(tm + C_A1 + (uint32_t)c.c0, a1);
#pragma unroll
            for (int j = 0; j < 16; ++j) { u[j] *= s1[j]; a1[j] *= s1[j]; }
            store16(c, P.b_hi, P.b_lo, u);
            store16(c, P.c_hi, P.c_lo, a1);
        }
        sync_issue(c, [&]() {
            issue_gemm(c.tmem + C_WK, AC, BW1, idesc_fwd, W / 16);
            umma_commit(mbA.addr);
            issue_gemm_acc(c.tmem + C_DW1, AT, CT, idesc_w64, ROWS / 16, 1u);
            issue_gemm_acc(c.tmem + C_DW0, BT, GT, idesc_w48, ROWS / 16, 1u);
            umma_commit(mbB.addr);
        });
        mbA.wait();
        // ---- P4: dwl += A2 * s2;  dz2 = S dh * s2 + A2 * wl * beta s2 (1 - s2) -> A;  dh1 = A W1;  dW1 += A^T [H | 1]
        {
            float a2[16], z[16], dhr[16];
            if (dh != nullptr && valid) {
                const float4 *src = reinterpret_cast<const float4 *>(dh + row * W + c.c0);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 t4 = __ldg(src + k);
                    dhr[4 * k] = t4.x; dhr[4 * k + 1] = t4.y; dhr[4 * k + 2] = t4.z; dhr[4 * k + 3] = t4.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) dhr[j] = 0.f;
            }
            tmem_ld16(tm + C_WK + (uint32_t)c.c0, a2);
            tmem_ld16(tm + C_Z2 + (uint32_t)c.c0, z);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const Sp p = sp_parts(z[j] + b1_s[c.c0 + j]);
                const float s2 = sp_sigmoid(p);
                gwl[j] = fmaf(a2[j], s2, gwl[j]);
                z[j] = fmaf(dhr[j] * scale, s2, a2[j] * wl_s[c.c0 + j] * sp_curv(p));
            }
            mbB.wait();          // the dW GEMMs of P3 have read A, B, C and the dg buffer
            store16(c, P.a_hi, P.a_lo, z);
        }
        sync_issue(c, [&]() {
            issue_gemm(c.tmem + C_WK, AA, BW1T, idesc_dh, W / 16);
            umma_commit(mbA.addr);
            issue_gemm_acc(c.tmem + C_DW1, AT, HT, idesc_w72, ROWS / 16, 1u);
            umma_commit(mbC.addr);
        });
        mbA.wait();
        // ---- P5: dz1 = dh1 * s1 + A1 * u1 * beta s1 (1 - s1) -> C;  dx = C W0;  dW0 += C^T [X | 1]
        {
            float v[16], a1[16], u[16], s1[16], curv[16];
            s1_from_z1<true>(tm + C_Z1 + (uint32_t)c.c0, b0_s + c.c0, s1, curv);
            tmem_ld16(tm + C_WK + (uint32_t)c.c0, v);
            tmem_ld16(tm + C_A1 + (uint32_t)c.c0, a1);
            tmem_ld16(tm + C_U1 + (uint32_t)c.c0, u);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], s1[j], a1[j] * u[j] * curv[j]);
            store16(c, P.c_hi, P.c_lo, v);
        }
        sync_issue(c, [&]() {
            issue_gemm(c.tmem + C_DX, AC, BW0T, idesc_dx, W / 16);
            umma_commit(mbA.addr);
            issue_gemm_acc(c.tmem + C_DW0, CT, XT, idesc_w48, ROWS / 16, 1u);
            umma_commit(mbD.addr);
        });
        mbA.wait();
        // ---- P6: dx -> global
        if (c.cg < 3 && (din0 != nullptr || din1 != nullptr)) {
            float v[16];
            tmem_ld16(tm + C_DX + (uint32_t)c.c0, v);
            if (valid) write_dx16(D, c.c0, v, inv_scale, row, din0, din1);
        }
        mbC.wait();
        mbD.wait();          // every operand buffer is free for the next tile
    }
    // ---- drain the persistent dW accumulators (M = 64 layout: output row o = 16 q + lane for lane < 16)
    if (had_tiles && dparams != nullptr) {
        const int o = 16 * c.q + c.lane;
        for (int ci = c.cg; ci * 16 < 72; ci += 4) {
            const int col0 = 16 * ci, m = 72 - col0 >= 16 ? 16 : 8;
            float v[16];
            if (m == 16) tmem_ld16(tm + C_DW1 + (uint32_t)col0, v);
            else tmem_ld8(tm + C_DW1 + (uint32_t)col0, v);
            if (c.lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = col0 + j;
                    if (j < m) {
                        if (col < W) atomicAdd(dparams + D.pW1 + o * W + col, v[j] * inv_scale);
                        else if (col == W) atomicAdd(dparams + D.pb1 + o, v[j] * inv_scale);
                    }
                }
            }
        }
        for (int ci = c.cg; ci * 16 < 48; ci += 4) {
            const int col0 = 16 * ci;
            float v[16];
            tmem_ld16(tm + C_DW0 + (uint32_t)col0, v);
            if (c.lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = col0 + j;
                    if (col < D.din) atomicAdd(dparams + D.pW0 + o * D.din + global_col(col, D.n_in0, D.n_in1), v[j] * inv_scale);
                    else if (col == D.din) atomicAdd(dparams + D.pb0 + o, v[j] * inv_scale);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float tot = warp_sum(gwl[j]);
            if (c.lane == 0) atomicAdd(&dwl[c.c0 + j], tot);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (had_tiles && dparams != nullptr && c.tid < W) atomicAdd(dparams + D.pWl + c.tid, dwl[c.tid] * inv_scale);
    if (c.tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(c.tmem), "r"(512u) : "memory");
}

int check_shape(const ia_mlp_desc *desc, TcDims *D)
{
    const int rc = make_dims(desc, 0, D);
    if (rc != IA_OK) return rc;
    if (!(D->n_in0 == 3 && D->n_in1 == 32 && D->nh == 2 && D->act == IA_ACT_SOFTPLUS100 && D->n_out >= 1)) {
        ia_set_error("ia_mlp_fwd_grad: implemented for the SDF network shape (3 + 32 inputs, two hidden layers of 64, Softplus(100))");
        return IA_ERR_UNSUPPORTED;
    }
    return IA_OK;
}

}  // namespace

extern "C" int32_t ia_mlp_fwd_grad(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                                   float *h_out, float *g0, float *g1, void *stream)
{
    TcDims D;
    const int rc = check_shape(desc, &D);
    if (rc != IA_OK) return rc;
    IA_REQUIRE(n >= 0, "ia_mlp_fwd_grad: n < 0");
    if (n == 0) return IA_OK;
    IA_REQUIRE(in0 && in1 && params, "ia_mlp_fwd_grad: NULL input");
    const Plan2 P = make_plan2(false);
    const int64_t n_tiles = ia_ceil_div(n, ROWS);
    const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ia_sm_count());
    IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_fwd_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.total));
    mlp_tc_fwd_grad_kernel<<<blocks, T2, P.total, (cudaStream_t)stream>>>(D, in0, in1, n, params, h_out, g0, g1);
    IA_LAUNCH_OK("mlp_tc_fwd_grad_kernel");
    return IA_OK;
}

extern "C" int32_t ia_mlp_fwd_grad_bwd(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                                       const float *dh, const float *dg0, const float *dg1, float *din0, float *din1,
                                       float *dparams, void *stream)
{
    TcDims D;
    const int rc = check_shape(desc, &D);
    if (rc != IA_OK) return rc;
    IA_REQUIRE(n >= 0, "ia_mlp_fwd_grad_bwd: n < 0");
    if (n == 0) return IA_OK;
    IA_REQUIRE(in0 && in1 && params, "ia_mlp_fwd_grad_bwd: NULL input");
    cudaStream_t s = (cudaStream_t)stream;
    float *hmax = nullptr, *gmax0 = nullptr, *gmax1 = nullptr;
    if (dh) { const int r2 = ia_tc_absmax_slot(dh, n, W, W, s, &hmax); if (r2 != IA_OK) return r2; }
    if (dg0) { const int r2 = ia_tc_absmax_slot(dg0, n, 3, 3, s, &gmax0); if (r2 != IA_OK) return r2; }
    if (dg1) { const int r2 = ia_tc_absmax_slot(dg1, n, 32, 32, s, &gmax1); if (r2 != IA_OK) return r2; }
    const Plan2 P = make_plan2(true);
    const int64_t n_tiles = ia_ceil_div(n, ROWS);
    const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ia_sm_count());
    IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_fwd_grad_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.total));
    mlp_tc_fwd_grad_bwd_kernel<<<blocks, T2, P.total, s>>>(D, in0, in1, n, params, dh, dg0, dg1, din0, din1, dparams, hmax, gmax0, gmax1);
    IA_LAUNCH_OK("mlp_tc_fwd_grad_bwd_kernel");
    return IA_OK;
}
