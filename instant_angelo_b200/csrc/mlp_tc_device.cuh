// Device-side building blocks shared by the tcgen05 MLP kernels (mlp_tc.cu: first-order forward / backward;
// mlp_tc2.cu: the analytic-gradient pair): PTX wrappers, the 3xF16 split operand layout, split GEMM issue, numerics.
#pragma once
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "ia_common.cuh"

namespace {

constexpr int ROWS = 128;
constexpr int CG = 4;                 // column groups: 4 threads share a row, 16 accumulator columns each
constexpr int THREADS = ROWS * CG;
constexpr int W = 64;
constexpr int MAX_OUT = 8;
constexpr float BETA = 100.f;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t D0_COL = 0;     // forward / dH / dX accumulator (<= 96 columns)
constexpr uint32_t D1_COL = 128;   // dW accumulator (<= 96 columns)

struct TcDims {
    int n_in0, n_in1, din, K0;  // K0 = round_up(din + 1, 16): column `din` holds the constant one (bias gradient)
    float s0, o0;
    int nh, n_out, nou, act;
    int pW0, pb0, pW1, pb1, pWl, pbl;
};

struct SmemPlan {  // byte offsets
    uint32_t ax_hi, ax_lo, ah_hi, ah_lo, dz_hi, dz_lo, w0_hi, w0_lo, w1_hi, w1_lo, wl, b0, b1, bl, dw0, dw1, dwl, dbl, red,
        part, mbar, tmem, ax2_hi, ax2_lo, ah2_hi, ah2_lo, total;
};

__host__ __device__ inline SmemPlan make_plan(const TcDims &D, bool bwd, bool pipe = false)
{
    SmemPlan p;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
    const uint32_t ax = (uint32_t)(D.K0 / 8) * 2048u;
    p.ax_hi = take(ax); p.ax_lo = take(ax);
    p.ah_hi = take(9 * 2048); p.ah_lo = take(9 * 2048);
    p.dz_hi = take(bwd ? 8 * 2048 : 0); p.dz_lo = take(bwd ? 8 * 2048 : 0);
    const uint32_t w0 = (uint32_t)(D.K0 / 8) * 1024u;
    p.w0_hi = take(w0); p.w0_lo = take(w0);
    p.w1_hi = take(8 * 1024); p.w1_lo = take(8 * 1024);
    p.wl = take(MAX_OUT * W * 4);
    p.b0 = take(W * 4); p.b1 = take(W * 4); p.bl = take(MAX_OUT * 4);
    p.dw0 = take(bwd ? (uint32_t)(W * (D.K0 + 1) * 4) : 0);
    p.dw1 = take(bwd ? W * 73 * 4 : 0);
    p.dwl = take(bwd ? MAX_OUT * W * 4 : 0);
    p.dbl = take(bwd ? MAX_OUT * 4 : 0);
    p.red = take(64 * 4);
    p.part = take(bwd ? 0 : (uint32_t)(CG * ROWS * MAX_OUT * 4));
    p.mbar = take(32);      // three mbarriers (8 bytes each)
    p.tmem = take(16);
    // software-pipelined backward: second set of X / H1 operand buffers (next tile's forward overlaps this tile's backward)
    p.ax2_hi = take(pipe ? ax : 0); p.ax2_lo = take(pipe ? ax : 0);
    p.ah2_hi = take(pipe ? 9 * 2048 : 0); p.ah2_lo = take(pipe ? 9 * 2048 : 0);
    p.total = o;
    return p;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // UMMA shared-memory matrix descriptor, SWIZZLE_NONE: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48)
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major)
{
    // kind::f16 instruction descriptor: D=f32 (1<<4), A=B=f16 (0), majors [15],[16], N>>3 [17,23), M>>4 [24,29)
    return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t mbar_addr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar_addr) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar_addr, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar_addr), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar_addr, uint32_t parity)
{
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(mbar_addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();  // never hang the GPU: a lost MMA completion is a bug, fail loudly
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[16])
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    tc_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- numerics ----------------------------------------------------------------------------------------------
// MUFU.EX2 / MUFU.LG2 without the denormal pre/post-scaling that __expf / __logf wrap around them (an FSETP and two
// FMULs per call): exp arguments here never produce results that matter below 2^-126, and lg2 sees 1 + e >= 1.
__device__ __forceinline__ float ex2_ftz(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_ftz(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float z)
{
    if (ACT == IA_ACT_SOFTPLUS100) {
        // torch.nn.Softplus(beta=100, threshold=20): z when beta z > 20, else log(1 + exp(beta z)) / beta
        constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
        const float t = z * (BETA * LOG2E);
        const float sp = lg2_ftz(1.0f + ex2_ftz(t)) * (LN2 / BETA);
        return t > 20.f * LOG2E ? z : sp;
    }
    return fmaxf(z, 0.f);
}

template <int ACT>
__device__ __forceinline__ float act_bwd_from_out(float h)
{
    // sigmoid(beta z) = 1 - exp(-beta h)
    if (ACT == IA_ACT_SOFTPLUS100) return 1.0f - ex2_ftz(h * (-BETA * 1.4426950408889634f));
    return h > 0.f ? 1.f : 0.f;
}

// store 8 consecutive K-values of a row as the fp16 (hi, lo) pair of 16-byte core-matrix rows
__device__ __forceinline__ void store_split8(char *hi_base, char *lo_base, uint32_t off, const float (&a)[8])
{
    __half2 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = __floats2half2_rn(a[2 * i], a[2 * i + 1]);
        const float2 f = __half22float2(h[i]);
        l[i] = __floats2half2_rn(a[2 * i] - f.x, a[2 * i + 1] - f.y);
    }
    *reinterpret_cast<uint4 *>(hi_base + off) = *reinterpret_cast<const uint4 *>(h);
    *reinterpret_cast<uint4 *>(lo_base + off) = *reinterpret_cast<const uint4 *>(l);
}

__device__ __forceinline__ void load_split8(const char *hi_base, const char *lo_base, uint32_t off, float (&a)[8])
{
    const uint4 uh = *reinterpret_cast<const uint4 *>(hi_base + off);
    const uint4 ul = *reinterpret_cast<const uint4 *>(lo_base + off);
    const __half2 *h = reinterpret_cast<const __half2 *>(&uh);
    const __half2 *l = reinterpret_cast<const __half2 *>(&ul);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 fh = __half22float2(h[i]), fl = __half22float2(l[i]);
        a[2 * i] = fh.x + fl.x;
        a[2 * i + 1] = fh.y + fl.y;
    }
}

// One "split" GEMM: D (+)= A * B^T with both operands as (hi, lo) pairs; 3 MMAs per 16-deep K step.
// Descriptors are built once per kernel; a K step only adds (byte advance >> 4) to the start-address field.
struct Operand {
    uint64_t dhi, dlo;    // descriptors of the hi / lo halves at K step 0
    uint32_t kstep16;     // (byte advance per K=16 step) >> 4
};

__device__ __forceinline__ Operand make_operand(uint32_t hi, uint32_t lo, uint32_t lbo, uint32_t sbo, uint32_t kstep)
{
    return {make_desc(hi, lbo, sbo), make_desc(lo, lbo, sbo), kstep >> 4};
}

__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const Operand &A, const Operand &B, uint32_t idesc, int n_ksteps)
{
    uint64_t ah = A.dhi, al = A.dlo, bh = B.dhi, bl = B.dlo;
    umma(tmem_d, ah, bh, idesc, 0u);
    umma(tmem_d, ah, bl, idesc, 1u);
    umma(tmem_d, al, bh, idesc, 1u);
#pragma unroll 4
    for (int ks = 1; ks < n_ksteps; ++ks) {
        ah += A.kstep16; al += A.kstep16; bh += B.kstep16; bl += B.kstep16;
        umma(tmem_d, ah, bh, idesc, 1u);
        umma(tmem_d, ah, bl, idesc, 1u);
        umma(tmem_d, al, bh, idesc, 1u);
    }
}

// same, with the accumulate flag of the very first MMA given by the caller (accumulators that persist across tiles)
__device__ __forceinline__ void issue_gemm_acc(uint32_t tmem_d, const Operand &A, const Operand &B, uint32_t idesc, int n_ksteps,
                                               uint32_t accumulate_first)
{
    uint64_t ah = A.dhi, al = A.dlo, bh = B.dhi, bl = B.dlo;
    umma(tmem_d, ah, bh, idesc, accumulate_first);
    umma(tmem_d, ah, bl, idesc, 1u);
    umma(tmem_d, al, bh, idesc, 1u);
#pragma unroll 4
    for (int ks = 1; ks < n_ksteps; ++ks) {
        ah += A.kstep16; al += A.kstep16; bh += B.kstep16; bl += B.kstep16;
        umma(tmem_d, ah, bh, idesc, 1u);
        umma(tmem_d, ah, bl, idesc, 1u);
        umma(tmem_d, al, bh, idesc, 1u);
    }
}

// activation buffer [128 rows, C cols] (chunk c at c*2048, row r at r*16)
__device__ __forceinline__ Operand act_as_A_kmajor(uint32_t hi, uint32_t lo) { return make_operand(hi, lo, 2048u, 128u, 4096u); }
__device__ __forceinline__ Operand act_as_mnmajor(uint32_t hi, uint32_t lo) { return make_operand(hi, lo, 128u, 2048u, 256u); }
// weight buffer [64 out rows, Cin cols] (chunk c at c*1024, row o at o*16)
__device__ __forceinline__ Operand w_as_B_kmajor(uint32_t hi, uint32_t lo) { return make_operand(hi, lo, 1024u, 128u, 2048u); }
__device__ __forceinline__ Operand w_as_B_mnmajor(uint32_t hi, uint32_t lo) { return make_operand(hi, lo, 128u, 1024u, 256u); }

// Internal column order of the first layer's input: [in1 (n_in1) | in0 (n_in0) | constant one | zeros].  Putting the wide,
// 16-byte aligned in1 block (the hash-grid features) first lets the input rows be read, and their gradients be written,
// with float4 accesses.  `shift` = n_in1 for the first layer (global column = (c < n_in1) ? n_in0 + c : c - n_in1), 0 otherwise.
__device__ __forceinline__ int global_col(int c, int n_in0, int n_in1) { return c < n_in1 ? n_in0 + c : c - n_in1; }

// W[64][n_in] fp32 (global) -> split fp16 canonical K-major B operand with Kpad columns (zero padded)
__device__ __forceinline__ void stage_weight(char *smem, uint32_t hi_off, uint32_t lo_off, const float *__restrict__ Wg, int n_in,
                                             int Kpad, int n_in0 = 0, int n_in1 = 0)
{
    for (int i = threadIdx.x; i < W * (Kpad / 8); i += THREADS) {
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * c8 + j;
            a[j] = c < n_in ? __ldg(Wg + o * n_in + (n_in1 > 0 ? global_col(c, n_in0, n_in1) : c)) : 0.f;
        }
        store_split8(smem + hi_off, smem + lo_off, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
}

// exactly one lane issues the warp-uniform tcgen05 instructions; behind a plain `threadIdx.x == 0` test it wraps every
// UTCHMMA in an ELECT / BRA.U.ANY loop over the "possibly several" active lanes (6 instructions and a branch per MMA).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}


__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Power-of-two tile scale from the tile's gradient bound x > 0: x = m * 2^e with m in [0.5, 1) (frexp), e clamped to
// [-60, 60]; scale = 2^(6-e), inv_scale = 2^(e-6).  Exponent-field arithmetic: frexpf / ldexpf are library routines with
// denormal branches, and every thread of the CTA runs this once per tile.
__device__ __forceinline__ void pow2_scale(float x, float &scale, float &inv_scale)
{
    x = fmaxf(x, 1e-30f);                                         // normal number
    int e = ((__float_as_int(x) >> 23) & 0xff) - 126;
    e = max(min(e, 60), -60);
    scale = __int_as_float((127 + 6 - e) << 23);
    inv_scale = __int_as_float((127 - 6 + e) << 23);
}

// Sum 32 per-lane values over the 32 lanes of a warp, for 32 different quantities at once: butterfly that halves the
// number of values a lane carries at every step (16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5).  Returns, on lane
// L, the warp-wide total of element L.
__device__ __forceinline__ float warp_sum32(float (&v)[32], int lane)
{
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = upper ? v[i] : v[i + half];
            const float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}


// this thread's 16-column share of d(input): internal columns [16ci, 16ci+16) of the dX accumulator -> din1 / din0
__device__ __forceinline__ void write_dx16(const TcDims &D, int c0, const float (&v)[16], float inv_scale, int64_t row,
                                           float *__restrict__ din0, float *__restrict__ din1)
{
    if (din1 && (D.n_in1 & 3) == 0 && c0 + 16 <= D.n_in1) {
        float4 *dst = reinterpret_cast<float4 *>(din1 + row * D.n_in1 + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(v[4 * q] * inv_scale, v[4 * q + 1] * inv_scale, v[4 * q + 2] * inv_scale, v[4 * q + 3] * inv_scale);
        return;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int col = c0 + j;
        if (col < D.n_in1) {
            if (din1) din1[row * D.n_in1 + col] = v[j] * inv_scale;
        } else if (col < D.din) {
            if (din0) din0[row * D.n_in0 + (col - D.n_in1)] = v[j] * inv_scale * D.s0;
        }
    }
}


// SPEC = 1 pins the input shape of the SDF network (3 xyz + 32 hash features -> K0 = 48, two hidden layers) at compile
// time: the staging loops, vector-path tests and K loops of the dominant launches (18.7 M tap rows per step) become
// straight-line code.  SPEC = 0 reads everything from TcDims.
template <int SPEC>
__device__ __forceinline__ TcDims specialise(TcDims D)
{
    if (SPEC == 1) { D.n_in0 = 3; D.n_in1 = 32; D.din = 35; D.K0 = 48; D.nh = 2; }
    if (SPEC == 2) { D.n_in0 = 0; D.n_in1 = 87; D.din = 87; D.K0 = 96; D.nh = 2; }      // colour head: 65+3 features | SH(4) | normal
    return D;
}


inline int make_dims(const ia_mlp_desc *d, int32_t n_out_used, TcDims *D)
{
    IA_REQUIRE(d != nullptr, "mlp_tc: desc is NULL");
    IA_REQUIRE(d->width == W, "mlp_tc: width must be 64 (got %d)", d->width);
    IA_REQUIRE(d->n_hidden_layers == 1 || d->n_hidden_layers == 2, "mlp_tc: n_hidden_layers must be 1 or 2");
    IA_REQUIRE(d->n_in0 >= 0 && d->n_in0 <= 8 && d->n_in1 >= 0, "mlp_tc: bad input split");
    const int din = d->n_in0 + d->n_in1;
    IA_REQUIRE(din >= 1 && din <= 95, "mlp_tc: input width %d not in [1,95]", din);
    IA_REQUIRE(n_out_used >= 0 && n_out_used <= d->n_out, "mlp_tc: n_out_used out of range");
    IA_REQUIRE(d->hidden_act == IA_ACT_RELU || d->hidden_act == IA_ACT_SOFTPLUS100, "mlp_tc: unsupported hidden activation");
    if (d->out_act != IA_ACT_NONE) {
        ia_set_error("mlp_tc: fused output activation not supported (apply it on the caller side)");
        return IA_ERR_UNSUPPORTED;
    }
    D->n_in0 = d->n_in0; D->n_in1 = d->n_in1; D->din = din; D->K0 = (din + 1 + 15) / 16 * 16;
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
