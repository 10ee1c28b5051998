// Width-64 fully fused MLP on the 5th-generation tensor cores (tcgen05 + TMEM), forward and backward.
// The "FullyFusedMLP" otype of reference models/network_utils.py:181-184 (tcnn.Network) and the fast path for
// VanillaMLP (models/network_utils.py:96-113) wherever its fp32 result can be reproduced to ~1e-6.
//
// Precision ("3xF16 split", fp32-equivalent): every GEMM operand is stored in shared memory as an fp16 pair
// hi = fp16(a), lo = fp16(a - hi) and each product is issued as three tcgen05.mma (hi*hi + hi*lo + lo*hi) into one
// fp32 TMEM accumulator, i.e. ~22 mantissa bits per operand.  This is REQUIRED by the workload, not a luxury: the
// finite-difference normals divide SDF differences by 2*eps ~ 3e-3, so a single fp16 pass (2^-11 relative) would put
// ~10 % noise on the eikonal term.  The MMA pipe is far from the bottleneck (the per-row activation epilogue is), so
// the 3x issue cost is hidden.  The last layer (<= 8 outputs: SDF taps, colours, densities) is evaluated in fp32
// registers.
//
// Mapping: one CTA = one 128-row tile = the 128 TMEM lanes; 512 threads = 4 "column groups" x 128 rows: thread
// (r, cg) owns accumulator columns [16cg, 16cg+16) of row r (tcgen05.ld 32x32b.x16), applies bias + activation and
// writes the next layer's A operand straight into the UMMA canonical K-major no-swizzle layout (core matrix = 8 rows
// x 16 B; address = chunk*2048 + row*16).  The same buffers are re-read as MN-major operands for the
// dW = dZ^T * H products (M=64, K=128 rows) of the backward pass.  Bias gradients come for free from a constant-one
// column appended to X / H1.  Parameter gradients accumulate in shared memory / registers across all tiles of the
// persistent CTA and are flushed with one atomic pass at the end.  Gradient operands are rescaled per tile by a
// power of two so that they sit in fp16's normal range.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "ia_common.cuh"
#include "hashgrid_device.cuh"
#include "mlp_tc_device.cuh"

namespace {


// optional cycle accounting (tools/prof_mlp.py --timing): [0] barrier wait, [1] MMA issue, [2] MMA completion wait,
// [3] whole tile, [4] tiles; accumulated by thread 0 of every CTA when enabled through ia_debug_tc_timing()
// Compiled out unless the library is built with -DIA_TC_TIMING=1: even a never-taken `if (g_tc_timing_on)` costs the hot
// loops a global load, registers and scheduling freedom (removing the sibling debug toggles was worth 4 % of the step).
#ifndef IA_TC_TIMING
#define IA_TC_TIMING 0
#endif
__device__ unsigned long long g_tc_cycles[8];
__device__ int g_tc_timing_flag = 0;
#define g_tc_timing_on (IA_TC_TIMING && g_tc_timing_flag)

struct Ctx {
    char *smem;
    SmemPlan P;
    uint32_t sbase;      // shared address of smem[0]
    uint32_t tmem;       // TMEM base address (lane 0, column 0)
    uint32_t lane_addr;  // this warp's lane-quarter offset in TMEM address format
    uint32_t phase;
    int r, cg;           // row of the tile / column group (16 accumulator columns) owned by this thread
};

// One lane of a CONVERGED warp (all callers sit right behind a CTA-wide barrier).  With elect.sync the compiler knows that
// all threads: make operand writes visible to the async proxy, sync, thread 0 issues via `issue`, everyone runs `overlap`
// (independent work: the next tile's input loads) and then waits
template <typename F, typename G>
__device__ __forceinline__ void run_mma_overlap(Ctx &c, F issue, G overlap)
{
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        if (elect_one()) {
            tc_fence_after();
            issue();
            umma_commit(c.sbase + c.P.mbar);
        }
        __syncwarp();
    }
    overlap();
    mbar_wait(c.sbase + c.P.mbar, c.phase & 1u);
    c.phase ^= 1u;
    tc_fence_after();
}

// all threads: make operand writes visible to the async proxy, sync, thread 0 issues via `issue`, everyone waits
template <typename F>
__device__ __forceinline__ void run_mma(Ctx &c, F issue)
{
    const bool timing = threadIdx.x == 0 && g_tc_timing_on;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (timing) t0 = clock64();
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        if (elect_one()) {
            if (timing) t1 = clock64();
            tc_fence_after();
            issue();
            umma_commit(c.sbase + c.P.mbar);
            if (timing) t2 = clock64();
        }
        __syncwarp();
    }
    mbar_wait(c.sbase + c.P.mbar, c.phase & 1u);
    c.phase ^= 1u;
    tc_fence_after();
    if (timing) {
        const long long t3 = clock64();
        atomicAdd(&g_tc_cycles[0], (unsigned long long)(t1 - t0));
        atomicAdd(&g_tc_cycles[1], (unsigned long long)(t2 - t1));
        atomicAdd(&g_tc_cycles[2], (unsigned long long)(t3 - t2));
    }
}

// Split form used by the software-pipelined backward: one CTA barrier publishes the operands, thread 0 issues several
// GEMM groups and commits each to its own mbarrier (mbar_commit(c, k)); the consumers wait per group (mma_wait(c, k)),
// so the epilogue of the first group runs while the tensor pipe is still busy with the later ones.
// Bit k of c.phase is the parity of mbarrier k.
template <typename F>
__device__ __forceinline__ void mma_issue(Ctx &c, F issue)
{
    const bool timing = threadIdx.x == 0 && g_tc_timing_on;
    const long long t0 = timing ? clock64() : 0;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t1 = timing ? clock64() : 0;
        tc_fence_after();
        issue();
        if (timing) {
            atomicAdd(&g_tc_cycles[0], (unsigned long long)(t1 - t0));
            atomicAdd(&g_tc_cycles[1], (unsigned long long)(clock64() - t1));
        }
    }
}
// Warp-specialised form (pipelined backward, WS = true): THREADS epilogue threads + one more warpgroup whose first warp only
// issues MMAs.  The epilogue threads publish their operand writes and meet the issuer warp on named barrier 1; one lane of
// the issuer warp then issues the GEMMs while every epilogue warp (including warp 0) is free to run ahead into its mbarrier
// waits.  Without it (WS = false) the issuing lane of warp 0 is blocked for ~3.7 k cycles per tile by the back-pressure of the
// tensor pipe's queue, and the other 15 warps wait for warp 0 at the next barrier (ncu: stall_barrier 15 %).
// Registers: a 17th warp alone capped ptxas at 96 registers for everyone (round 1: the spills cost more than the issuer
// bought).  Here the issuer's whole warpgroup drops to 32 registers with setmaxnreg.dec and the four epilogue warpgroups
// raise theirs to 112 with setmaxnreg.inc: 128 * 32 + 512 * 112 = 640 * 96, exactly the allocation of a 640-thread CTA.
constexpr int WS_EXTRA = 128;                   // setmaxnreg works on whole warpgroups
constexpr int WS_BAR = THREADS + 32;            // named barrier 1: the epilogue threads and the issuer warp
template <bool WS, typename F>
__device__ __forceinline__ void ws_publish(Ctx &c, F issue)
{
    const bool timing = threadIdx.x == 0 && g_tc_timing_on;
    const long long t0 = timing ? clock64() : 0;
    fence_async_smem();
    tc_fence_before();
    // WS: the epilogue threads only ARRIVE -- the issuer warp is the one that waits (ws_issue).  They meet no CTA-wide barrier
    // in the tile loop at all: everything they wait for afterwards is an mbarrier of an MMA group that the issuer commits
    // after this rendezvous, so a slow warp (or the issuer, blocked on the tensor pipe's queue) delays nobody but itself,
    // and the generations of the named barrier cannot mix.
    if (WS) asm volatile("bar.arrive 1, %0;\n" ::"n"(WS_BAR) : "memory");
    else asm volatile("bar.sync 1, %0;\n" ::"n"(THREADS) : "memory");
    if (timing) atomicAdd(&g_tc_cycles[0], (unsigned long long)(clock64() - t0));
    if (!WS && threadIdx.x < 32) {      // no issuer warp: one lane of warp 0 issues, then warp 0 joins the epilogue
        if (elect_one()) {
            const long long t1 = timing ? clock64() : 0;
            tc_fence_after();
            issue();
            if (timing) atomicAdd(&g_tc_cycles[1], (unsigned long long)(clock64() - t1));
        }
        __syncwarp();
    }
}
template <typename F>
__device__ __forceinline__ void ws_issue(Ctx &c, F issue)
{
    asm volatile("bar.sync 1, %0;\n" ::"n"(WS_BAR) : "memory");
    if (elect_one()) {
        const bool timing = g_tc_timing_on;
        const long long t1 = timing ? clock64() : 0;
        tc_fence_after();
        issue();
        if (timing) atomicAdd(&g_tc_cycles[1], (unsigned long long)(clock64() - t1));
    }
    __syncwarp();
}
__device__ __forceinline__ void mbar_commit(const Ctx &c, int k) { umma_commit(c.sbase + c.P.mbar + 8u * (uint32_t)k); }
__device__ __forceinline__ void mma_wait(Ctx &c, int k)
{
    const bool timing = threadIdx.x == 0 && g_tc_timing_on;
    const long long t0 = timing ? clock64() : 0;
    mbar_wait(c.sbase + c.P.mbar + 8u * (uint32_t)k, (c.phase >> k) & 1u);
    c.phase ^= 1u << k;
    tc_fence_after();
    if (timing) atomicAdd(&g_tc_cycles[5 + k], (unsigned long long)(clock64() - t0));   // [5..7]: wait per mbarrier
}

template <int NOU>
__device__ __forceinline__ void setup_common(Ctx &c, const TcDims &D, const float *__restrict__ params)
{
    const int tid = threadIdx.x;
    char *smem = c.smem;
    c.r = tid & (ROWS - 1);
    c.cg = tid >> 7;
    stage_weight(smem, c.P.w0_hi, c.P.w0_lo, params + D.pW0, D.din, D.K0, D.n_in0, D.n_in1);
    if (D.nh == 2) stage_weight(smem, c.P.w1_hi, c.P.w1_lo, params + D.pW1, W, W);
    float *wl = reinterpret_cast<float *>(smem + c.P.wl);
    for (int i = tid; i < NOU * W; i += THREADS) wl[i] = i < D.nou * W ? __ldg(params + D.pWl + i) : 0.f;
    float *b0 = reinterpret_cast<float *>(smem + c.P.b0), *b1 = reinterpret_cast<float *>(smem + c.P.b1);
    float *bl = reinterpret_cast<float *>(smem + c.P.bl);
    for (int i = tid; i < W; i += THREADS) {
        b0[i] = __ldg(params + D.pb0 + i);
        b1[i] = D.nh == 2 ? __ldg(params + D.pb1 + i) : 0.f;
    }
    if (tid < MAX_OUT) bl[tid] = tid < D.nou ? __ldg(params + D.pbl + tid) : 0.f;
    // constant-one column of the H1 buffer (chunk 8): hi = (1, 0, ..., 0), lo = 0  -> bias gradient of layer 1
    if (tid < ROWS) {
        float a[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_split8(smem + c.P.ah_hi, smem + c.P.ah_lo, 8u * 2048u + (uint32_t)tid * 16u, a);
    }
    if (tid == 0) {
        mbar_init(c.sbase + c.P.mbar, 1);
        mbar_init(c.sbase + c.P.mbar + 8, 1);
        mbar_init(c.sbase + c.P.mbar + 16, 1);
    }
    if (tid < 32) {  // warp 0 owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(c.sbase + c.P.tmem), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    c.tmem = *reinterpret_cast<volatile uint32_t *>(smem + c.P.tmem);
    c.lane_addr = ((uint32_t)((tid >> 5) & 3) * 32u) << 16;
    c.phase = 0;
}

__device__ __forceinline__ void teardown(Ctx &c)
{
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(c.tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ---- first-layer input rows.  Thread (r, cg) owns the 8-column chunks cg, cg+4, ... of row r.
constexpr int MAX_IN_CHUNKS = 3;   // K0 <= 96 -> at most 3 chunks per thread

struct InRegs { float v[MAX_IN_CHUNKS][8]; };

__device__ __forceinline__ void load_input_regs(const Ctx &c, const TcDims &D, const float *__restrict__ in0,
                                                const float *__restrict__ in1, int64_t row, bool valid, InRegs &R)
{
    const bool vec = (D.n_in1 & 3) == 0;
#pragma unroll
    for (int q = 0; q < MAX_IN_CHUNKS; ++q) {
        const int c8 = c.cg + CG * q;
        if (c8 * 8 >= D.K0) break;
        if (valid && vec && c8 * 8 + 8 <= D.n_in1) {
            const float4 *src = reinterpret_cast<const float4 *>(in1 + row * D.n_in1 + 8 * c8);
            const float4 a = __ldg(src), b = __ldg(src + 1);
            R.v[q][0] = a.x; R.v[q][1] = a.y; R.v[q][2] = a.z; R.v[q][3] = a.w;
            R.v[q][4] = b.x; R.v[q][5] = b.y; R.v[q][6] = b.z; R.v[q][7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = 8 * c8 + j;
                // in0 columns are kept RAW here (scale / offset are applied by store_input_regs): the caller prefetches these
                // registers one pipeline phase ahead, and arithmetic on a just-loaded value would stall the warp on it
                float v = 0.f;
                if (valid) {
                    if (col < D.n_in1) v = __ldg(in1 + row * D.n_in1 + col);
                    else if (col < D.din) v = __ldg(in0 + row * D.n_in0 + (col - D.n_in1));
                    else if (col == D.din) v = 1.0f;
                }
                R.v[q][j] = v;
            }
        }
    }
}

// RAW_IN0: R holds the in0 columns as loaded (load_input_regs); valid == false rows are all zeros and stay so
template <bool RAW_IN0 = true>
__device__ __forceinline__ void store_input_regs(Ctx &c, const TcDims &D, const InRegs &R, uint32_t hi_off, uint32_t lo_off,
                                                 bool valid = true)
{
#pragma unroll
    for (int q = 0; q < MAX_IN_CHUNKS; ++q) {
        const int c8 = c.cg + CG * q;
        if (c8 * 8 >= D.K0) break;
        if (RAW_IN0 && c8 * 8 + 8 > D.n_in1 && c8 * 8 < D.din) {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = 8 * c8 + j;
                a[j] = (valid && col >= D.n_in1 && col < D.din) ? fmaf(R.v[q][j], D.s0, D.o0) : R.v[q][j];
            }
            store_split8(c.smem + hi_off, c.smem + lo_off, (uint32_t)c8 * 2048u + (uint32_t)c.r * 16u, a);
        } else {
            store_split8(c.smem + hi_off, c.smem + lo_off, (uint32_t)c8 * 2048u + (uint32_t)c.r * 16u, R.v[q]);
        }
    }
}

__device__ __forceinline__ void stage_input(Ctx &c, const TcDims &D, const float *__restrict__ in0, const float *__restrict__ in1,
                                            int64_t row, bool valid)
{
    InRegs R;
    load_input_regs(c, D, in0, in1, row, valid, R);
    store_input_regs<true>(c, D, R, c.P.ax_hi, c.P.ax_lo, valid);
}

// ---- fused encoder: the first-layer input row is cat[hash-grid features (2 per level) | x*s0+o0 | 1 | 0], with the features
// gathered here instead of read from an [n, L*F] tensor (reference models/geometry.py:206, 233, 266: network(encoding(x))).
// Thread (r, cg) owns 8-column chunks cg, cg+4, ...: a feature chunk is 4 consecutive levels, so the warps of column group cg
// gather levels 4cg .. 4cg+3 (n_in1 is a multiple of 8 on this path).  x: [n,3] in the encoder's [0,1] coordinates.
__device__ __forceinline__ void gather_input_regs(const Ctx &c, const TcDims &D, const GridParams &G, const float2 *__restrict__ table,
                                                  const float *__restrict__ x, int64_t row_clamped, InRegs &R)
{
    const float px = __ldg(x + 3 * row_clamped), py = __ldg(x + 3 * row_clamped + 1), pz = __ldg(x + 3 * row_clamped + 2);
    const bool oob = point_outside(px, py, pz);
#pragma unroll
    for (int q = 0; q < MAX_IN_CHUNKS; ++q) {
        const int c8 = c.cg + CG * q;
        if (c8 * 8 >= D.K0) break;
        if (c8 * 8 < D.n_in1) {
#pragma unroll
            for (int lv = 0; lv < 4; ++lv) {
                const int l = 4 * c8 + lv;
                float f0 = 0.f, f1 = 0.f;
                if (l < G.active) {                                                      // masked levels: exact zeros
                    if (!oob) hg_gather_level<false>(G, table, l, px, py, pz, f0, f1);
                    else hg_gather_level<true>(G, table, l, px, py, pz, f0, f1);
                }
                R.v[q][2 * lv] = f0;
                R.v[q][2 * lv + 1] = f1;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = 8 * c8 + j - D.n_in1;       // 0..2: xyz, 3: the constant one, then zero padding
                R.v[q][j] = col == 0 ? fmaf(px, D.s0, D.o0) : col == 1 ? fmaf(py, D.s0, D.o0) : col == 2 ? fmaf(pz, D.s0, D.o0)
                                                                                               : col == 3 ? 1.0f : 0.f;
            }
        }
    }
}

// this thread's 16 accumulator columns of a hidden layer: h = act(z + b)
template <int ACT>
__device__ __forceinline__ void hidden_cols(Ctx &c, const float *__restrict__ bias, float (&h)[16])
{
    tmem_ld16(c.tmem + c.lane_addr + D0_COL + 16u * c.cg, h);
#pragma unroll
    for (int j = 0; j < 16; ++j) h[j] = act_fwd<ACT>(h[j] + bias[16 * c.cg + j]);
}

__device__ __forceinline__ void store_cols16(Ctx &c, uint32_t hi_off, uint32_t lo_off, const float (&v)[16])
{
    float a[8];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = v[8 * half + j];
        store_split8(c.smem + hi_off, c.smem + lo_off, (uint32_t)(2 * c.cg + half) * 2048u + (uint32_t)c.r * 16u, a);
    }
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
template <int ACT, int NOU, int SPEC = 0, bool FUSED = false>
__global__ void __launch_bounds__(THREADS, (NOU <= 3 ? 2 : 1))
mlp_tc_fwd_kernel(const TcDims Din, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
                  const float *__restrict__ params, float *__restrict__ out, int64_t ld_out, const GridParams G,
                  const float2 *__restrict__ table)
{
    const TcDims D = specialise<SPEC>(Din);
    extern __shared__ __align__(1024) char smem[];
    Ctx c;
    c.smem = smem;
    c.P = make_plan(D, false);
    c.sbase = smem_u32(smem);
    setup_common<NOU>(c, D, params);
    const float *b0 = reinterpret_cast<const float *>(smem + c.P.b0), *b1 = reinterpret_cast<const float *>(smem + c.P.b1);
    const float *bl = reinterpret_cast<const float *>(smem + c.P.bl), *wl = reinterpret_cast<const float *>(smem + c.P.wl);
    float *part = reinterpret_cast<float *>(smem + c.P.part);      // [CG][ROWS][NOU] partial output sums
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0);
    // operand descriptors live in shared memory: only the issuing thread needs them, keep them out of everyone's registers
    __shared__ Operand s_op[4];
    if (threadIdx.x == 0) {
        s_op[0] = act_as_A_kmajor(c.sbase + c.P.ax_hi, c.sbase + c.P.ax_lo);
        s_op[1] = act_as_A_kmajor(c.sbase + c.P.ah_hi, c.sbase + c.P.ah_lo);
        s_op[2] = w_as_B_kmajor(c.sbase + c.P.w0_hi, c.sbase + c.P.w0_lo);
        s_op[3] = w_as_B_kmajor(c.sbase + c.P.w1_hi, c.sbase + c.P.w1_lo);
    }
    const Operand &AX = s_op[0], &AH = s_op[1], &BW0 = s_op[2], &BW1 = s_op[3];
    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    InRegs Rn;                                             // this thread's input chunks of the CTA's NEXT tile (raw)
    if constexpr (!FUSED) {
        const int64_t r0 = (int64_t)blockIdx.x * ROWS + c.r;
        if ((int64_t)blockIdx.x < n_tiles) load_input_regs(c, D, in0, in1, r0, r0 < n, Rn);
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS + c.r;
        const bool valid = row < n;
        const long long tile_t0 = (threadIdx.x == 0 && g_tc_timing_on) ? clock64() : 0;
        if constexpr (FUSED) {
            InRegs R;
            gather_input_regs(c, D, G, table, in0, valid ? row : n - 1, R);      // rows past n: a copy of the last row, not written
            store_input_regs<false>(c, D, R, c.P.ax_hi, c.P.ax_lo);
        } else {
            store_input_regs<true>(c, D, Rn, c.P.ax_hi, c.P.ax_lo, valid);
        }
        if constexpr (!FUSED) {
            // the next tile's rows are requested behind this tile's first GEMM and consumed a whole tile later: the global-load
            // latency no longer sits at the head of every tile's dependent chain (2.39 -> 2.06 ms per 20 M tap rows)
            run_mma_overlap(c, [&]() { issue_gemm(c.tmem + D0_COL, AX, BW0, idesc_fwd, D.K0 / 16); }, [&]() {
                const int64_t nt = tile + gridDim.x, nr = nt * ROWS + c.r;
                if (nt < n_tiles) load_input_regs(c, D, in0, in1, nr, nr < n, Rn);
            });
        } else {
            run_mma(c, [&]() { issue_gemm(c.tmem + D0_COL, AX, BW0, idesc_fwd, D.K0 / 16); });
        }
        float h[16];
        hidden_cols<ACT>(c, b0, h);
        if (D.nh == 2) {
            store_cols16(c, c.P.ah_hi, c.P.ah_lo, h);
            run_mma(c, [&]() { issue_gemm(c.tmem + D0_COL, AH, BW1, idesc_fwd, W / 16); });
            hidden_cols<ACT>(c, b1, h);
        }
        if constexpr (NOU == 0) {
            // feature mode: hand the last hidden layer's activations to the caller (wide output layers are a plain GEMM)
            if (valid) {
                float4 *dst = reinterpret_cast<float4 *>(out + row * ld_out + 16 * c.cg);
#pragma unroll
                for (int q = 0; q < 4; ++q) dst[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
            }
        } else {
            // output layer in fp32: partial dot products over this thread's 16 hidden units, combined through smem
#pragma unroll
            for (int o = 0; o < NOU; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) acc = fmaf(h[j], wl[o * W + 16 * c.cg + j], acc);
                part[(c.cg * ROWS + c.r) * NOU + o] = acc;
            }
            __syncthreads();
            if (c.cg == 0 && valid) {
#pragma unroll
                for (int o = 0; o < NOU; ++o) {
                    if (o < D.nou) {
                        float acc = bl[o];
#pragma unroll
                        for (int g = 0; g < CG; ++g) acc += part[(g * ROWS + c.r) * NOU + o];
                        out[row * ld_out + o] = acc;
                    }
                }
            }
        }
        // `part` is written again only after the barriers of the next tile's run_mma
        if (threadIdx.x == 0 && g_tc_timing_on) {
            atomicAdd(&g_tc_cycles[3], (unsigned long long)(clock64() - tile_t0));
            atomicAdd(&g_tc_cycles[4], 1ull);
        }
    }
    teardown(c);
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
// dW_last[o][16 cg + j] += sum over the warp's 32 rows of dy[o] * h[j], two outputs (32 products) per butterfly
template <int NOU>
__device__ __forceinline__ void accumulate_dwl(float *__restrict__ dwl, const float (&dy)[NOU > 0 ? NOU : 1], const float (&h)[16],
                                               int cg, int lane)
{
#pragma unroll
    for (int o = 0; o < NOU; o += 2) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            v[j] = dy[o] * h[j];
            v[16 + j] = (o + 1 < NOU) ? dy[o + 1 < NOU ? o + 1 : o] * h[j] : 0.f;
        }
        const float tot = warp_sum32(v, lane);
        const int oo = o + (lane >> 4);
        if (oo < NOU) atomicAdd(&dwl[oo * W + 16 * cg + (lane & 15)], tot);
    }
}

// read this thread's share of a dW accumulator (M = 64 layout: output row o = 16*(lane quarter) + lane for lane < 16;
// 16-column chunk ci is handled by column group ci % CG) and add it, unscaled, to the padded smem accumulator
__device__ __forceinline__ void drain_dw(Ctx &c, float *__restrict__ acc_smem, int n_cols, int ld, float inv_scale)
{
    const int lane = threadIdx.x & 31;
    const int o = 16 * ((threadIdx.x >> 5) & 3) + lane;
    for (int ci = c.cg; ci * 16 < n_cols; ci += CG) {
        const int c0 = 16 * ci;
        float v[16];
        const int m = n_cols - c0 >= 16 ? 16 : 8;
        if (m == 16) tmem_ld16(c.tmem + c.lane_addr + D1_COL + (uint32_t)c0, v);
        else tmem_ld8(c.tmem + c.lane_addr + D1_COL + (uint32_t)c0, v);
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < m) acc_smem[o * ld + c0 + j] += v[j] * inv_scale;
        }
    }
}

template <int ACT, int NOU, int SPEC = 0>
__global__ void __launch_bounds__(THREADS, 1)
mlp_tc_bwd_kernel(const TcDims Din, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
                  const float *__restrict__ params, const float *__restrict__ dout, int64_t ld_dout,
                  float *__restrict__ din0, float *__restrict__ din1, float *__restrict__ dparams)
{
    const TcDims D = specialise<SPEC>(Din);
    extern __shared__ __align__(1024) char smem[];
    Ctx c;
    c.smem = smem;
    c.P = make_plan(D, true);
    c.sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld0 = D.K0 + 1, ld1 = 73;   // odd leading dimensions: conflict-free row-per-lane accumulation
    float *dw0 = reinterpret_cast<float *>(smem + c.P.dw0), *dw1 = reinterpret_cast<float *>(smem + c.P.dw1);
    float *dwl = reinterpret_cast<float *>(smem + c.P.dwl), *dbl = reinterpret_cast<float *>(smem + c.P.dbl);
    float *red = reinterpret_cast<float *>(smem + c.P.red);
    for (int i = tid; i < W * ld0; i += THREADS) dw0[i] = 0.f;
    for (int i = tid; i < W * ld1; i += THREADS) dw1[i] = 0.f;
    for (int i = tid; i < MAX_OUT * W; i += THREADS) dwl[i] = 0.f;
    if (tid < MAX_OUT) dbl[tid] = 0.f;
    setup_common<NOU>(c, D, params);
    const float *b0 = reinterpret_cast<const float *>(smem + c.P.b0), *b1 = reinterpret_cast<const float *>(smem + c.P.b1);
    const float *wl = reinterpret_cast<const float *>(smem + c.P.wl);
    // bound of sum_o |W_last[o][k]| for the per-tile gradient scale
    constexpr int NO = NOU > 0 ? NOU : 1;      // NOU == 0: feature mode, the incoming gradient is d(last hidden) [n, 64]
    float wmax = NOU == 0 ? 1.f : 0.f;
    for (int i = tid; i < NOU * W; i += THREADS) wmax = fmaxf(wmax, fabsf(wl[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[warp] = wmax;
    __syncthreads();
    wmax = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) wmax = fmaxf(wmax, red[w]);
    wmax = fmaxf(wmax * (float)NO, 1e-30f);
    __syncthreads();

    const bool want_dx = din0 != nullptr || din1 != nullptr;
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0);          // A K-major, B K-major
    const uint32_t idesc_dh = make_idesc(128, W, 0, 1);           // dH1 = dZ2 * W1      (B MN-major)
    const uint32_t idesc_dx = make_idesc(128, D.K0, 0, 1);        // dX  = dZ1 * W0      (B MN-major)
    const uint32_t idesc_dw1 = make_idesc(64, 72, 1, 1);          // dW1 = dZ2^T [H1|1]  (both MN-major)
    const uint32_t idesc_dw0 = make_idesc(64, D.K0, 1, 1);        // dW0 = dZ1^T [X|1]
    __shared__ Operand s_op[10];
    if (threadIdx.x == 0) {
        s_op[0] = act_as_A_kmajor(c.sbase + c.P.ax_hi, c.sbase + c.P.ax_lo);
        s_op[1] = act_as_A_kmajor(c.sbase + c.P.ah_hi, c.sbase + c.P.ah_lo);
        s_op[2] = act_as_A_kmajor(c.sbase + c.P.dz_hi, c.sbase + c.P.dz_lo);
        s_op[3] = act_as_mnmajor(c.sbase + c.P.ax_hi, c.sbase + c.P.ax_lo);
        s_op[4] = act_as_mnmajor(c.sbase + c.P.ah_hi, c.sbase + c.P.ah_lo);
        s_op[5] = act_as_mnmajor(c.sbase + c.P.dz_hi, c.sbase + c.P.dz_lo);
        s_op[6] = w_as_B_kmajor(c.sbase + c.P.w0_hi, c.sbase + c.P.w0_lo);
        s_op[7] = w_as_B_kmajor(c.sbase + c.P.w1_hi, c.sbase + c.P.w1_lo);
        s_op[8] = w_as_B_mnmajor(c.sbase + c.P.w0_hi, c.sbase + c.P.w0_lo);
        s_op[9] = w_as_B_mnmajor(c.sbase + c.P.w1_hi, c.sbase + c.P.w1_lo);
    }
    const Operand &AX = s_op[0], &AH = s_op[1], &ADZ = s_op[2], &XT = s_op[3], &HT = s_op[4], &DZT = s_op[5];
    const Operand &BW0 = s_op[6], &BW1 = s_op[7], &BW0T = s_op[8], &BW1T = s_op[9];

    // persistent per-thread partial sums of dW_last[o][16*cg + j] and db_last[o] (reduced over rows at the very end)
    constexpr bool REG_DWL = NOU >= 1 && NOU <= 3;
    constexpr int NREG = REG_DWL ? NOU : 1;
    float gwl[NREG][16];
    float gbl[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) gbl[o] = 0.f;
#pragma unroll
    for (int o = 0; o < NREG; ++o)
#pragma unroll
        for (int j = 0; j < 16; ++j) gwl[o][j] = 0.f;

    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * ROWS + c.r;
        const bool valid = row < n;
        const long long tile_t0 = (threadIdx.x == 0 && g_tc_timing_on) ? clock64() : 0;
        stage_input(c, D, in0, in1, row, valid);
        float dy[NO];
        float dhf[NOU == 0 ? 16 : 1];
        float dymax = 0.f;
        if constexpr (NOU == 0) {
            const float4 *src = reinterpret_cast<const float4 *>(dout + row * ld_dout + 16 * c.cg);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                dhf[4 * q] = t.x; dhf[4 * q + 1] = t.y; dhf[4 * q + 2] = t.z; dhf[4 * q + 3] = t.w;
                dymax = fmaxf(dymax, fmaxf(fmaxf(fabsf(t.x), fabsf(t.y)), fmaxf(fabsf(t.z), fabsf(t.w))));
            }
            dy[0] = 0.f;
        } else {
#pragma unroll
            for (int o = 0; o < NOU; ++o) {
                dy[o] = (valid && o < D.nou) ? __ldg(dout + row * ld_dout + o) : 0.f;
                dymax = fmaxf(dymax, fabsf(dy[o]));
                if (c.cg == 0) gbl[o] += dy[o];
            }
        }
        // per-tile power-of-two scale that bounds |dZ| of the tile by 2^6 (fp16 normal range, 2^10 headroom for dH)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dymax = fmaxf(dymax, __shfl_xor_sync(0xffffffffu, dymax, o));
        if (lane == 0) red[warp] = dymax;
        // ---- recompute forward
        run_mma(c, [&]() { issue_gemm(c.tmem + D0_COL, AX, BW0, idesc_fwd, D.K0 / 16); });   // (its barrier publishes `red`)
        dymax = 0.f;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) dymax = fmaxf(dymax, red[w]);
        float scale, inv_scale;
        pow2_scale(dymax * wmax, scale, inv_scale);
        float h[16];
        const float *bL = b0;
        if (D.nh == 2) {
            hidden_cols<ACT>(c, b0, h);
            store_cols16(c, c.P.ah_hi, c.P.ah_lo, h);
            run_mma(c, [&]() { issue_gemm(c.tmem + D0_COL, AH, BW1, idesc_fwd, W / 16); });
            bL = b1;
        }
        // ---- last hidden layer: h (recomputed), output-layer gradients in fp32 registers, dZ_last -> smem (scaled)
        hidden_cols<ACT>(c, bL, h);
        {
            float dz[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = 16 * c.cg + j;
                float dh = 0.f;
                if constexpr (NOU == 0) {
                    dh = dhf[j];
                } else {
#pragma unroll
                    for (int o = 0; o < NOU; ++o) dh = fmaf(dy[o], wl[o * W + k], dh);
                }
                dz[j] = dh * act_bwd_from_out<ACT>(h[j]) * scale;
            }
            if (NOU == 0) {
                // nothing: the output layer lives outside this kernel
            } else if (REG_DWL) {
#pragma unroll
                for (int o = 0; o < NREG; ++o)
#pragma unroll
                    for (int j = 0; j < 16; ++j) gwl[o][j] = fmaf(dy[o], h[j], gwl[o][j]);
            } else {
                accumulate_dwl<NOU>(dwl, dy, h, c.cg, lane);
            }
            store_cols16(c, c.P.dz_hi, c.P.dz_lo, dz);
        }
        if (D.nh == 2) {
            // ---- dH1 = dZ2 W1 ; dW1 (+db1) = dZ2^T [H1 | 1]
            run_mma(c, [&]() {
                issue_gemm(c.tmem + D0_COL, ADZ, BW1T, idesc_dh, W / 16);
                issue_gemm(c.tmem + D1_COL, DZT, HT, idesc_dw1, ROWS / 16);
            });
            drain_dw(c, dw1, 72, ld1, inv_scale);
            // dZ1 = dH1 (*) act'(H1), H1 re-read from its operand buffer; overwrites the dZ buffer (its MMAs are complete)
            float v[16];
            tmem_ld16(c.tmem + c.lane_addr + D0_COL + 16u * c.cg, v);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float hh[8], a[8];
                const uint32_t off = (uint32_t)(2 * c.cg + half) * 2048u + (uint32_t)c.r * 16u;
                load_split8(smem + c.P.ah_hi, smem + c.P.ah_lo, off, hh);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = v[8 * half + j] * act_bwd_from_out<ACT>(hh[j]);
                store_split8(smem + c.P.dz_hi, smem + c.P.dz_lo, off, a);
            }
        }
        // ---- dX = dZ1 W0 ; dW0 (+db0) = dZ1^T [X | 1]
        run_mma(c, [&]() {
            if (want_dx) issue_gemm(c.tmem + D0_COL, ADZ, BW0T, idesc_dx, W / 16);
            issue_gemm(c.tmem + D1_COL, DZT, XT, idesc_dw0, ROWS / 16);
        });
        drain_dw(c, dw0, D.K0, ld0, inv_scale);
        if (want_dx) {
            for (int ci = c.cg; ci * 16 < D.din; ci += CG) {
                const int c0 = 16 * ci;
                float v[16];
                tmem_ld16(c.tmem + c.lane_addr + D0_COL + (uint32_t)c0, v);
                if (valid) write_dx16(D, c0, v, inv_scale, row, din0, din1);
            }
        }
        if (threadIdx.x == 0 && g_tc_timing_on) {
            atomicAdd(&g_tc_cycles[3], (unsigned long long)(clock64() - tile_t0));
            atomicAdd(&g_tc_cycles[4], 1ull);
        }
    }
    // ---- reduce the register-resident output-layer gradients over the rows of the CTA
    if (REG_DWL) {
#pragma unroll
        for (int o = 0; o < NREG; ++o) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float s = warp_sum(gwl[o][j]);
                if (lane == 0) atomicAdd(&dwl[o * W + 16 * c.cg + j], s);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < NOU; ++o) {
        const float sb = warp_sum(gbl[o]);
        if (lane == 0 && c.cg == 0) atomicAdd(&dbl[o], sb);
    }
    __syncthreads();
    // ---- flush parameter gradients
    if (dparams != nullptr) {
        for (int i = tid; i < W * D.K0; i += THREADS) {
            const int o = i / D.K0, col = i - o * D.K0;
            const float g = dw0[o * ld0 + col];
            if (col < D.din) atomicAdd(dparams + D.pW0 + o * D.din + global_col(col, D.n_in0, D.n_in1), g);
            else if (col == D.din) atomicAdd(dparams + D.pb0 + o, g);
        }
        if (D.nh == 2) {
            for (int i = tid; i < W * 72; i += THREADS) {
                const int o = i / 72, col = i - o * 72;
                const float g = dw1[o * ld1 + col];
                if (col < W) atomicAdd(dparams + D.pW1 + o * W + col, g);
                else if (col == W) atomicAdd(dparams + D.pb1 + o, g);
            }
        }
        for (int i = tid; i < D.nou * W; i += THREADS) atomicAdd(dparams + D.pWl + i, dwl[i]);
        if (tid < D.nou) atomicAdd(dparams + D.pbl + tid, dbl[tid]);
    }
    teardown(c);
}


// ------------------------------------------------------------------------------------------------------------
// backward, software pipelined (two hidden layers, K0 <= 64): the forward recomputation of tile i+1 shares the two
// MMA phases of tile i's backward, so a tile costs 2 barrier/MMA round trips instead of 4 and every phase carries more
// independent epilogue work.
//   phase 1 MMAs:  dH1(i) = dZ2 W1 -> D2 | dW1 += dZ2^T [H1(i)|1] -> D1 | L0(i+1) = X(i+1) W0^T -> D0
//   phase 2 MMAs:  dX(i)  = dZ1 W0 -> D2 | dW0 += dZ1^T [X(i)|1]  -> D1 | L1(i+1) = H1(i+1) W1^T -> D0
// X / H1 are double buffered (tile parity), dZ is single buffered (each MMA phase has completed before it is rewritten).
// ------------------------------------------------------------------------------------------------------------
constexpr uint32_t PD0 = 0, PD2 = 64, PD1 = 128;   // TMEM columns: forward acc | dH/dX acc | dW acc

__device__ __forceinline__ void drain_dw_at(Ctx &c, uint32_t col0, float *__restrict__ acc_smem, int n_cols, int ld, float inv_scale)
{
    const int lane = threadIdx.x & 31;
    const int o = 16 * ((threadIdx.x >> 5) & 3) + lane;
    for (int ci = c.cg; ci * 16 < n_cols; ci += CG) {
        const int c0 = 16 * ci;
        float v[16];
        const int m = n_cols - c0 >= 16 ? 16 : 8;
        if (m == 16) tmem_ld16(c.tmem + c.lane_addr + col0 + (uint32_t)c0, v);
        else tmem_ld8(c.tmem + c.lane_addr + col0 + (uint32_t)c0, v);
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < m) acc_smem[o * ld + c0 + j] += v[j] * inv_scale;
        }
    }
}

template <int ACT, int NOU, int SPEC = 0, bool FUSED = false, bool WS = false>
__global__ void __launch_bounds__(THREADS + (WS ? WS_EXTRA : 0), 1)
mlp_tc_bwd_pipe_kernel(const TcDims Din, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
                       const float *__restrict__ params, const float *__restrict__ dout, int64_t ld_dout,
                       float *__restrict__ din0, float *__restrict__ din1, float *__restrict__ dparams, const GridParams G,
                       const float2 *__restrict__ table, const float *__restrict__ gmax)
{
    const TcDims D = specialise<SPEC>(Din);
    extern __shared__ __align__(1024) char smem[];
    Ctx c;
    c.smem = smem;
    c.P = make_plan(D, true, true);
    c.sbase = smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld0 = D.K0 + 1, ld1 = 73;
    float *dw0 = reinterpret_cast<float *>(smem + c.P.dw0), *dw1 = reinterpret_cast<float *>(smem + c.P.dw1);
    float *dwl = reinterpret_cast<float *>(smem + c.P.dwl), *dbl = reinterpret_cast<float *>(smem + c.P.dbl);
    float *red = reinterpret_cast<float *>(smem + c.P.red);      // [0,16): per-warp maxima of the current tile's |dy|
    for (int i = tid; i < W * ld0; i += THREADS) dw0[i] = 0.f;
    for (int i = tid; i < W * ld1; i += THREADS) dw1[i] = 0.f;
    for (int i = tid; i < MAX_OUT * W; i += THREADS) dwl[i] = 0.f;
    if (tid < MAX_OUT) dbl[tid] = 0.f;
    setup_common<NOU>(c, D, params);
    if (tid < ROWS) {   // constant-one column of the second H1 buffer
        float a[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        store_split8(smem + c.P.ah2_hi, smem + c.P.ah2_lo, 8u * 2048u + (uint32_t)tid * 16u, a);
    }
    const float *b0 = reinterpret_cast<const float *>(smem + c.P.b0), *b1 = reinterpret_cast<const float *>(smem + c.P.b1);
    const float *wl = reinterpret_cast<const float *>(smem + c.P.wl);
    constexpr int NO = NOU > 0 ? NOU : 1;
    float wmax = NOU == 0 ? 1.f : 0.f;
    for (int i = tid; i < NOU * W; i += THREADS) wmax = fmaxf(wmax, fabsf(wl[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[warp] = wmax;
    __syncthreads();
    wmax = 0.f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) wmax = fmaxf(wmax, red[w]);
    wmax = fmaxf(wmax * (float)NO, 1e-30f);
    __syncthreads();

    const bool want_dx = din0 != nullptr || din1 != nullptr;
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0);
    const uint32_t idesc_dh = make_idesc(128, W, 0, 1);
    const uint32_t idesc_dx = make_idesc(128, D.K0, 0, 1);
    const uint32_t idesc_dw1 = make_idesc(64, 72, 1, 1);
    const uint32_t idesc_dw0 = make_idesc(64, D.K0, 1, 1);
    // [parity][0: X as A (K-major), 1: H1 as A, 2: X^T (MN-major), 3: H1^T]; then dZ, dZ^T, W0, W1, W0^T, W1^T
    __shared__ Operand s_buf[2][4];
    __shared__ Operand s_op[6];
    if (threadIdx.x == 0) {
        const uint32_t axh[2] = {c.sbase + c.P.ax_hi, c.sbase + c.P.ax2_hi}, axl[2] = {c.sbase + c.P.ax_lo, c.sbase + c.P.ax2_lo};
        const uint32_t ahh[2] = {c.sbase + c.P.ah_hi, c.sbase + c.P.ah2_hi}, ahl[2] = {c.sbase + c.P.ah_lo, c.sbase + c.P.ah2_lo};
        for (int b = 0; b < 2; ++b) {
            s_buf[b][0] = act_as_A_kmajor(axh[b], axl[b]);
            s_buf[b][1] = act_as_A_kmajor(ahh[b], ahl[b]);
            s_buf[b][2] = act_as_mnmajor(axh[b], axl[b]);
            s_buf[b][3] = act_as_mnmajor(ahh[b], ahl[b]);
        }
        s_op[0] = act_as_A_kmajor(c.sbase + c.P.dz_hi, c.sbase + c.P.dz_lo);
        s_op[1] = act_as_mnmajor(c.sbase + c.P.dz_hi, c.sbase + c.P.dz_lo);
        s_op[2] = w_as_B_kmajor(c.sbase + c.P.w0_hi, c.sbase + c.P.w0_lo);
        s_op[3] = w_as_B_kmajor(c.sbase + c.P.w1_hi, c.sbase + c.P.w1_lo);
        s_op[4] = w_as_B_mnmajor(c.sbase + c.P.w0_hi, c.sbase + c.P.w0_lo);
        s_op[5] = w_as_B_mnmajor(c.sbase + c.P.w1_hi, c.sbase + c.P.w1_lo);
    }
    const Operand &ADZ = s_op[0], &DZT = s_op[1], &BW0 = s_op[2], &BW1 = s_op[3], &BW0T = s_op[4], &BW1T = s_op[5];
    const uint32_t ax_hi[2] = {c.P.ax_hi, c.P.ax2_hi}, ax_lo[2] = {c.P.ax_lo, c.P.ax2_lo};
    const uint32_t ah_hi[2] = {c.P.ah_hi, c.P.ah2_hi}, ah_lo[2] = {c.P.ah_lo, c.P.ah2_lo};

    constexpr bool REG_DWL = NOU >= 1 && NOU <= 3;
    constexpr int NREG = REG_DWL ? NOU : 1;
    float gwl[NREG][16];
    float gbl[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) gbl[o] = 0.f;
#pragma unroll
    for (int o = 0; o < NREG; ++o)
#pragma unroll
        for (int j = 0; j < 16; ++j) gwl[o][j] = 0.f;

    // One power-of-two gradient scale for the whole launch, from max |dout| (absmax pre-pass, *gmax) and the output layer's
    // weights: |dZ| <= 2^6 everywhere, fp16 normal range with 2^10 headroom for dH.  (A per-tile scale tracked each tile's own
    // maximum, but needed a CTA-wide exchange of the per-warp maxima in every tile; rows whose gradient is 2^-20 of the
    // launch maximum lose relative precision, their absolute error stays at 2^-24 of the maximum.)
    float scale, inv_scale;
    pow2_scale(fmaxf(__ldg(gmax), 1e-30f) * wmax, scale, inv_scale);
    // this thread's share of the incoming gradient of a tile: plain loads, consumed in stage A of that tile (a whole pipeline
    // phase later, so the warp never waits on them)
    float dy[NO];
    float dhf[NOU == 0 ? 16 : 1];
    auto load_dy = [&](int64_t row, bool valid) {
        if constexpr (NOU == 0) {
            const float4 *src = reinterpret_cast<const float4 *>(dout + row * ld_dout + 16 * c.cg);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = valid ? __ldg(src + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                dhf[4 * q] = t.x; dhf[4 * q + 1] = t.y; dhf[4 * q + 2] = t.z; dhf[4 * q + 3] = t.w;
            }
            dy[0] = 0.f;
        } else {
#pragma unroll
            for (int o = 0; o < NOU; ++o) dy[o] = (valid && o < D.nou) ? __ldg(dout + row * ld_dout + o) : 0.f;
        }
    };

    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    int64_t tile = blockIdx.x;
    // GEMM batches, in the order the tensor pipe executes them (issue order); each group commits to its own mbarrier
    auto issue_first0 = [&]() {
        issue_gemm(c.tmem + PD0, s_buf[0][0], BW0, idesc_fwd, D.K0 / 16);
        mbar_commit(c, 0);
    };
    auto issue_first1 = [&]() {
        issue_gemm(c.tmem + PD0, s_buf[0][1], BW1, idesc_fwd, W / 16);
        mbar_commit(c, 2);
    };
    // phase 1: the next tile's first forward GEMM goes first so that its epilogue (H1(i+1)) runs while the pipe works
    // through dH1 and the (longest) dW1 GEMM of tile i
    auto issue_phase1 = [&](int cur, int nxt, bool has_next) {
        if (has_next) issue_gemm(c.tmem + PD0, s_buf[nxt][0], BW0, idesc_fwd, D.K0 / 16);
        mbar_commit(c, 0);
        issue_gemm(c.tmem + PD2, ADZ, BW1T, idesc_dh, W / 16);
        issue_gemm(c.tmem + PD1, DZT, s_buf[cur][3], idesc_dw1, ROWS / 16);
        mbar_commit(c, 1);
    };
    // phase 2: dX first (its store streams out while dW0 and the next tile's second forward GEMM execute)
    auto issue_phase2 = [&](int cur, int nxt, bool has_next) {
        if (want_dx) issue_gemm(c.tmem + PD2, ADZ, BW0T, idesc_dx, W / 16);
        mbar_commit(c, 0);
        issue_gemm(c.tmem + PD1, DZT, s_buf[cur][2], idesc_dw0, ROWS / 16);
        mbar_commit(c, 1);
        if (has_next) issue_gemm(c.tmem + PD0, s_buf[nxt][1], BW1, idesc_fwd, W / 16);
        mbar_commit(c, 2);
    };
    if (WS && warp >= THREADS / 32) {
        // ---- the issuer's warpgroup (every warp of a warpgroup executes the same setmaxnreg, see ws_publish); its first warp
        // mirrors the tile loop of the epilogue threads, one named-barrier rendezvous per batch; the other three idle
        asm volatile("setmaxnreg.dec.sync.aligned.u32 32;\n");
        if (warp == THREADS / 32) {
        if (tile < n_tiles) {
            ws_issue(c, issue_first0);
            ws_issue(c, issue_first1);
        }
        for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
            const int cur = it & 1, nxt = cur ^ 1;
            const bool has_next = tile + gridDim.x < n_tiles;
            ws_issue(c, [&]() { issue_phase1(cur, nxt, has_next); });
            ws_issue(c, [&]() { issue_phase2(cur, nxt, has_next); });
        }
        }
    } else {
        if constexpr (WS) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;\n");
        if (tile < n_tiles) {
            // ---- prologue: forward of the first tile, its incoming gradient and scale
            const int64_t row = tile * ROWS + c.r;
            const bool valid = row < n;
            InRegs R0;
            if constexpr (FUSED) gather_input_regs(c, D, G, table, in0, valid ? row : n - 1, R0);
            else load_input_regs(c, D, in0, in1, row, valid, R0);
            store_input_regs<!FUSED>(c, D, R0, ax_hi[0], ax_lo[0], valid);
            load_dy(row, valid);
            ws_publish<WS>(c, issue_first0);
            mma_wait(c, 0);
            float h[16];
            tmem_ld16(c.tmem + c.lane_addr + PD0 + 16u * c.cg, h);
    #pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = act_fwd<ACT>(h[j] + b0[16 * c.cg + j]);
            store_cols16(c, ah_hi[0], ah_lo[0], h);
            ws_publish<WS>(c, issue_first1);
        }
        // input rows of the NEXT tile are fetched into registers one phase ahead of their use (global latency hidden)
        InRegs Rn;
        {
            const int64_t ntile0 = tile + gridDim.x;
            const int64_t nrow0 = ntile0 * ROWS + c.r;
            if (ntile0 < n_tiles) {
                if constexpr (FUSED) gather_input_regs(c, D, G, table, in0, nrow0 < n ? nrow0 : n - 1, Rn);
                else load_input_regs(c, D, in0, in1, nrow0, nrow0 < n, Rn);
            }
        }
        for (int it = 0; tile < n_tiles; ++it, tile += gridDim.x) {
            const int cur = it & 1, nxt = cur ^ 1;
            const int64_t row = tile * ROWS + c.r;
            const bool valid = row < n;
            const int64_t ntile = tile + gridDim.x;
            const bool has_next = ntile < n_tiles;
            const int64_t nrow = ntile * ROWS + c.r;
            const bool nvalid = has_next && nrow < n;
            const long long tile_t0 = (threadIdx.x == 0 && g_tc_timing_on) ? clock64() : 0;
            // ---- A: D0 = pre-activations of the last hidden layer of tile i; next tile's input goes to the other buffer
            if (has_next) store_input_regs<!FUSED>(c, D, Rn, ax_hi[nxt], ax_lo[nxt], nvalid);
            mma_wait(c, 2);      // second forward GEMM of tile i (issued by the prologue / the previous phase 2)
            {
                float h[16], dz[16];
                if constexpr (NOU > 0) {
                    if (c.cg == 0) {
#pragma unroll
                        for (int o = 0; o < NOU; ++o) gbl[o] += dy[o];
                    }
                }
                tmem_ld16(c.tmem + c.lane_addr + PD0 + 16u * c.cg, h);
    #pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = 16 * c.cg + j;
                    h[j] = act_fwd<ACT>(h[j] + b1[k]);
                    float dh = 0.f;
                    if constexpr (NOU == 0) {
                        dh = dhf[j];
                    } else {
    #pragma unroll
                        for (int o = 0; o < NOU; ++o) dh = fmaf(dy[o], wl[o * W + k], dh);
                    }
                    dz[j] = dh * act_bwd_from_out<ACT>(h[j]) * scale;
                }
                if (NOU == 0) {
                } else if (REG_DWL) {
    #pragma unroll
                    for (int o = 0; o < NREG; ++o)
    #pragma unroll
                        for (int j = 0; j < 16; ++j) gwl[o][j] = fmaf(dy[o], h[j], gwl[o][j]);
                } else {
                    accumulate_dwl<NOU>(dwl, dy, h, c.cg, lane);
                }
                store_cols16(c, c.P.dz_hi, c.P.dz_lo, dz);
            }
            // ---- phase 1
                ws_publish<WS>(c, [&]() { issue_phase1(cur, nxt, has_next); });
            const float inv_cur = inv_scale;
            // incoming gradient of the next tile (dy / dhf of tile i are dead from here on: the output-layer work of tile i
            // happened in stage A)
            if (has_next) load_dy(nrow, nvalid);
            mma_wait(c, 0);
            if (has_next) {
                // H1(i+1) = act(L0(i+1) + b0)
                float h[16];
                tmem_ld16(c.tmem + c.lane_addr + PD0 + 16u * c.cg, h);
    #pragma unroll
                for (int j = 0; j < 16; ++j) h[j] = act_fwd<ACT>(h[j] + b0[16 * c.cg + j]);
                store_cols16(c, ah_hi[nxt], ah_lo[nxt], h);
            }
            mma_wait(c, 1);      // dH1 in PD2, dW1 in PD1; the dZ buffer is free to be overwritten
            {
                // dZ1(i) = dH1 (*) act'(H1(i))
                float v[16];
                tmem_ld16(c.tmem + c.lane_addr + PD2 + 16u * c.cg, v);
    #pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float hh[8], a[8];
                    const uint32_t off = (uint32_t)(2 * c.cg + half) * 2048u + (uint32_t)c.r * 16u;
                    load_split8(smem + ah_hi[cur], smem + ah_lo[cur], off, hh);
    #pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = v[8 * half + j] * act_bwd_from_out<ACT>(hh[j]);
                    store_split8(smem + c.P.dz_hi, smem + c.P.dz_lo, off, a);
                }
            }
            drain_dw_at(c, PD1, dw1, 72, ld1, inv_cur);
            // ---- phase 2
            ws_publish<WS>(c, [&]() { issue_phase2(cur, nxt, has_next); });
            {
                const int64_t n2tile = ntile + gridDim.x;          // prefetch the input rows of tile i+2
                const int64_t n2row = n2tile * ROWS + c.r;
                if (n2tile < n_tiles) {
                    if constexpr (FUSED) gather_input_regs(c, D, G, table, in0, n2row < n ? n2row : n - 1, Rn);
                    else load_input_regs(c, D, in0, in1, n2row, n2row < n, Rn);
                }
            }
            mma_wait(c, 0);
            if (want_dx) {
                for (int ci = c.cg; ci * 16 < D.din; ci += CG) {
                    const int c0 = 16 * ci;
                    float v[16];
                    tmem_ld16(c.tmem + c.lane_addr + PD2 + (uint32_t)c0, v);
                    if (valid) write_dx16(D, c0, v, inv_cur, row, din0, din1);
                }
            }
            mma_wait(c, 1);
            drain_dw_at(c, PD1, dw0, D.K0, ld0, inv_cur);
            if (threadIdx.x == 0 && g_tc_timing_on) {
                atomicAdd(&g_tc_cycles[3], (unsigned long long)(clock64() - tile_t0));
                atomicAdd(&g_tc_cycles[4], 1ull);
            }
        }
    }
    // ---- reduce the register-resident output-layer gradients over the rows of the CTA
    if (REG_DWL && tid < THREADS) {
#pragma unroll
        for (int o = 0; o < NREG; ++o) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float s = warp_sum(gwl[o][j]);
                if (lane == 0) atomicAdd(&dwl[o * W + 16 * c.cg + j], s);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < NOU; ++o) {
        const float sb = warp_sum(gbl[o]);
        if (lane == 0 && c.cg == 0) atomicAdd(&dbl[o], sb);
    }
    __syncthreads();
    if (dparams != nullptr && tid < THREADS) {
        for (int i = tid; i < W * D.K0; i += THREADS) {
            const int o = i / D.K0, col = i - o * D.K0;
            const float g = dw0[o * ld0 + col];
            if (col < D.din) atomicAdd(dparams + D.pW0 + o * D.din + global_col(col, D.n_in0, D.n_in1), g);
            else if (col == D.din) atomicAdd(dparams + D.pb0 + o, g);
        }
        for (int i = tid; i < W * 72; i += THREADS) {
            const int o = i / 72, col = i - o * 72;
            const float g = dw1[o * ld1 + col];
            if (col < W) atomicAdd(dparams + D.pW1 + o * W + col, g);
            else if (col == W) atomicAdd(dparams + D.pb1 + o, g);
        }
        for (int i = tid; i < D.nou * W; i += THREADS) atomicAdd(dparams + D.pWl + i, dwl[i]);
        if (tid < D.nou) atomicAdd(dparams + D.pbl + tid, dbl[tid]);
    }
    teardown(c);
}

// ------------------------------------------------------------------------------------------------------------
// backward, two tiles in flight ("duo"): the 512 threads of the CTA are two TEAMS of 256; each team runs the plain
// recompute-forward / backward sequence on its own 128-row tile with its own operand buffers, TMEM columns, mbarriers and
// named barrier, and the two teams share only the staged weights.  Nothing orders one team against the other, so while
// one team sits in an epilogue (MUFU / FP32 pipes) or waits for its GEMMs, the other team's instructions and MMAs fill the
// issue slots and the tensor pipe -- the overlap two CTAs per SM would give, which the 64 K registers / 227 KB of an SM
// cannot hold for this kernel (ncu on the single-tile pipeline: tensor pipe 23 %, issue slots 38 %, stalls dominated by
// dependency waits of the 4 warps per scheduler).
//   thread (team, warp w, lane): TMEM lane quarter q = w % 4 (rows 32 q + lane), column half = w / 4 (32 columns).
//   dW0 / dW1 accumulate in TMEM across ALL tiles of the team (launch-wide gradient scale => no per-tile rescale, no drain):
//   TMEM columns per team: [0,64) forward / dH, [64,128) dX, [128,200) dW1, [200,248) dW0.
// Shape: the SDF network (3 + 32 inputs, K0 = 48, two hidden layers), NOU outputs used.
// ------------------------------------------------------------------------------------------------------------
constexpr int TEAM = 256;
constexpr uint32_t DUO_COLS = 512, DUO_TEAM_COLS = 256, DUO_D0 = 0, DUO_D2 = 64, DUO_W1 = 128, DUO_W0 = 200;

struct DuoPlan {
    uint32_t ax_hi[2], ax_lo[2], ah_hi[2], ah_lo[2], dz_hi[2], dz_lo[2], w0_hi, w0_lo, w1_hi, w1_lo, wl, b0, b1, bl, dwl, dbl, red,
        mbar, tmem, ops, total;
};

__host__ __device__ inline DuoPlan make_duo_plan()
{
    DuoPlan p;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
    for (int t = 0; t < 2; ++t) {
        p.ax_hi[t] = take(6 * 2048); p.ax_lo[t] = take(6 * 2048);
        p.ah_hi[t] = take(9 * 2048); p.ah_lo[t] = take(9 * 2048);
        p.dz_hi[t] = take(8 * 2048); p.dz_lo[t] = take(8 * 2048);
    }
    p.w0_hi = take(6 * 1024); p.w0_lo = take(6 * 1024);
    p.w1_hi = take(8 * 1024); p.w1_lo = take(8 * 1024);
    p.wl = take(MAX_OUT * W * 4);
    p.b0 = take(W * 4); p.b1 = take(W * 4); p.bl = take(MAX_OUT * 4);
    p.dwl = take(MAX_OUT * W * 4); p.dbl = take(MAX_OUT * 4);
    p.red = take(64 * 4);
    p.mbar = take(64);       // 2 teams x 4 mbarriers
    p.tmem = take(16);
    p.ops = take(2 * 6 * sizeof(Operand) + 4 * sizeof(Operand));
    p.total = o;
    return p;
}

template <int ACT, int NOU>
__global__ void __launch_bounds__(2 * TEAM, 1)
mlp_tc_bwd_duo_kernel(const TcDims Din, const float *__restrict__ in0, const float *__restrict__ in1, int64_t n,
                      const float *__restrict__ params, const float *__restrict__ dout, int64_t ld_dout,
                      float *__restrict__ din0, float *__restrict__ din1, float *__restrict__ dparams, const float *__restrict__ gmax)
{
    static_assert(NOU >= 0 && NOU <= 3, "duo backward: feature mode (0) or a fused output layer with 1..3 outputs");
    constexpr int NO = NOU > 0 ? NOU : 1;
    const TcDims D = specialise<1>(Din);
    extern __shared__ __align__(1024) char smem[];
    const DuoPlan P = make_duo_plan();
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, team = tid >> 8, tt = tid & (TEAM - 1), lane = tid & 31, wt = tt >> 5;
    const int q = wt & 3, half = wt >> 2, r = 32 * q + lane;
    float *dwl = reinterpret_cast<float *>(smem + P.dwl), *dbl = reinterpret_cast<float *>(smem + P.dbl);
    float *red = reinterpret_cast<float *>(smem + P.red);
    // ---- one-time setup by all 512 threads
    for (int i = tid; i < W * 6; i += 2 * TEAM) {          // W0 [64][35] -> split K-major B operand, 48 columns, internal order
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * c8 + j;
            a[j] = c < D.din ? __ldg(params + D.pW0 + o * D.din + global_col(c, D.n_in0, D.n_in1)) : 0.f;
        }
        store_split8(smem + P.w0_hi, smem + P.w0_lo, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
    for (int i = tid; i < W * 8; i += 2 * TEAM) {
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __ldg(params + D.pW1 + o * W + 8 * c8 + j);
        store_split8(smem + P.w1_hi, smem + P.w1_lo, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
    float *wl_s = reinterpret_cast<float *>(smem + P.wl);
    float *b0_s = reinterpret_cast<float *>(smem + P.b0), *b1_s = reinterpret_cast<float *>(smem + P.b1);
    for (int i = tid; i < NOU * W; i += 2 * TEAM) wl_s[i] = i < D.nou * W ? __ldg(params + D.pWl + i) : 0.f;
    for (int i = tid; i < W; i += 2 * TEAM) { b0_s[i] = __ldg(params + D.pb0 + i); b1_s[i] = __ldg(params + D.pb1 + i); }
    for (int i = tid; i < MAX_OUT * W; i += 2 * TEAM) dwl[i] = 0.f;
    if (tid < MAX_OUT) dbl[tid] = 0.f;
    {   // constant columns of each team's buffers: H1 chunk 8 = (1, 0, ..., 0) (bias gradient of layer 1), X chunk 5 = zeros
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (tt < ROWS) {
            store_split8(smem + P.ah_hi[team], smem + P.ah_lo[team], 8u * 2048u + (uint32_t)tt * 16u, one);
            store_split8(smem + P.ax_hi[team], smem + P.ax_lo[team], 5u * 2048u + (uint32_t)tt * 16u, zero);
        }
    }
    if (tid < 8) mbar_init(sbase + P.mbar + 8u * (uint32_t)tid, 1);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + P.tmem), "r"(DUO_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    float wmax = NOU == 0 ? 1.f : 0.f;      // NOU == 0: the incoming gradient is d(last hidden layer) itself
    for (int i = tid; i < NOU * W; i += 2 * TEAM) wmax = fmaxf(wmax, fabsf(__ldg(params + D.pWl + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[tid >> 5] = wmax;
    Operand *ops = reinterpret_cast<Operand *>(smem + P.ops);       // [team][6] then 4 shared weight operands
    if (tt == 0) {
        Operand *o = ops + 6 * team;
        o[0] = act_as_A_kmajor(sbase + P.ax_hi[team], sbase + P.ax_lo[team]);
        o[1] = act_as_A_kmajor(sbase + P.ah_hi[team], sbase + P.ah_lo[team]);
        o[2] = act_as_A_kmajor(sbase + P.dz_hi[team], sbase + P.dz_lo[team]);
        o[3] = act_as_mnmajor(sbase + P.ax_hi[team], sbase + P.ax_lo[team]);
        o[4] = act_as_mnmajor(sbase + P.ah_hi[team], sbase + P.ah_lo[team]);
        o[5] = act_as_mnmajor(sbase + P.dz_hi[team], sbase + P.dz_lo[team]);
    }
    if (tid == 0) {
        ops[12] = w_as_B_kmajor(sbase + P.w0_hi, sbase + P.w0_lo);
        ops[13] = w_as_B_kmajor(sbase + P.w1_hi, sbase + P.w1_lo);
        ops[14] = w_as_B_mnmajor(sbase + P.w0_hi, sbase + P.w0_lo);
        ops[15] = w_as_B_mnmajor(sbase + P.w1_hi, sbase + P.w1_lo);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem + P.tmem) + DUO_TEAM_COLS * (uint32_t)team;
    const uint32_t lane_addr = ((uint32_t)q * 32u) << 16;
    wmax = 0.f;
#pragma unroll
    for (int w = 0; w < 2 * TEAM / 32; ++w) wmax = fmaxf(wmax, red[w]);
    wmax = fmaxf(wmax * (float)NO, 1e-30f);
    float scale, inv_scale;
    pow2_scale(fmaxf(__ldg(gmax), 1e-30f) * wmax, scale, inv_scale);

    const Operand &AX = ops[6 * team + 0], &AH = ops[6 * team + 1], &ADZ = ops[6 * team + 2];
    const Operand &XT = ops[6 * team + 3], &HT = ops[6 * team + 4], &DZT = ops[6 * team + 5];
    const Operand &BW0 = ops[12], &BW1 = ops[13], &BW0T = ops[14], &BW1T = ops[15];
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0), idesc_dh = make_idesc(128, W, 0, 1), idesc_dx = make_idesc(128, 48, 0, 1);
    const uint32_t idesc_dw1 = make_idesc(64, 72, 1, 1), idesc_dw0 = make_idesc(64, 48, 1, 1);
    const uint32_t mb0 = sbase + P.mbar + 32u * (uint32_t)team, mb1 = mb0 + 8u;
    uint32_t ph0 = 0, ph1 = 0;
    char *ax_hi = smem + P.ax_hi[team], *ax_lo = smem + P.ax_lo[team], *ah_hi = smem + P.ah_hi[team], *ah_lo = smem + P.ah_lo[team];
    char *dz_hi = smem + P.dz_hi[team], *dz_lo = smem + P.dz_lo[team];

    // team barrier + MMA issue by one lane of the team's first warp
    auto team_issue = [&](auto issue) {
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync %0, %1;\n" ::"r"(1 + team), "n"(TEAM) : "memory");
        if (wt == 0) {
            if (elect_one()) {
                tc_fence_after();
                issue();
            }
            __syncwarp();
        }
    };
    auto wait0 = [&]() { mbar_wait(mb0, ph0); ph0 ^= 1u; tc_fence_after(); };
    auto wait1 = [&]() { mbar_wait(mb1, ph1); ph1 ^= 1u; tc_fence_after(); };

    // input rows: thread (r, half) owns chunks half, half + 2 (hash features) and, for half == 0, chunk 4 (xyz | 1 | 0 0 0 0)
    float xr[2][8], xyz[3];
    auto load_x = [&](int64_t row, bool valid) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 *src = reinterpret_cast<const float4 *>(in1 + row * 32 + 8 * (half + 2 * k));
            const float4 a = valid ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f), b = valid ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            xr[k][0] = a.x; xr[k][1] = a.y; xr[k][2] = a.z; xr[k][3] = a.w;
            xr[k][4] = b.x; xr[k][5] = b.y; xr[k][6] = b.z; xr[k][7] = b.w;
        }
        if (half == 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) xyz[j] = valid ? __ldg(in0 + row * 3 + j) : 0.f;
        }
    };
    auto store_x = [&](bool valid) {
#pragma unroll
        for (int k = 0; k < 2; ++k) store_split8(ax_hi, ax_lo, (uint32_t)(half + 2 * k) * 2048u + (uint32_t)r * 16u, xr[k]);
        if (half == 0) {
            const float a[8] = {valid ? fmaf(xyz[0], D.s0, D.o0) : 0.f, valid ? fmaf(xyz[1], D.s0, D.o0) : 0.f,
                                valid ? fmaf(xyz[2], D.s0, D.o0) : 0.f, valid ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f};
            store_split8(ax_hi, ax_lo, 4u * 2048u + (uint32_t)r * 16u, a);
        }
    };

    float gwl[NO][NOU > 0 ? 32 : 1], gbl[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        gbl[o] = 0.f;
#pragma unroll
        for (int j = 0; j < (NOU > 0 ? 32 : 1); ++j) gwl[o][j] = 0.f;
    }
    const bool want_dx = din0 != nullptr || din1 != nullptr;
    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    const int64_t stride = (int64_t)gridDim.x * 2;
    int64_t tile = (int64_t)blockIdx.x * 2 + team;
    const bool had_tiles = tile < n_tiles;
    float dy_next[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) dy_next[o] = 0.f;
    if (had_tiles) {
        const int64_t row = tile * ROWS + r;
        load_x(row, row < n);
#pragma unroll
        for (int o = 0; o < NOU; ++o) dy_next[o] = (row < n && o < D.nou) ? __ldg(dout + row * ld_dout + o) : 0.f;
    }
    uint32_t acc = 0u;        // 0 for the team's first tile: the persistent dW accumulators start from the first MMA
    for (; tile < n_tiles; tile += stride) {
        const int64_t row = tile * ROWS + r;
        const bool valid = row < n;
        float dy[NO];
#pragma unroll
        for (int o = 0; o < NO; ++o) dy[o] = dy_next[o];
        // ---- X(i) -> smem
        store_x(valid);
        // ---- L0 = X W0^T
        team_issue([&]() { issue_gemm(tmem_base + DUO_D0, AX, BW0, idesc_fwd, 48 / 16); umma_commit(mb0); });
        {
            // raw prefetch of tile i+1 (consumed one whole tile later).  Issued AFTER the team barrier: the proxy fence in
            // front of a barrier waits for outstanding loads (ncu: long-scoreboard stalls on FENCE.VIEW.ASYNC / BAR.SYNC)
            const int64_t nrow = (tile + stride) * ROWS + r;
            const bool nvalid = tile + stride < n_tiles && nrow < n;
            load_x(nvalid ? nrow : 0, nvalid);
#pragma unroll
            for (int o = 0; o < NOU; ++o) dy_next[o] = (nvalid && o < D.nou) ? __ldg(dout + nrow * ld_dout + o) : 0.f;
        }
        wait0();
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = 32 * half + 16 * pass;
            float h[16];
            tmem_ld16(tmem_base + lane_addr + DUO_D0 + (uint32_t)c0, h);
#pragma unroll
            for (int j = 0; j < 16; ++j) h[j] = act_fwd<ACT>(h[j] + b0_s[c0 + j]);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = h[8 * hf + j];
                store_split8(ah_hi, ah_lo, (uint32_t)(c0 / 8 + hf) * 2048u + (uint32_t)r * 16u, a);
            }
        }
        // ---- L1 = H1 W1^T
        team_issue([&]() { issue_gemm(tmem_base + DUO_D0, AH, BW1, idesc_fwd, W / 16); umma_commit(mb0); });
        wait0();
        if (NOU > 0 && half == 0) {
#pragma unroll
            for (int o = 0; o < NOU; ++o) gbl[o] += dy[o];
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = 32 * half + 16 * pass;
            float h[16], dz[16];
            float dhf[NOU == 0 ? 16 : 1];
            if constexpr (NOU == 0) {          // feature mode: d(last hidden layer) [n, 64] comes from the caller
                const float4 *src = reinterpret_cast<const float4 *>(dout + row * ld_dout + c0);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const float4 t4 = valid ? __ldg(src + qq) : make_float4(0.f, 0.f, 0.f, 0.f);
                    dhf[4 * qq] = t4.x; dhf[4 * qq + 1] = t4.y; dhf[4 * qq + 2] = t4.z; dhf[4 * qq + 3] = t4.w;
                }
            }
            tmem_ld16(tmem_base + lane_addr + DUO_D0 + (uint32_t)c0, h);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int k = c0 + j;
                h[j] = act_fwd<ACT>(h[j] + b1_s[k]);
                float dh = 0.f;
                if constexpr (NOU == 0) {
                    dh = dhf[j];
                } else {
#pragma unroll
                    for (int o = 0; o < NOU; ++o) {
                        dh = fmaf(dy[o], wl_s[o * W + k], dh);
                        gwl[o][16 * pass + j] = fmaf(dy[o], h[j], gwl[o][16 * pass + j]);
                    }
                }
                dz[j] = dh * act_bwd_from_out<ACT>(h[j]) * scale;
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = dz[8 * hf + j];
                store_split8(dz_hi, dz_lo, (uint32_t)(c0 / 8 + hf) * 2048u + (uint32_t)r * 16u, a);
            }
        }
        // ---- dH1 = dZ2 W1 ; dW1 (+db1) += dZ2^T [H1 | 1]   (persistent accumulator)
        team_issue([&]() {
            issue_gemm(tmem_base + DUO_D0, ADZ, BW1T, idesc_dh, W / 16);
            umma_commit(mb0);
            issue_gemm_acc(tmem_base + DUO_W1, DZT, HT, idesc_dw1, ROWS / 16, acc);
            umma_commit(mb1);
        });
        wait0();
        float dz1[32];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = 32 * half + 16 * pass;
            float v[16];
            tmem_ld16(tmem_base + lane_addr + DUO_D0 + (uint32_t)c0, v);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float hh[8];
                load_split8(ah_hi, ah_lo, (uint32_t)(c0 / 8 + hf) * 2048u + (uint32_t)r * 16u, hh);
#pragma unroll
                for (int j = 0; j < 8; ++j) dz1[16 * pass + 8 * hf + j] = v[8 * hf + j] * act_bwd_from_out<ACT>(hh[j]);
            }
        }
        wait1();          // dW1 has read dZ2 and H1: the dZ buffer may be overwritten
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = dz1[8 * c8 + j];
            store_split8(dz_hi, dz_lo, (uint32_t)(4 * half + c8) * 2048u + (uint32_t)r * 16u, a);
        }
        // ---- dX = dZ1 W0 ; dW0 (+db0) += dZ1^T [X | 1]
        team_issue([&]() {
            if (want_dx) issue_gemm(tmem_base + DUO_D2, ADZ, BW0T, idesc_dx, W / 16);
            umma_commit(mb0);
            issue_gemm_acc(tmem_base + DUO_W0, DZT, XT, idesc_dw0, ROWS / 16, acc);
            umma_commit(mb1);
        });
        acc = 1u;
        wait0();
        if (want_dx) {
            // internal columns [0,32): hash features -> din1 ; [32,35): xyz -> din0.  half 0: columns 0-15 and 32-47, half 1: 16-31
            float v[16];
            tmem_ld16(tmem_base + lane_addr + DUO_D2 + 16u * (uint32_t)half, v);
            if (valid) write_dx16(D, 16 * half, v, inv_scale, row, din0, din1);
            if (half == 0) {
                tmem_ld16(tmem_base + lane_addr + DUO_D2 + 32u, v);
                if (valid) write_dx16(D, 32, v, inv_scale, row, din0, din1);
            }
        }
        wait1();          // dW0 has read X and dZ1: both buffers are free for the next tile
    }
    // ---- per team: drain the persistent dW accumulators (M = 64 layout: output row o = 16 q + lane for lane < 16)
    if (had_tiles && dparams != nullptr) {
        const int o = 16 * q + lane;
        for (int ci = half; ci * 16 < 72; ci += 2) {
            const int c0 = 16 * ci, m = 72 - c0 >= 16 ? 16 : 8;
            float v[16];
            if (m == 16) tmem_ld16(tmem_base + lane_addr + DUO_W1 + (uint32_t)c0, v);
            else tmem_ld8(tmem_base + lane_addr + DUO_W1 + (uint32_t)c0, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = c0 + j;
                    if (j < m) {
                        if (col < W) atomicAdd(dparams + D.pW1 + o * W + col, v[j] * inv_scale);
                        else if (col == W) atomicAdd(dparams + D.pb1 + o, v[j] * inv_scale);
                    }
                }
            }
        }
        for (int ci = half; ci * 16 < 48; ci += 2) {
            const int c0 = 16 * ci;
            float v[16];
            tmem_ld16(tmem_base + lane_addr + DUO_W0 + (uint32_t)c0, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = c0 + j;
                    if (col < D.din) atomicAdd(dparams + D.pW0 + o * D.din + global_col(col, D.n_in0, D.n_in1), v[j] * inv_scale);
                    else if (col == D.din) atomicAdd(dparams + D.pb0 + o, v[j] * inv_scale);
                }
            }
        }
    }
    // ---- output-layer gradients: reduce the per-thread partial sums over the 32 rows of a warp, then over warps in smem
    if constexpr (NOU > 0) {
#pragma unroll
        for (int o = 0; o < NOU; ++o) {
            const float tot = warp_sum32(gwl[o], lane);          // lane L holds the warp total of column 32 half + L
            atomicAdd(&dwl[o * W + 32 * half + lane], tot);
            const float sb = warp_sum(gbl[o]);
            if (lane == 0 && half == 0) atomicAdd(&dbl[o], sb);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (dparams != nullptr) {
        for (int i = tid; i < D.nou * W; i += 2 * TEAM) atomicAdd(dparams + D.pWl + i, dwl[i]);
        if (tid < D.nou) atomicAdd(dparams + D.pbl + tid, dbl[tid]);
    }
    if (tid < 32) {
        const uint32_t base = *reinterpret_cast<volatile uint32_t *>(smem + P.tmem);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(DUO_COLS) : "memory");
    }
}

}  // namespace

// debug hook (not part of the public ABI header): enable/disable cycle accounting and read the counters back
extern "C" int32_t ia_debug_tc_timing(int32_t enable, unsigned long long *out8_host)
{
    if (out8_host) cudaMemcpyFromSymbol(out8_host, g_tc_cycles, sizeof(unsigned long long) * 8);
    unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_tc_cycles, zero, sizeof(zero));
    cudaMemcpyToSymbol(g_tc_timing_flag, &enable, sizeof(int));
    if (enable && !IA_TC_TIMING) {
        ia_set_error("ia_debug_tc_timing: library built without -DIA_TC_TIMING=1");
        return IA_ERR_UNSUPPORTED;
    }
    return 0;
}

int ia_tc_launch_bwd_duo96(int pW0, int pb0, int pW1, int pb1, int pWl, int pbl, const float *in1, int64_t n, const float *params,
                           const float *dout, int64_t ld_dout, float *din1, float *dparams, const float *gmax, cudaStream_t stream);
int ia_mlp_fwd_fp32(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, int32_t, float *, int64_t, void *);
int ia_mlp_bwd_fp32(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, const float *, int32_t, int64_t,
                    float *, float *, float *, void *);

namespace {

int launch_fwd_tc(const TcDims &D, const float *in0, const float *in1, int64_t n, const float *params, float *out, int64_t ld_out,
                  const GridParams *G, const float2 *table, void *stream)
{
    const SmemPlan P = make_plan(D, false);
    IA_REQUIRE(P.total <= 227 * 1024, "mlp_tc_fwd: needs %u B of shared memory", P.total);
    const int64_t n_tiles = ia_ceil_div(n, ROWS);
    const int per_sm = std::max(1, std::min(2, (int)((227u * 1024u) / (P.total + 1024u))));   // TMEM: 2 x 256 columns
    const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ia_sm_count() * per_sm);
    static const GridParams no_grid = {};
    const GridParams &GP = G ? *G : no_grid;
#define IA_TC_FWD(ACT, NOU, SPEC, FUSED)                                                                                          \
    do {                                                                                                                          \
        IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_fwd_kernel<ACT, NOU, SPEC, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                        (int)P.total));                                                                           \
        mlp_tc_fwd_kernel<ACT, NOU, SPEC, FUSED><<<blocks, THREADS, P.total, (cudaStream_t)stream>>>(D, in0, in1, n, params, out, \
                                                                                                     ld_out, GP, table);          \
    } while (0)
    const bool sp = D.act == IA_ACT_SOFTPLUS100;
    const bool geo = sp && D.n_in0 == 3 && D.n_in1 == 32 && D.nh == 2 && D.K0 == 48;     // the SDF network's input shape
    if (G != nullptr) {
        // fused encoder: the SDF network (Softplus) and the background density network (ReLU), SDF column / feature mode
        if (geo && D.nou == 0) IA_TC_FWD(IA_ACT_SOFTPLUS100, 0, 1, true);
        else if (geo && D.nou == 1) IA_TC_FWD(IA_ACT_SOFTPLUS100, 1, 1, true);
        else if (sp && D.nou == 0) IA_TC_FWD(IA_ACT_SOFTPLUS100, 0, 0, true);
        else if (sp && D.nou == 1) IA_TC_FWD(IA_ACT_SOFTPLUS100, 1, 0, true);
        else if (!sp && D.nou == 1) IA_TC_FWD(IA_ACT_RELU, 1, 0, true);
        else if (!sp && D.nou <= 8 && D.nou > 3) IA_TC_FWD(IA_ACT_RELU, 8, 0, true);
        else {
            ia_set_error("sdf_taps_fused_fwd: unsupported (activation, n_out_used) = (%d, %d)", D.act, D.nou);
            return IA_ERR_UNSUPPORTED;
        }
    }
    else if (geo && D.nou == 0) IA_TC_FWD(IA_ACT_SOFTPLUS100, 0, 1, false);
    else if (geo && D.nou == 1) IA_TC_FWD(IA_ACT_SOFTPLUS100, 1, 1, false);
    else if (D.nou == 0) { if (sp) IA_TC_FWD(IA_ACT_SOFTPLUS100, 0, 0, false); else IA_TC_FWD(IA_ACT_RELU, 0, 0, false); }
    else if (D.nou == 1) { if (sp) IA_TC_FWD(IA_ACT_SOFTPLUS100, 1, 0, false); else IA_TC_FWD(IA_ACT_RELU, 1, 0, false); }
    else if (D.nou <= 3) { if (sp) IA_TC_FWD(IA_ACT_SOFTPLUS100, 3, 0, false); else IA_TC_FWD(IA_ACT_RELU, 3, 0, false); }
    else { if (sp) IA_TC_FWD(IA_ACT_SOFTPLUS100, 8, 0, false); else IA_TC_FWD(IA_ACT_RELU, 8, 0, false); }
#undef IA_TC_FWD
    IA_LAUNCH_OK("mlp_tc_fwd_kernel");
    return IA_OK;
}

// max |dout| over the rows of a launch (device scalar, one of 64 rotating slots so that launches in flight on different
// streams do not share one): the pipelined backward's launch-wide gradient scale
__global__ void absmax_kernel(const float *__restrict__ v, int64_t n, int cols, int64_t ld, float *__restrict__ out)
{
    float m = 0.f;
    const int64_t total = n * cols, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    if (ld == cols && ((uintptr_t)v & 15) == 0) {           // contiguous rows: one flat array, 16-byte loads
        const float4 *v4 = reinterpret_cast<const float4 *>(v);
        const int64_t n4 = total >> 2;
        for (int64_t i = t0; i < n4; i += step) {
            const float4 a = __ldg(v4 + i);
            m = fmaxf(fmaxf(m, fmaxf(fabsf(a.x), fabsf(a.y))), fmaxf(fabsf(a.z), fabsf(a.w)));
        }
        for (int64_t i = (n4 << 2) + t0; i < total; i += step) m = fmaxf(m, fabsf(__ldg(v + i)));
    } else {
        for (int64_t i = t0; i < total; i += step) {
            const int64_t r = i / cols;
            m = fmaxf(m, fabsf(__ldg(v + r * ld + (i - r * cols))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int *>(out), __float_as_int(m));     // m >= 0: int order == float order
}

int absmax_slot(const float *dout, int64_t n, int cols, int64_t ld, cudaStream_t stream, float **slot_out)
{
    static float *slots = nullptr;
    static unsigned next = 0;
    if (slots == nullptr) IA_CUDA_OK(cudaMalloc(&slots, 64 * sizeof(float)));
    float *slot = slots + (next++ & 63u);
    IA_CUDA_OK(cudaMemsetAsync(slot, 0, sizeof(float), stream));
    const int64_t total = n * cols;
    const unsigned blocks = (unsigned)std::min<int64_t>(ia_ceil_div(total, 256 * 16), (int64_t)ia_sm_count() * 8);
    absmax_kernel<<<std::max(1u, blocks), 256, 0, stream>>>(dout, n, cols, ld, slot);
    IA_LAUNCH_OK("absmax_kernel");
    *slot_out = slot;
    return IA_OK;
}

int launch_bwd_tc(const TcDims &D, const float *in0, const float *in1, int64_t n, const float *params, const float *dout,
                  int64_t ld_dout, float *din0, float *din1, float *dparams, const GridParams *G, const float2 *table, void *stream)
{
    // software-pipelined variant when two hidden layers and the doubled X / H1 buffers fit in shared memory
    const SmemPlan Ppipe = make_plan(D, true, true);
    const bool pipe = D.nh == 2 && D.K0 <= 64 && Ppipe.total <= 227 * 1024 && (G != nullptr || getenv("IA_TC_NO_PIPE") == nullptr);
    const SmemPlan P = pipe ? Ppipe : make_plan(D, true);
    IA_REQUIRE(P.total <= 227 * 1024, "mlp_tc_bwd: needs %u B of shared memory", P.total);
    const int64_t n_tiles = ia_ceil_div(n, ROWS);
    const int per_sm = std::max(1, std::min(2, (int)((227u * 1024u) / (P.total + 1024u))));
    const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)ia_sm_count() * per_sm);
    static const GridParams no_grid = {};
    const GridParams &GP = G ? *G : no_grid;
    float *gmax = nullptr;
    if (pipe) {
        int rc = absmax_slot(dout, n, D.nou == 0 ? W : D.nou, ld_dout, (cudaStream_t)stream, &gmax);
        if (rc) return rc;
    }
#define IA_TC_BWD_PIPE_WS(ACT, NOU, SPEC, FUSED, WS)                                                                              \
    do {                                                                                                                          \
        IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_bwd_pipe_kernel<ACT, NOU, SPEC, FUSED, WS>,                                        \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.total));                              \
        mlp_tc_bwd_pipe_kernel<ACT, NOU, SPEC, FUSED, WS><<<blocks, THREADS + (WS ? WS_EXTRA : 0), P.total, (cudaStream_t)stream>>>( \
            D, in0, in1, n, params, dout, ld_dout, din0, din1, dparams, GP, table, gmax);                                         \
    } while (0)
    // the SDF-network shapes (SPEC = 1) have the warp-specialised variant; IA_TC_WS=0 selects the single-role kernel (A/B runs)
    // Measured on a B200 (profiles/r02_ab_tc_bwd_ws.md): 6.73 ms/step with the issuer warpgroup against 6.44 ms without, so the
    // single-role kernel stays the default; IA_TC_WS=1 selects the warp-specialised one.
    static const bool ws_env = getenv("IA_TC_WS") != nullptr && atoi(getenv("IA_TC_WS")) != 0;
#define IA_TC_BWD_PIPE(ACT, NOU, SPEC, FUSED)                                                                                     \
    do {                                                                                                                          \
        if (SPEC == 1 && ws_env) IA_TC_BWD_PIPE_WS(ACT, NOU, SPEC, FUSED, (SPEC == 1));                                           \
        else IA_TC_BWD_PIPE_WS(ACT, NOU, SPEC, FUSED, false);                                                                     \
    } while (0)
#define IA_TC_BWD_GEN(ACT, NOU, SPEC)                                                                                             \
    do {                                                                                                                          \
        IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_bwd_kernel<ACT, NOU, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                        (int)P.total));                                                                           \
        mlp_tc_bwd_kernel<ACT, NOU, SPEC><<<blocks, THREADS, P.total, (cudaStream_t)stream>>>(D, in0, in1, n, params, dout,      \
                                                                                              ld_dout, din0, din1, dparams);     \
    } while (0)
#define IA_TC_BWD(ACT, NOU)                                                                                                       \
    do {                                                                                                                          \
        if (pipe) IA_TC_BWD_PIPE(ACT, NOU, 0, false);                                                                             \
        else IA_TC_BWD_GEN(ACT, NOU, 0);                                                                                          \
    } while (0)
    const bool sp = D.act == IA_ACT_SOFTPLUS100;
    const bool geo = pipe && sp && D.n_in0 == 3 && D.n_in1 == 32 && D.nh == 2 && D.K0 == 48;
    const bool tex = !pipe && !sp && D.n_in0 == 0 && D.n_in1 == 87 && D.nh == 2 && D.K0 == 96 && D.nou == 3;
    // the tap evaluations of the SDF network (one output column): two tiles in flight per CTA (mlp_tc_bwd_duo_kernel);
    // IA_TC_DUO=0 selects the single-tile software pipeline (A/B runs)
    static const bool duo_env = getenv("IA_TC_DUO") == nullptr || atoi(getenv("IA_TC_DUO")) != 0;
    // the colour head after the output-layer fold (88 input columns, ReLU, 3 outputs): its own two-tile kernel (mlp_tc3.cu)
    if (G == nullptr && !sp && D.n_in0 == 0 && D.n_in1 == 88 && D.nh == 2 && D.nou == 3 && D.n_out == 3 && duo_env) {
        int rc = absmax_slot(dout, n, 3, ld_dout, (cudaStream_t)stream, &gmax);
        if (rc) return rc;
        return ia_tc_launch_bwd_duo96(D.pW0, D.pb0, D.pW1, D.pb1, D.pWl, D.pbl, in1, n, params, dout, ld_dout, din1, dparams, gmax,
                                      (cudaStream_t)stream);
    }
    if (G == nullptr && geo && (D.nou == 0 || D.nou == 1) && duo_env) {
        const DuoPlan DP = make_duo_plan();
        const unsigned dblocks = (unsigned)std::min<int64_t>(ia_ceil_div(n_tiles, 2), (int64_t)ia_sm_count());
        if (D.nou == 1) {
            IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_bwd_duo_kernel<IA_ACT_SOFTPLUS100, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP.total));
            mlp_tc_bwd_duo_kernel<IA_ACT_SOFTPLUS100, 1><<<dblocks, 2 * TEAM, DP.total, (cudaStream_t)stream>>>(
                D, in0, in1, n, params, dout, ld_dout, din0, din1, dparams, gmax);
        } else {
            IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_bwd_duo_kernel<IA_ACT_SOFTPLUS100, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DP.total));
            mlp_tc_bwd_duo_kernel<IA_ACT_SOFTPLUS100, 0><<<dblocks, 2 * TEAM, DP.total, (cudaStream_t)stream>>>(
                D, in0, in1, n, params, dout, ld_dout, din0, din1, dparams, gmax);
        }
        IA_LAUNCH_OK("mlp_tc_bwd_duo_kernel");
        return IA_OK;
    }
    if (G != nullptr) {
        // fused encoder (in-kernel re-gather of the first-layer input): two-hidden-layer Softplus networks, i.e. VolumeSDF
        if (!(pipe && sp && (D.nou == 0 || D.nou == 1))) {
            ia_set_error("sdf_taps_fused_bwd: unsupported network shape (hidden layers %d, activation %d, n_out_used %d)", D.nh, D.act, D.nou);
            return IA_ERR_UNSUPPORTED;
        }
        if (geo && D.nou == 0) IA_TC_BWD_PIPE(IA_ACT_SOFTPLUS100, 0, 1, true);
        else if (geo && D.nou == 1) IA_TC_BWD_PIPE(IA_ACT_SOFTPLUS100, 1, 1, true);
        else if (D.nou == 0) IA_TC_BWD_PIPE(IA_ACT_SOFTPLUS100, 0, 0, true);
        else IA_TC_BWD_PIPE(IA_ACT_SOFTPLUS100, 1, 0, true);
    }
    else if (geo && D.nou == 0) IA_TC_BWD_PIPE(IA_ACT_SOFTPLUS100, 0, 1, false);
    else if (geo && D.nou == 1) IA_TC_BWD_PIPE(IA_ACT_SOFTPLUS100, 1, 1, false);
    else if (tex) IA_TC_BWD_GEN(IA_ACT_RELU, 3, 2);
    else if (D.nou == 0) { if (sp) IA_TC_BWD(IA_ACT_SOFTPLUS100, 0); else IA_TC_BWD(IA_ACT_RELU, 0); }
    else if (D.nou == 1) { if (sp) IA_TC_BWD(IA_ACT_SOFTPLUS100, 1); else IA_TC_BWD(IA_ACT_RELU, 1); }
    else if (D.nou <= 3) { if (sp) IA_TC_BWD(IA_ACT_SOFTPLUS100, 3); else IA_TC_BWD(IA_ACT_RELU, 3); }
    else { if (sp) IA_TC_BWD(IA_ACT_SOFTPLUS100, 8); else IA_TC_BWD(IA_ACT_RELU, 8); }
#undef IA_TC_BWD
#undef IA_TC_BWD_GEN
#undef IA_TC_BWD_PIPE
#undef IA_TC_BWD_PIPE_WS
    IA_LAUNCH_OK("mlp_tc_bwd_kernel");
    return IA_OK;
}

}  // namespace

// abs-max pre-pass into a rotating device slot, for the kernels of mlp_tc2.cu
int ia_tc_absmax_slot(const float *v, int64_t n, int cols, int64_t ld, cudaStream_t stream, float **slot_out)
{
    return absmax_slot(v, n, cols, ld, stream, slot_out);
}

int ia_mlp_fwd_tc(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                  int32_t n_out_used, float *out, int64_t ld_out, void *stream)
{
    if (desc && n_out_used > MAX_OUT) {
        ia_set_error("mlp_tc: more than %d fused outputs; request n_out_used = 0 (last hidden layer) and apply the output "
                     "layer as a GEMM", MAX_OUT);
        return IA_ERR_UNSUPPORTED;
    }
    TcDims D;
    int rc = make_dims(desc, n_out_used, &D);
    if (rc) return rc;
    IA_REQUIRE(n >= 0 && (n == 0 || (params && out)), "mlp_tc_fwd: NULL pointer");
    IA_REQUIRE(n == 0 || ((D.n_in0 == 0 || in0) && (D.n_in1 == 0 || in1)), "mlp_tc_fwd: missing input pointer");
    IA_REQUIRE(ld_out >= (n_out_used == 0 ? W : n_out_used), "mlp_tc_fwd: ld_out too small");
    IA_REQUIRE(n_out_used != 0 || (ld_out % 4 == 0 && ((uintptr_t)out & 15) == 0), "mlp_tc_fwd: feature output must be 16-byte aligned");
    if (n == 0) return IA_OK;
    return launch_fwd_tc(D, in0, in1, n, params, out, ld_out, nullptr, nullptr, stream);
}

int ia_mlp_bwd_tc(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                  const float *dout, int32_t n_out_used, int64_t ld_dout, float *din0, float *din1, float *dparams,
                  void *stream)
{
    if (desc && n_out_used > MAX_OUT) {
        ia_set_error("mlp_tc: more than %d fused outputs; request n_out_used = 0 (last hidden layer)", MAX_OUT);
        return IA_ERR_UNSUPPORTED;
    }
    TcDims D;
    int rc = make_dims(desc, n_out_used, &D);
    if (rc) return rc;
    IA_REQUIRE(n >= 0 && (n == 0 || (params && dout)), "mlp_tc_bwd: NULL pointer");
    IA_REQUIRE(n == 0 || ((D.n_in0 == 0 || in0) && (D.n_in1 == 0 || in1)), "mlp_tc_bwd: missing input pointer");
    IA_REQUIRE(ld_dout >= (n_out_used == 0 ? W : n_out_used), "mlp_tc_bwd: ld_dout too small");
    IA_REQUIRE(n_out_used != 0 || (ld_dout % 4 == 0 && ((uintptr_t)dout & 15) == 0), "mlp_tc_bwd: feature gradient must be 16-byte aligned");
    if (n == 0) return IA_OK;
    return launch_bwd_tc(D, in0, in1, n, params, dout, ld_dout, din0, din1, dparams, nullptr, nullptr, stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused VolumeSDF evaluation (include/ia_b200.h: ia_sdf_taps_fused_fwd / _bwd)
// ---------------------------------------------------------------------------------------------------------------------
namespace {
int fused_dims(const ia_mlp_desc *desc, const ia_grid_plan *plan, int32_t active_levels, int32_t n_out_used, TcDims *D, GridParams *G)
{
    IA_REQUIRE(desc != nullptr && plan != nullptr, "sdf_taps_fused: NULL descriptor");
    IA_REQUIRE(desc->precision == IA_MLP_TC_F16, "sdf_taps_fused: the fused kernels are the tensor-core (IA_MLP_TC_F16) path");
    int rc = fill_params(plan, active_levels, G);
    if (rc) return rc;
    IA_REQUIRE(desc->n_in0 == 3, "sdf_taps_fused: the network input must be cat[xyz pass-through (3), encoding] (n_in0 = %d)", desc->n_in0);
    IA_REQUIRE(desc->n_in1 == plan->n_levels * plan->n_features, "sdf_taps_fused: n_in1 = %d but the grid has %d x %d features",
               desc->n_in1, plan->n_levels, plan->n_features);
    IA_REQUIRE(plan->n_levels % 4 == 0, "sdf_taps_fused: n_levels must be a multiple of 4 (got %d)", plan->n_levels);
    if (n_out_used > MAX_OUT) {
        ia_set_error("sdf_taps_fused: more than %d fused outputs; request n_out_used = 0 (last hidden layer)", MAX_OUT);
        return IA_ERR_UNSUPPORTED;
    }
    return make_dims(desc, n_out_used, D);
}
}  // namespace

extern "C" int32_t ia_sdf_taps_fused_fwd(const ia_mlp_desc *desc, const ia_grid_plan *plan, int32_t active_levels, const float *x,
                                         int64_t n, const float *table, const float *params, int32_t n_out_used, float *out,
                                         int64_t ld_out, void *stream)
{
    TcDims D;
    GridParams G;
    int rc = fused_dims(desc, plan, active_levels, n_out_used, &D, &G);
    if (rc) return rc;
    IA_REQUIRE(n >= 0 && (n == 0 || (x && table && params && out)), "sdf_taps_fused_fwd: NULL pointer");
    IA_REQUIRE(ld_out >= (n_out_used == 0 ? W : n_out_used), "sdf_taps_fused_fwd: ld_out too small");
    IA_REQUIRE(n_out_used != 0 || (ld_out % 4 == 0 && ((uintptr_t)out & 15) == 0), "sdf_taps_fused_fwd: feature output must be 16-byte aligned");
    if (n == 0) return IA_OK;
    return launch_fwd_tc(D, x, nullptr, n, params, out, ld_out, &G, reinterpret_cast<const float2 *>(table), stream);
}

extern "C" int32_t ia_hashgrid_bwd_grouped(const float *x, int64_t n, const float *table, const float *dy,
                                           const ia_grid_plan *plan, int32_t active_levels, int32_t group, float *dtable,
                                           float *dx, void *stream);

extern "C" int32_t ia_sdf_taps_fused_bwd(const ia_mlp_desc *desc, const ia_grid_plan *plan, int32_t active_levels, const float *x,
                                         int64_t n, const float *table, const float *params, const float *dout, int32_t n_out_used,
                                         int64_t ld_dout, int32_t group, float *dtable, float *dx_enc, float *dx_direct,
                                         float *dparams, float *denc_ws, void *stream)
{
    TcDims D;
    GridParams G;
    int rc = fused_dims(desc, plan, active_levels, n_out_used, &D, &G);
    if (rc) return rc;
    IA_REQUIRE(n >= 0 && (n == 0 || (x && table && params && dout)), "sdf_taps_fused_bwd: NULL pointer");
    IA_REQUIRE(ld_dout >= (n_out_used == 0 ? W : n_out_used), "sdf_taps_fused_bwd: ld_dout too small");
    IA_REQUIRE(n_out_used != 0 || (ld_dout % 4 == 0 && ((uintptr_t)dout & 15) == 0), "sdf_taps_fused_bwd: feature gradient must be 16-byte aligned");
    IA_REQUIRE((dtable == nullptr && dx_enc == nullptr) || denc_ws != nullptr,
               "sdf_taps_fused_bwd: table / position gradients need the [n, L*F] workspace");
    if (n == 0) return IA_OK;
    rc = launch_bwd_tc(D, x, nullptr, n, params, dout, ld_dout, dx_direct, denc_ws, dparams, &G, reinterpret_cast<const float2 *>(table), stream);
    if (rc) return rc;
    if (dtable != nullptr || dx_enc != nullptr)
        return ia_hashgrid_bwd_grouped(x, n, table, denc_ws, plan, active_levels, group, dtable, dx_enc, stream);
    return IA_OK;
}
