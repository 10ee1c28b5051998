// Backward of the colour head (88 -> 64 -> 64 -> 3, ReLU) on the tensor cores with two tiles in flight per CTA -- the "duo"
// organisation of mlp_tc.cu (mlp_tc_bwd_duo_kernel) for a first layer that is too wide for it: two resident copies of an
// 88-column input tile, its hidden layer and the dZ buffer (116 KB per team) do not fit beside the weights.
//   * The input tile X and the first hidden layer H1 share one buffer R1 (12 chunks of 8 columns): H1 overwrites chunks 0..7 once
//     L0 = X W0^T has completed, and those eight chunks of X are RE-STAGED from global memory (loads issued under the dX GEMM)
//     for the very last product of the tile, dW0 += dZ1^T X.  Chunks 8..11 (24 input columns, the constant one and zero padding)
//     stay resident.  80 KB per team.
//   * Column 88 of X is the constant one: the first layer's bias rides in the GEMM (column 88 of the staged W0 holds b0), its
//     gradient comes out of dW0 for free.  H1 has no such column (that would need TMEM columns the team does not have:
//     96 for forward / dH / dX, 64 for dW1, 96 for dW0 = 256); db1 and the output layer's gradient dWl are reduced over the 32
//     rows of a warp with the transposing butterfly (warp_sum32) once per tile and then live in ONE register per lane.
// Everything else as in the duo kernel: teams of 256 threads (thread = row r, 32-column half), own TMEM half, own mbarriers and
// named barrier, dW0 / dW1 persistent in TMEM across the tiles of a team, launch-wide power-of-two gradient scale.
// Reference: VanillaMLP.forward / autograd of the colour network, models/network_utils.py:96-113 via models/texture.py:26-33.
#include "mlp_tc_device.cuh"

namespace {

constexpr int TEAM96 = 256;
constexpr int DIN96 = 88;              // input columns (the folded colour-head row: h (64) | pts (3) | SH (16) | normal (3) | pad (2))
constexpr uint32_t T96_D0 = 0, T96_W1 = 96, T96_W0 = 160, T96_TEAM_COLS = 256;

struct Duo96Plan {
    uint32_t r1_hi[2], r1_lo[2], dz_hi[2], dz_lo[2], w0_hi, w0_lo, w1_hi, w1_lo, wl, b1, red, mbar, tmem, ops, total;
};

__host__ __device__ inline Duo96Plan make_duo96_plan()
{
    Duo96Plan p;
    uint32_t o = 0;
    auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 127u) & ~127u; return r; };
    for (int t = 0; t < 2; ++t) {
        p.r1_hi[t] = take(12 * 2048); p.r1_lo[t] = take(12 * 2048);
        p.dz_hi[t] = take(8 * 2048); p.dz_lo[t] = take(8 * 2048);
    }
    p.w0_hi = take(12 * 1024); p.w0_lo = take(12 * 1024);
    p.w1_hi = take(8 * 1024); p.w1_lo = take(8 * 1024);
    p.wl = take(3 * W * 4);
    p.b1 = take(W * 4);
    p.red = take(64 * 4);
    p.mbar = take(64);
    p.tmem = take(16);
    p.ops = take(12 * sizeof(Operand));
    p.total = o;
    return p;
}

struct P96 { int pW0, pb0, pW1, pb1, pWl, pbl; };

__global__ void __launch_bounds__(2 * TEAM96, 1)
mlp_tc_bwd_duo96_kernel(const P96 Q, const float *__restrict__ in1, int64_t n, const float *__restrict__ params,
                        const float *__restrict__ dout, int64_t ld_dout, float *__restrict__ din1, float *__restrict__ dparams,
                        const float *__restrict__ gmax)
{
    constexpr int NOU = 3;
    extern __shared__ __align__(1024) char smem[];
    const Duo96Plan P = make_duo96_plan();
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, team = tid >> 8, tt = tid & (TEAM96 - 1), lane = tid & 31, wt = tt >> 5;
    const int q = wt & 3, half = wt >> 2, r = 32 * q + lane;
    float *red = reinterpret_cast<float *>(smem + P.red);
    // ---- one-time setup by all 512 threads
    for (int i = tid; i < W * 12; i += 2 * TEAM96) {          // [W0 | b0 | 0] -> split K-major B operand, 96 columns
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = 8 * c8 + j;
            a[j] = c < DIN96 ? __ldg(params + Q.pW0 + o * DIN96 + c) : (c == DIN96 ? __ldg(params + Q.pb0 + o) : 0.f);
        }
        store_split8(smem + P.w0_hi, smem + P.w0_lo, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
    for (int i = tid; i < W * 8; i += 2 * TEAM96) {
        const int o = i % W, c8 = i / W;
        float a[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __ldg(params + Q.pW1 + o * W + 8 * c8 + j);
        store_split8(smem + P.w1_hi, smem + P.w1_lo, (uint32_t)c8 * 1024u + (uint32_t)o * 16u, a);
    }
    float *wl_s = reinterpret_cast<float *>(smem + P.wl), *b1_s = reinterpret_cast<float *>(smem + P.b1);
    for (int i = tid; i < NOU * W; i += 2 * TEAM96) wl_s[i] = __ldg(params + Q.pWl + i);
    for (int i = tid; i < W; i += 2 * TEAM96) b1_s[i] = __ldg(params + Q.pb1 + i);
    if (tid < 8) mbar_init(sbase + P.mbar + 8u * (uint32_t)tid, 1);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(sbase + P.tmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    float wmax = 0.f;
    for (int i = tid; i < NOU * W; i += 2 * TEAM96) wmax = fmaxf(wmax, fabsf(__ldg(params + Q.pWl + i)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) red[tid >> 5] = wmax;
    Operand *ops = reinterpret_cast<Operand *>(smem + P.ops);       // [team][4] then 4 shared weight operands
    if (tt == 0) {
        Operand *o = ops + 4 * team;
        o[0] = act_as_A_kmajor(sbase + P.r1_hi[team], sbase + P.r1_lo[team]);
        o[1] = act_as_mnmajor(sbase + P.r1_hi[team], sbase + P.r1_lo[team]);
        o[2] = act_as_A_kmajor(sbase + P.dz_hi[team], sbase + P.dz_lo[team]);
        o[3] = act_as_mnmajor(sbase + P.dz_hi[team], sbase + P.dz_lo[team]);
    }
    if (tid == 0) {
        ops[8] = w_as_B_kmajor(sbase + P.w0_hi, sbase + P.w0_lo);
        ops[9] = w_as_B_kmajor(sbase + P.w1_hi, sbase + P.w1_lo);
        ops[10] = w_as_B_mnmajor(sbase + P.w0_hi, sbase + P.w0_lo);
        ops[11] = w_as_B_mnmajor(sbase + P.w1_hi, sbase + P.w1_lo);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(smem + P.tmem) + T96_TEAM_COLS * (uint32_t)team;
    const uint32_t lane_addr = ((uint32_t)q * 32u) << 16;
    wmax = 0.f;
#pragma unroll
    for (int w = 0; w < 2 * TEAM96 / 32; ++w) wmax = fmaxf(wmax, red[w]);
    wmax = fmaxf(wmax * (float)NOU, 1e-30f);
    float scale, inv_scale;
    pow2_scale(fmaxf(__ldg(gmax), 1e-30f) * wmax, scale, inv_scale);

    const Operand &AR1 = ops[4 * team + 0], &R1T = ops[4 * team + 1], &ADZ = ops[4 * team + 2], &DZT = ops[4 * team + 3];
    const Operand &BW0 = ops[8], &BW1 = ops[9], &BW0T = ops[10], &BW1T = ops[11];
    const uint32_t idesc_fwd = make_idesc(128, W, 0, 0), idesc_dh = make_idesc(128, W, 0, 1), idesc_dx = make_idesc(128, 96, 0, 1);
    const uint32_t idesc_dw1 = make_idesc(64, 64, 1, 1), idesc_dw0 = make_idesc(64, 96, 1, 1);
    const uint32_t mb0 = sbase + P.mbar + 32u * (uint32_t)team, mb1 = mb0 + 8u;
    uint32_t ph0 = 0, ph1 = 0;
    char *r1_hi = smem + P.r1_hi[team], *r1_lo = smem + P.r1_lo[team], *dz_hi = smem + P.dz_hi[team], *dz_lo = smem + P.dz_lo[team];

    auto team_issue = [&](auto issue) {
        fence_async_smem();
        tc_fence_before();
        asm volatile("bar.sync %0, %1;\n" ::"r"(1 + team), "n"(TEAM96) : "memory");
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

    // input rows: thread (r, half) owns the re-staged chunks half, half + 2, half + 4, half + 6 (columns < 64) and the resident
    // chunks 8, 10 (half 0) / 9, 11 (half 1; chunk 11 = constant one | zeros)
    float xr[4][8], xe[2][8];
    auto load8 = [&](float (&dst)[8], int64_t row, int chunk, bool valid) {
        const float4 *src = reinterpret_cast<const float4 *>(in1 + row * DIN96 + 8 * chunk);
        const float4 a = valid ? __ldg(src) : make_float4(0.f, 0.f, 0.f, 0.f), b = valid ? __ldg(src + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[0] = a.x; dst[1] = a.y; dst[2] = a.z; dst[3] = a.w; dst[4] = b.x; dst[5] = b.y; dst[6] = b.z; dst[7] = b.w;
    };
    auto load_main = [&](int64_t row, bool valid) {
#pragma unroll
        for (int k = 0; k < 4; ++k) load8(xr[k], row, half + 2 * k, valid);
    };
    auto load_extra = [&](int64_t row, bool valid) {
        load8(xe[0], row, 8 + half, valid);
        if (half == 0) load8(xe[1], row, 10, valid);
    };
    auto store_main = [&]() {
#pragma unroll
        for (int k = 0; k < 4; ++k) store_split8(r1_hi, r1_lo, (uint32_t)(half + 2 * k) * 2048u + (uint32_t)r * 16u, xr[k]);
    };
    auto store_extra = [&](bool valid) {
        store_split8(r1_hi, r1_lo, (uint32_t)(8 + half) * 2048u + (uint32_t)r * 16u, xe[0]);
        if (half == 0) {
            store_split8(r1_hi, r1_lo, 10u * 2048u + (uint32_t)r * 16u, xe[1]);
        } else {
            const float a[8] = {valid ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            store_split8(r1_hi, r1_lo, 11u * 2048u + (uint32_t)r * 16u, a);
        }
    };

    float gwl[NOU] = {0.f, 0.f, 0.f}, gbl[NOU] = {0.f, 0.f, 0.f}, gb1 = 0.f;
    const int64_t n_tiles = (n + ROWS - 1) / ROWS;
    const int64_t stride = (int64_t)gridDim.x * 2;
    int64_t tile = (int64_t)blockIdx.x * 2 + team;
    const bool had_tiles = tile < n_tiles;
    float dy_next[NOU] = {0.f, 0.f, 0.f};
    if (had_tiles) {
        const int64_t row = tile * ROWS + r;
        load_main(row, row < n);
        load_extra(row, row < n);
#pragma unroll
        for (int o = 0; o < NOU; ++o) dy_next[o] = row < n ? __ldg(dout + row * ld_dout + o) : 0.f;
    }
    uint32_t acc = 0u;
    for (; tile < n_tiles; tile += stride) {
        const int64_t row = tile * ROWS + r;
        const bool valid = row < n;
        float dy[NOU];
#pragma unroll
        for (int o = 0; o < NOU; ++o) dy[o] = dy_next[o];
        // ---- X(i) -> R1;  L0 = X [W0 | b0]^T
        store_main();
        store_extra(valid);
        team_issue([&]() { issue_gemm(tmem_base + T96_D0, AR1, BW0, idesc_fwd, 96 / 16); umma_commit(mb0); });
        wait0();
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {       // H1 = relu(L0) over chunks 0..7 of R1
            const int c0 = 32 * half + 16 * pass;
            float h[16];
            tmem_ld16(tmem_base + lane_addr + T96_D0 + (uint32_t)c0, h);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = fmaxf(h[8 * hf + j], 0.f);
                store_split8(r1_hi, r1_lo, (uint32_t)(c0 / 8 + hf) * 2048u + (uint32_t)r * 16u, a);
            }
        }
        // ---- L1 = H1 W1^T
        team_issue([&]() { issue_gemm(tmem_base + T96_D0, AR1, BW1, idesc_fwd, W / 16); umma_commit(mb0); });
        wait0();
        {
            float h2[32];
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int c0 = 32 * half + 16 * pass;
                float h[16], dz[16];
                tmem_ld16(tmem_base + lane_addr + T96_D0 + (uint32_t)c0, h);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int k = c0 + j;
                    h[j] = fmaxf(h[j] + b1_s[k], 0.f);
                    h2[16 * pass + j] = h[j];
                    float dh = 0.f;
#pragma unroll
                    for (int o = 0; o < NOU; ++o) dh = fmaf(dy[o], wl_s[o * W + k], dh);
                    dz[j] = h[j] > 0.f ? dh * scale : 0.f;
                }
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    float a[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[j] = dz[8 * hf + j];
                    store_split8(dz_hi, dz_lo, (uint32_t)(c0 / 8 + hf) * 2048u + (uint32_t)r * 16u, a);
                }
            }
            // output-layer and b1 gradients: column sums over the warp's 32 rows; lane L ends up with column 32 half + L
            float v[32];
#pragma unroll
            for (int o = 0; o < NOU; ++o) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = dy[o] * h2[j];
                gwl[o] += warp_sum32(v, lane);
                if (half == 0) gbl[o] += dy[o];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float dh = 0.f;
#pragma unroll
                for (int o = 0; o < NOU; ++o) dh = fmaf(dy[o], wl_s[o * W + 32 * half + j], dh);
                v[j] = h2[j] > 0.f ? dh * scale : 0.f;
            }
            gb1 += warp_sum32(v, lane);
        }
        // ---- dH1 = dZ2 W1 ; dW1 += dZ2^T H1
        team_issue([&]() {
            issue_gemm(tmem_base + T96_D0, ADZ, BW1T, idesc_dh, W / 16);
            umma_commit(mb0);
            issue_gemm_acc(tmem_base + T96_W1, DZT, R1T, idesc_dw1, ROWS / 16, acc);
            umma_commit(mb1);
        });
        wait0();
        float dz1[32];
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int c0 = 32 * half + 16 * pass;
            float v[16];
            tmem_ld16(tmem_base + lane_addr + T96_D0 + (uint32_t)c0, v);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                float hh[8];
                load_split8(r1_hi, r1_lo, (uint32_t)(c0 / 8 + hf) * 2048u + (uint32_t)r * 16u, hh);
#pragma unroll
                for (int j = 0; j < 8; ++j) dz1[16 * pass + 8 * hf + j] = hh[j] > 0.f ? v[8 * hf + j] : 0.f;
            }
        }
        wait1();          // dW1 has read dZ2 and H1
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            float a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = dz1[8 * c8 + j];
            store_split8(dz_hi, dz_lo, (uint32_t)(4 * half + c8) * 2048u + (uint32_t)r * 16u, a);
        }
        // ---- dX = dZ1 [W0 | b0 | 0]  (96 columns; 88 are written out);  X chunks 0..7 come back from global memory meanwhile
        team_issue([&]() { issue_gemm(tmem_base + T96_D0, ADZ, BW0T, idesc_dx, W / 16); umma_commit(mb0); });
        load_main(valid ? row : 0, valid);
        wait0();
        if (din1 != nullptr) {
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                const int c0 = 48 * half + 16 * part;
                if (c0 < DIN96) {
                    float v[16];
                    tmem_ld16(tmem_base + lane_addr + T96_D0 + (uint32_t)c0, v);
                    if (valid) {
                        float4 *dst = reinterpret_cast<float4 *>(din1 + row * DIN96 + c0);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (c0 + 4 * k < DIN96)
                                dst[k] = make_float4(v[4 * k] * inv_scale, v[4 * k + 1] * inv_scale, v[4 * k + 2] * inv_scale, v[4 * k + 3] * inv_scale);
                    }
                }
            }
        }
        store_main();
        // ---- dW0 (+db0) += dZ1^T [X | 1 | 0]
        team_issue([&]() { issue_gemm_acc(tmem_base + T96_W0, DZT, R1T, idesc_dw0, ROWS / 16, acc); umma_commit(mb1); });
        acc = 1u;
        {
            const int64_t nrow = (tile + stride) * ROWS + r;
            const bool nvalid = tile + stride < n_tiles && nrow < n;
            load_main(nvalid ? nrow : 0, nvalid);
            load_extra(nvalid ? nrow : 0, nvalid);
#pragma unroll
            for (int o = 0; o < NOU; ++o) dy_next[o] = nvalid ? __ldg(dout + nrow * ld_dout + o) : 0.f;
        }
        wait1();          // dW0 has read X and dZ1: both buffers are free for the next tile
    }
    // ---- drain
    if (had_tiles && dparams != nullptr) {
        const int o = 16 * q + lane;
        for (int ci = half; ci * 16 < 64; ci += 2) {
            float v[16];
            tmem_ld16(tmem_base + lane_addr + T96_W1 + 16u * (uint32_t)ci, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(dparams + Q.pW1 + o * W + 16 * ci + j, v[j] * inv_scale);
            }
        }
        for (int ci = half; ci * 16 < 96; ci += 2) {
            float v[16];
            tmem_ld16(tmem_base + lane_addr + T96_W0 + 16u * (uint32_t)ci, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int col = 16 * ci + j;
                    if (col < DIN96) atomicAdd(dparams + Q.pW0 + o * DIN96 + col, v[j] * inv_scale);
                    else if (col == DIN96) atomicAdd(dparams + Q.pb0 + o, v[j] * inv_scale);
                }
            }
        }
        atomicAdd(dparams + Q.pb1 + 32 * half + lane, gb1 * inv_scale);
#pragma unroll
        for (int oo = 0; oo < NOU; ++oo) {
            atomicAdd(dparams + Q.pWl + oo * W + 32 * half + lane, gwl[oo]);
            const float sb = warp_sum(gbl[oo]);
            if (lane == 0 && half == 0) atomicAdd(dparams + Q.pbl + oo, sb);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        const uint32_t base = *reinterpret_cast<volatile uint32_t *>(smem + P.tmem);
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(512u) : "memory");
    }
}

}  // namespace

// called by launch_bwd_tc (mlp_tc.cu) for the shape n_in0 = 0, n_in1 = 88, two hidden layers, ReLU, 3 outputs all used
int ia_tc_launch_bwd_duo96(int pW0, int pb0, int pW1, int pb1, int pWl, int pbl, const float *in1, int64_t n, const float *params,
                           const float *dout, int64_t ld_dout, float *din1, float *dparams, const float *gmax, cudaStream_t stream)
{
    const Duo96Plan P = make_duo96_plan();
    const int64_t n_tiles = ia_ceil_div(n, ROWS);
    const unsigned blocks = (unsigned)std::min<int64_t>(ia_ceil_div(n_tiles, 2), (int64_t)ia_sm_count());
    IA_CUDA_OK(cudaFuncSetAttribute(mlp_tc_bwd_duo96_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.total));
    const P96 Q{pW0, pb0, pW1, pb1, pWl, pbl};
    mlp_tc_bwd_duo96_kernel<<<blocks, 2 * TEAM96, P.total, stream>>>(Q, in1, n, params, dout, ld_dout, din1, dparams, gmax);
    IA_LAUNCH_OK("mlp_tc_bwd_duo96_kernel");
    return IA_OK;
}
