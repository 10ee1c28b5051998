// Flat effective parameters of a VanillaMLP in one launch (reference models/network_utils.py:115-134: nn.Linear layers,
// optionally wrapped in torch weight_norm: W = g * v / ||v||_row).  The fused MLP kernels take the ABI's flat layout
//   W0[out0, in0] b0[out0] W1[...] b1[...] ... ;
// assembling it with torch operators costs ~13 tiny kernels per network forward and ~30 in backward, five times per
// training step.  One warp per output row; the backward applies the weight-norm adjoint
//   dg = dW . v_hat,  dv = (g / ||v||) (dW - (dW . v_hat) v_hat),  db = d(flat bias).
#include "ia_common.cuh"

namespace {

struct WnParams {
    int n_layers;
    int n_out[IA_WN_MAX_LAYERS], n_in[IA_WN_MAX_LAYERS];
    int row0[IA_WN_MAX_LAYERS + 1];           // first global row of each layer
    long long off[IA_WN_MAX_LAYERS];          // offset of the layer's W block in the flat vector (its bias follows the block)
    const float *g[IA_WN_MAX_LAYERS], *v[IA_WN_MAX_LAYERS], *b[IA_WN_MAX_LAYERS];
    float *dg[IA_WN_MAX_LAYERS], *dv[IA_WN_MAX_LAYERS], *db[IA_WN_MAX_LAYERS];
};

__device__ __forceinline__ float warp_sum(float x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// ACC (backward only): add to dg / dv / db instead of overwriting them -- the destinations are the parameters' slices of the
// gradient arena (zeroed once per step), so no autograd accumulation pass follows.  Each element is owned by one warp.
template <bool BWD, bool ACC>
__global__ void weightnorm_kernel(const WnParams P, float *__restrict__ flat, const float *__restrict__ dflat)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= P.row0[P.n_layers]) return;
    int L = 0;
#pragma unroll
    for (int i = 1; i < IA_WN_MAX_LAYERS; ++i)
        if (i < P.n_layers && row >= P.row0[i]) L = i;
    const int o = row - P.row0[L], n_in = P.n_in[L], n_out = P.n_out[L];
    const float *__restrict__ v = P.v[L] + (long long)o * n_in;
    const long long wofs = P.off[L] + (long long)o * n_in, bofs = P.off[L] + (long long)n_out * n_in + o;
    const bool wn = P.g[L] != nullptr;
    float ss = 0.f;
    if (wn) {
        for (int k = lane; k < n_in; k += 32) { const float t = v[k]; ss = fmaf(t, t, ss); }
        ss = warp_sum(ss);
    }
    const float nrm = wn ? sqrtf(ss) : 1.f, gv = wn ? P.g[L][o] : 1.f;
    const float s = wn ? gv / nrm : 1.f;
    if (!BWD) {
        for (int k = lane; k < n_in; k += 32) flat[wofs + k] = v[k] * s;
        if (lane == 0) flat[bofs] = P.b[L][o];
    } else {
        const float *__restrict__ dw = dflat + wofs;
        float dot = 0.f;
        if (wn) {
            for (int k = lane; k < n_in; k += 32) dot = fmaf(dw[k], v[k], dot);
            dot = warp_sum(dot) / nrm;                                   // dW . v_hat
        }
        if (P.dv[L] != nullptr) {
            float *__restrict__ dv = P.dv[L] + (long long)o * n_in;
            for (int k = lane; k < n_in; k += 32) {
                const float t = wn ? s * (dw[k] - dot * v[k] / nrm) : dw[k];
                dv[k] = ACC ? dv[k] + t : t;
            }
        }
        if (lane == 0) {
            if (wn && P.dg[L] != nullptr) P.dg[L][o] = ACC ? P.dg[L][o] + dot : dot;
            if (P.db[L] != nullptr) P.db[L][o] = ACC ? P.db[L][o] + dflat[bofs] : dflat[bofs];
        }
    }
}

int fill(const ia_wn_desc *d, WnParams *P, bool bwd)
{
    IA_REQUIRE(d != nullptr, "weightnorm: desc is NULL");
    IA_REQUIRE(d->n_layers >= 1 && d->n_layers <= IA_WN_MAX_LAYERS, "weightnorm: n_layers %d not in [1,%d]", d->n_layers, IA_WN_MAX_LAYERS);
    P->n_layers = d->n_layers;
    long long off = 0;
    int row = 0;
    for (int i = 0; i < IA_WN_MAX_LAYERS; ++i) {
        const bool live = i < d->n_layers;
        P->n_out[i] = live ? d->n_out[i] : 0;
        P->n_in[i] = live ? d->n_in[i] : 0;
        P->row0[i] = row;
        P->off[i] = off;
        P->g[i] = live ? d->g[i] : nullptr;
        P->v[i] = live ? d->v[i] : nullptr;
        P->b[i] = live ? d->b[i] : nullptr;
        P->dg[i] = live && bwd ? d->dg[i] : nullptr;
        P->dv[i] = live && bwd ? d->dv[i] : nullptr;
        P->db[i] = live && bwd ? d->db[i] : nullptr;
        if (live) {
            IA_REQUIRE(d->n_out[i] >= 1 && d->n_in[i] >= 1 && d->v[i] && d->b[i], "weightnorm: bad layer %d", i);
            row += d->n_out[i];
            off += (long long)d->n_out[i] * d->n_in[i] + d->n_out[i];
        }
    }
    P->row0[IA_WN_MAX_LAYERS] = row;
    for (int i = d->n_layers; i < IA_WN_MAX_LAYERS; ++i) P->row0[i] = row;
    return IA_OK;
}

}  // namespace

extern "C" int32_t ia_weightnorm_flat_fwd(const ia_wn_desc *desc, float *flat, void *stream)
{
    WnParams P;
    int rc = fill(desc, &P, false);
    if (rc) return rc;
    IA_REQUIRE(flat != nullptr, "weightnorm_flat_fwd: flat is NULL");
    const int rows = P.row0[IA_WN_MAX_LAYERS];
    weightnorm_kernel<false, false><<<(unsigned)ia_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(P, flat, nullptr);
    IA_LAUNCH_OK("weightnorm_kernel<fwd>");
    return IA_OK;
}

extern "C" int32_t ia_weightnorm_flat_bwd(const ia_wn_desc *desc, const float *dflat, void *stream)
{
    WnParams P;
    int rc = fill(desc, &P, true);
    if (rc) return rc;
    IA_REQUIRE(dflat != nullptr, "weightnorm_flat_bwd: dflat is NULL");
    const int rows = P.row0[IA_WN_MAX_LAYERS];
    weightnorm_kernel<true, false><<<(unsigned)ia_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(P, nullptr, dflat);
    IA_LAUNCH_OK("weightnorm_kernel<bwd>");
    return IA_OK;
}

extern "C" int32_t ia_weightnorm_flat_bwd_acc(const ia_wn_desc *desc, const float *dflat, void *stream)
{
    WnParams P;
    int rc = fill(desc, &P, true);
    if (rc) return rc;
    IA_REQUIRE(dflat != nullptr, "weightnorm_flat_bwd_acc: dflat is NULL");
    const int rows = P.row0[IA_WN_MAX_LAYERS];
    weightnorm_kernel<true, true><<<(unsigned)ia_ceil_div(rows, 8), 256, 0, (cudaStream_t)stream>>>(P, nullptr, dflat);
    IA_LAUNCH_OK("weightnorm_kernel<bwd, acc>");
    return IA_OK;
}
