/* ia_b200.h -- C ABI of the B200-native (sm_100a) Instant-angelo training hot path.
 *
 * This is the drop-in boundary: plain pointers + sizes + a cudaStream_t (passed as void*), an int32
 * status return (0 = OK, negative = ia_status), no exceptions and no torch types.  Every entry point
 * names the reference interface it replaces (file:line under the hugoycj/Instant-angelo tree).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host or the parameter is a plan/desc
 *     struct (plain POD, read on the host, passed to kernels by value);
 *   - all tensors are contiguous, row-major, fp32 unless stated; the caller owns every allocation;
 *   - kernels are enqueued on `stream`; no call synchronises the device except ia_march_total();
 *   - functions are re-entrant; the only global state is the thread-local last-error string.
 */
#ifndef IA_B200_H
#define IA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IA_MAX_LEVELS 32
#define IA_ABI_VERSION 1

typedef enum ia_status {
    IA_OK = 0,
    IA_ERR_INVALID_ARG = -1,
    IA_ERR_UNSUPPORTED = -2,
    IA_ERR_CUDA = -3,
    IA_ERR_NO_DEVICE = -4
} ia_status;

/* Thread-local description of the last non-zero status returned on this thread. */
const char *ia_last_error_string(void);
int32_t ia_abi_version(void);
/* Compute capability of the current device (major*10+minor); IA_ERR_NO_DEVICE without a GPU. */
int32_t ia_device_arch(void);

/* ------------------------------------------------------------------------------------------------
 * Multiresolution hash grid            replaces tcnn.Encoding(otype=HashGrid) as constructed at
 *                                      models/network_utils.py:44-48 and called at :56-59
 * ------------------------------------------------------------------------------------------------ */
typedef struct ia_grid_plan {
    int32_t n_levels;
    int32_t n_features;                  /* must be 2 */
    int32_t log2_hashmap_size;
    int32_t base_resolution;
    float per_level_scale;
    float scale[IA_MAX_LEVELS];          /* exp2f(l*log2f(pls))*base - 1, float32 arithmetic */
    uint32_t res[IA_MAX_LEVELS];         /* ceilf(scale)+1 */
    uint32_t size[IA_MAX_LEVELS];        /* entries in the level */
    uint32_t offset[IA_MAX_LEVELS + 1];  /* entry offset of each level; offset[L] = total entries */
    uint32_t hashed[IA_MAX_LEVELS];      /* 1: coherent prime hash, 0: dense index */
} ia_grid_plan;

/* Host-only: fills `plan` with tcnn's per-level geometry (flat param layout: level-major, entry-major,
 * feature-minor; config keys of configs/neuralangelo-colmap_sparse.yaml:45-51). */
int32_t ia_hashgrid_plan(int32_t n_levels, int32_t n_features, int32_t log2_hashmap_size,
                         int32_t base_resolution, float per_level_scale, ia_grid_plan *plan_host);

/* out[n, L*F] = encode(x[n,3] in [0,1]); levels >= active_levels are written as exact zeros, which is
 * the progressive mask of ProgressiveBandHashGrid.forward (models/network_utils.py:56-59) folded in.
 * `table` is the flat fp32 parameter vector (tcnn Encoding.params). */
int32_t ia_hashgrid_fwd(const float *x, int64_t n, const float *table, const ia_grid_plan *plan_host,
                        int32_t active_levels, float *out, void *stream);

/* As ia_hashgrid_fwd, for rows that arrive in groups of `group` consecutive, spatially close points (the six finite-difference
 * taps of one sample, models/geometry.py:221-233): one thread walks the taps of a (group, level) and re-fetches the cell's
 * corners only when a tap leaves the cell.  Same values as ia_hashgrid_fwd up to fp32 rounding of the interpolation
 * (lerp chain instead of weight products).  group == 6 and n_levels <= 16 are specialised; anything else uses ia_hashgrid_fwd. */
int32_t ia_hashgrid_fwd_grouped(const float *x, int64_t n, const float *table, const ia_grid_plan *plan_host,
                                int32_t active_levels, int32_t group, float *out, void *stream);

/* dtable[...] += scatter(w_corner * dy[n, L*F]) (accumulates; caller zeroes).  Autograd of the call at
 * models/network_utils.py:57 w.r.t. Encoding.params. */
int32_t ia_hashgrid_bwd_table(const float *x, int64_t n, const float *dy, const ia_grid_plan *plan_host,
                              int32_t active_levels, float *dtable, void *stream);

/* dx[n,3] = d enc / d x contracted with dy (overwrites).  Needed because curvature tap positions
 * depend on parameters (models/geometry.py:238-246, no detach). */
int32_t ia_hashgrid_bwd_input(const float *x, int64_t n, const float *table, const float *dy,
                              const ia_grid_plan *plan_host, int32_t active_levels, float *dx, void *stream);

/* Both of the above in one pass; dtable and/or dx may be NULL. */
int32_t ia_hashgrid_bwd(const float *x, int64_t n, const float *table, const float *dy,
                        const ia_grid_plan *plan_host, int32_t active_levels, float *dtable, float *dx,
                        void *stream);

/* As ia_hashgrid_bwd, for rows that arrive in groups of `group` consecutive, spatially close points (the six
 * finite-difference taps of one sample, models/geometry.py:221-233): contributions of a group that fall into the same
 * grid cell are summed in registers and scattered once.  group == 6 is specialised; other values use ia_hashgrid_bwd.
 * Identical to ia_hashgrid_bwd up to fp32 summation order. */
int32_t ia_hashgrid_bwd_grouped(const float *x, int64_t n, const float *table, const float *dy,
                                const ia_grid_plan *plan_host, int32_t active_levels, int32_t group, float *dtable,
                                float *dx, void *stream);

/* Second-order adjoints of ia_hashgrid_bwd_input (dx = J(x; table)^T dy), needed when the SDF normals are autograd's
 * d sdf / d x with create_graph=True (grad_type: analytic, models/geometry.py:214-218; tcnn: kernel_grid_backward_input's
 * own backward, kernel_grid_backward_input_backward_grid).  v[n,3] = dL/d(dx):
 *   ia_hashgrid_jvp:                  out[n, L*F] = dL/d(dy)  = J(x; table) v              (masked levels: zeros)
 *   ia_hashgrid_bwd_input_bwd_table:  dtable (ACCUMULATED)   += d(v^T J(x; table)^T dy) / d(table)
 * d/dx of dx (mixed second derivatives of the trilinear interpolant) is not provided: positions carry no gradient on
 * the training path. */
int32_t ia_hashgrid_jvp(const float *x, int64_t n, const float *table, const float *v, const ia_grid_plan *plan,
                        int32_t active_levels, float *out, void *stream);
int32_t ia_hashgrid_bwd_input_bwd_table(const float *x, int64_t n, const float *v, const float *dy,
                                        const ia_grid_plan *plan, int32_t active_levels, float *dtable, void *stream);

/* fp16 shadow tables (opt-in, `table_precision: fp16` in the encoding config): the arena the optimizer owns stays fp32,
 * the gathers read a __half2-per-entry copy refreshed by ia_table_to_half once per optimizer step -- tcnn's own arrangement
 * (fp32 master parameters, fp16 copy for compute; the reference runs tcnn in fp16, models/network_utils.py:57).  Same
 * indices and interpolation as the fp32 entry points; table_h has plan->offset[n_levels] entries of 4 bytes; interpolation
 * and every gradient stay fp32.  group == 6 selects the grouped scatter as in ia_hashgrid_bwd_grouped, any other value the
 * plain kernel; dtable / dx may be NULL. */
int32_t ia_table_to_half(const float *table, int64_t n_floats, void *table_h, void *stream);
int32_t ia_hashgrid_fwd_h(const float *x, int64_t n, const void *table_h, const ia_grid_plan *plan_host,
                          int32_t active_levels, float *out, void *stream);
int32_t ia_hashgrid_bwd_h(const float *x, int64_t n, const void *table_h, const float *dy,
                          const ia_grid_plan *plan_host, int32_t active_levels, int32_t group, float *dtable, float *dx,
                          void *stream);
int32_t ia_hashgrid_jvp_h(const float *x, int64_t n, const void *table_h, const float *v, const ia_grid_plan *plan,
                          int32_t active_levels, float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Spherical harmonics                  replaces tcnn.Encoding(otype=SphericalHarmonics) built at
 *                                      models/network_utils.py:90-91, called at models/texture.py:25,52,129,134
 * ------------------------------------------------------------------------------------------------ */
int32_t ia_sh_fwd(const float *d01, int64_t n, int32_t degree, float *out /*[n,degree^2]*/, void *stream);
int32_t ia_sh_bwd(const float *d01, int64_t n, int32_t degree, const float *dout, float *dd01, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Width-64 fused MLP                   replaces VanillaMLP.forward (models/network_utils.py:108-113) and
 *                                      tcnn.Network (models/network_utils.py:181-184)
 * ------------------------------------------------------------------------------------------------ */
typedef enum ia_act { IA_ACT_NONE = 0, IA_ACT_RELU = 1, IA_ACT_SOFTPLUS100 = 2, IA_ACT_SIGMOID = 3 } ia_act;
typedef enum ia_mlp_precision { IA_MLP_FP32 = 0, IA_MLP_TC_F16 = 1 } ia_mlp_precision;

typedef struct ia_mlp_desc {
    int32_t n_in0;            /* leading input columns taken from in0 as in0*in0_scale+in0_offset (0..8) */
    float in0_scale, in0_offset;
    int32_t n_in1;            /* remaining input columns taken from in1 (row stride n_in1) */
    int32_t n_hidden_layers;  /* 1 or 2 hidden layers of `width` neurons */
    int32_t width;            /* must be 64 */
    int32_t n_out;            /* 1..128 */
    int32_t hidden_act;       /* IA_ACT_RELU | IA_ACT_SOFTPLUS100 */
    int32_t out_act;          /* IA_ACT_NONE | IA_ACT_SIGMOID */
    int32_t precision;        /* ia_mlp_precision */
} ia_mlp_desc;

/* Flat parameter layout (effective weights, i.e. after weight-norm is folded on the host side):
 *   W0[width, n_in0+n_in1], b0[width], (W1[width,width], b1[width]), Wl[n_out, width], bl[n_out]
 * each W row-major [out, in] like nn.Linear.weight. */
int64_t ia_mlp_param_count(const ia_mlp_desc *desc_host);

/* out[n, n_out_used] (row stride ld_out) = first n_out_used outputs of the network.  n_out_used < n_out
 * is the SDF-only evaluation of the finite-difference taps (models/geometry.py:233, 266).
 * IA_MLP_TC_F16 only: n_out_used <= 8 are fused; n_out_used == 0 returns the last hidden layer's activations
 * [n, width] instead (and ia_mlp_bwd then takes their gradient), so that a wide output layer -- the 65-feature centre
 * evaluation -- can be applied by the caller as one plain GEMM. */
int32_t ia_mlp_fwd(const ia_mlp_desc *desc_host, const float *in0, const float *in1, int64_t n,
                   const float *params, int32_t n_out_used, float *out, int64_t ld_out, void *stream);

/* Backward with in-kernel recomputation of the hidden activations.  dout[n, n_out_used] (row stride
 * ld_dout).  din0/din1 may be NULL; dparams (same layout as params) is ACCUMULATED. */
int32_t ia_mlp_bwd(const ia_mlp_desc *desc_host, const float *in0, const float *in1, int64_t n,
                   const float *params, const float *dout, int32_t n_out_used, int64_t ld_dout,
                   float *din0, float *din1, float *dparams, void *stream);

/* grad_type 'analytic' (models/geometry.py:206 + :214-218: `torch.autograd.grad(sdf, points_, create_graph=True)`): the SDF
 * network together with the derivative of its output column 0 (the SDF) w.r.t. its own inputs, on the tensor cores.
 * Shape: precision IA_MLP_TC_F16, n_in0 == 3, n_in1 == 32, two hidden layers, IA_ACT_SOFTPLUS100 (anything else returns
 * IA_ERR_UNSUPPORTED).  h_out[n, 64]: the last hidden layer (apply the output layer with ia_linear64_fwd / ia_sdf_head_fwd,
 * as after ia_mlp_fwd with n_out_used == 0); g0[n, 3] = d out0 / d in0, g1[n, 32] = d out0 / d in1.  Outputs may be NULL. */
int32_t ia_mlp_fwd_grad(const ia_mlp_desc *desc_host, const float *in0, const float *in1, int64_t n, const float *params,
                        float *h_out, float *g0, float *g1, void *stream);

/* Adjoint of ia_mlp_fwd_grad (a second-order adjoint of the network: what autograd's double backward computes for the
 * reference): cotangents dh[n, 64], dg0[n, 3], dg1[n, 32] (each may be NULL = zero) -> din0[n, 3], din1[n, 32] (overwritten,
 * may be NULL) and dparams (ACCUMULATED; row 0 of the output layer receives the contribution that arrives through g, the
 * contribution through h_out belongs to the caller's output-layer backward). */
int32_t ia_mlp_fwd_grad_bwd(const ia_mlp_desc *desc_host, const float *in0, const float *in1, int64_t n, const float *params,
                            const float *dh, const float *dg0, const float *dg1, float *din0, float *din1, float *dparams,
                            void *stream);

/* Fused VolumeSDF evaluation: out = network(cat[x*in0_scale+in0_offset, hashgrid(x)]) -- the composition
 * `self.network(self.encoding(points))` of models/geometry.py:206 (centre), :233 (six finite-difference taps) and :266 (six
 * curvature taps), i.e. the 13 evaluations per sample of SURVEY.md section 8(b) -- with the hash-grid gather done inside the
 * tensor-core MLP kernel's operand staging: thread (row, column group) gathers its four levels and writes the fp16 hi/lo
 * pair straight into the UMMA operand layout, so the [n, L*F] encoding never exists in HBM (neither in forward nor as a
 * saved tensor: backward re-gathers it).  desc: precision IA_MLP_TC_F16, n_in0 == 3 (the xyz pass-through of
 * CompositeEncoding, models/network_utils.py:75-78), n_in1 == L*F, L % 4 == 0.  x: [n,3] in the encoder's [0,1] coordinates
 * (any value is legal: tcnn's wrap-around indexing).  n_out_used as for ia_mlp_fwd (0: last hidden layer [n, 64]). */
int32_t ia_sdf_taps_fused_fwd(const ia_mlp_desc *desc_host, const ia_grid_plan *plan_host, int32_t active_levels,
                              const float *x, int64_t n, const float *table, const float *params, int32_t n_out_used,
                              float *out, int64_t ld_out, void *stream);

/* Backward of the above (two-hidden-layer Softplus networks = VolumeSDF): dparams and dtable are ACCUMULATED; the
 * position gradient comes in two parts, dx_enc[n,3] through the encoding and dx_direct[n,3] through the xyz pass-through
 * (each overwritten, each may be NULL; d out / d x = dx_enc + dx_direct).  `group` as for ia_hashgrid_bwd_grouped (6: rows
 * are the six taps of one sample).  denc_ws: caller-provided [n, L*F] fp32 workspace for the gradient w.r.t. the encoding
 * on its way from the MLP kernel to the table scatter (required unless dtable and dx_enc are both NULL). */
int32_t ia_sdf_taps_fused_bwd(const ia_mlp_desc *desc_host, const ia_grid_plan *plan_host, int32_t active_levels,
                              const float *x, int64_t n, const float *table, const float *params, const float *dout,
                              int32_t n_out_used, int64_t ld_dout, int32_t group, float *dtable, float *dx_enc,
                              float *dx_direct, float *dparams, float *denc_ws, void *stream);

/* Flat effective parameter vector of a VanillaMLP (models/network_utils.py:115-134) in one launch: per layer
 * W = g * v / ||v||_row when g != NULL (torch weight_norm, dim=0), else W = v; layout as above (W block, then bias).
 * Backward writes dg[n_out] / dv[n_out, n_in] / db[n_out] (each may be NULL) from dflat; pointers are device pointers,
 * the descriptor itself is read on the host. */
#define IA_WN_MAX_LAYERS 4
typedef struct ia_wn_desc {
    int32_t n_layers;
    int32_t n_out[IA_WN_MAX_LAYERS], n_in[IA_WN_MAX_LAYERS];
    const float *g[IA_WN_MAX_LAYERS], *v[IA_WN_MAX_LAYERS], *b[IA_WN_MAX_LAYERS];
    float *dg[IA_WN_MAX_LAYERS], *dv[IA_WN_MAX_LAYERS], *db[IA_WN_MAX_LAYERS];
} ia_wn_desc;
int32_t ia_weightnorm_flat_fwd(const ia_wn_desc *desc_host, float *flat, void *stream);
int32_t ia_weightnorm_flat_bwd(const ia_wn_desc *desc_host, const float *dflat, void *stream);
/* Same adjoint, ADDED to dg / dv / db (the parameters' slices of a gradient arena that is zeroed once per step). */
int32_t ia_weightnorm_flat_bwd_acc(const ia_wn_desc *desc_host, const float *dflat, void *stream);

/* Wide output layer for the feature mode above: out[n, n_out] = h[n, 64] W[n_out, 64]^T + b (n_out <= 128), fp32.
 * Backward: dh[n,64] = dout W (may be NULL); dW[n_out,64] += dout^T h and db[n_out] += sum dout (ACCUMULATED; may be NULL). */
int32_t ia_linear64_fwd(const float *h, int64_t n, const float *W, const float *b, int32_t n_out, float *out,
                        int64_t ld_out, void *stream);
/* dextra[n, n_extra] (may be NULL with n_extra = 0) is added to the first n_extra columns of dout on the fly: gradients that
 * reach a few output columns through a second consumer (the SDF value and the diffuse albedo columns of the feature
 * vector, models/geometry.py:206-207 + models/texture.py:58) need no separate accumulation pass. */
int32_t ia_linear64_bwd(const float *h, int64_t n, const float *W, const float *dout, int64_t ld_dout, int32_t n_out,
                        const float *dextra, int32_t n_extra, float *dh, float *dW, float *db, void *stream);

/* SDF output layer fused with the colour head's input row (models/geometry.py:206-207 `feature = cat[out, points*2-1]`
 * + models/texture.py:26-27 `cat[features, dirs_embd, normals]`):
 *   tin[n, ld_tin] = [ h W^T + b (n_feat <= 72 cols) | pts01*2-1 (3) | enc (n_enc <= 26) | normal (3) ],
 *   sdf[n] = column 0, rgb_raw[n,3] = columns 1..3 (the dual-colour head's diffuse term; may be NULL).
 * One kernel writes whole rows; nothing of `out` / `feature` / `network_inp` is materialised separately.
 * Backward: dtin[n, ld_tin] is the colour MLP's input gradient; dextra[n, n_extra] (gradients reaching columns
 * 0..n_extra-1 through sdf / rgb_raw) is added on the fly.  dh[n,64] required; dW/db ACCUMULATED (may be NULL);
 * dpts01 = 2 dtin[:, n_feat:n_feat+3], denc, dnormal = the matching column blocks (each may be NULL). */
int32_t ia_sdf_head_fwd(const float *h, int64_t n, const float *W, const float *b, int32_t n_feat, const float *pts01,
                        const float *enc, int32_t n_enc, const float *normal, float *tin, int64_t ld_tin, float *sdf,
                        float *rgb_raw, void *stream);
int32_t ia_sdf_head_bwd(const float *h, int64_t n, const float *W, const float *dtin, int64_t ld_tin, int32_t n_feat,
                        int32_t n_enc, const float *dextra, int32_t n_extra, float *dh, float *dW, float *db,
                        float *dpts01, float *denc, float *dnormal, void *stream);

/* The same input row with the SDF network's wide output layer FOLDED into the colour network's first layer: with
 * out = Wl h + bl, colour layer 0 is z = (Wc0[:, :n_feat] Wl) h + Wc0[:, n_feat:] [pts | enc | normal] + (bc0 + Wc0[:, :n_feat] bl), so the
 * colour network reads h directly and the [n, n_feat] geometry output (models/geometry.py:206-207) is never formed; the caller
 * composes the two small matrices.  tin[n, ld] = [h (64) | pts01*2-1 (3) | enc (n_enc) | normal (3) | zeros up to ld] (ld % 4 == 0);
 * the four geometry outputs used outside the colour network -- sdf = out[:, 0] and the dual-colour diffuse term out[:, 1:4]
 * (models/texture.py:58) -- come from W4 = Wl[0:4], b4 = bl[0:4]. */
int32_t ia_colour_in_fwd(const float *h, int64_t n, const float *W4, const float *b4, const float *pts01, const float *enc,
                         int32_t n_enc, const float *normal, float *tin, int64_t ld, float *sdf, float *rgb_raw, void *stream);
/* dh[n,64] = dtin[:, :64] + [dsdf | drgb] W4 (overwritten); dpts01 / denc / dnormal: column blocks of dtin (overwritten, may be
 * NULL); dW4[4,64], db4[4] ACCUMULATED (may be NULL). */
int32_t ia_colour_in_bwd(const float *h, int64_t n, const float *W4, const float *dtin, int64_t ld, int32_t n_enc,
                         const float *dsdf, const float *drgb, float *dh, float *dW4, float *db4, float *dpts01, float *denc,
                         float *dnormal, void *stream);
/* The fold in parameter space (once per step): flat = colour network [Wc0[64, n_in] | bc0[64] | rest[n_rest]], the geometry
 * output layer w_last[n_feat, 64], b_last[n_feat] ->
 *   flat_eff = [W_eff[64, ld] | b_eff[64] | rest],  W_eff[o] = [Wc0[o, :n_feat] w_last | Wc0[o, n_feat:] | 0], b_eff = bc0 + Wc0[:, :n_feat] b_last
 * (Wc0 [feature | rest] = (Wc0[:, :n_feat] Wl) h + ...: models/texture.py:26-29 applied to models/geometry.py:206-207).
 * Backward: dflat is ADDED to (NULL = not wanted), dw_last / db_last are written (NULL = not wanted). */
int32_t ia_fold_head_fwd(const float *flat, const float *w_last, const float *b_last, int32_t n_in, int32_t n_feat, int32_t ld,
                         int64_t n_rest, float *flat_eff, void *stream);
int32_t ia_fold_head_bwd(const float *dflat_eff, const float *flat, const float *w_last, const float *b_last, int32_t n_in,
                         int32_t n_feat, int32_t ld, int64_t n_rest, float *dflat, float *dw_last, float *db_last, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Finite-difference / curvature stages  fused elementwise stages of VolumeSDF.forward
 *                                       (models/geometry.py:219-234 FD taps + gradient, :236-275 curvature)
 * ------------------------------------------------------------------------------------------------ */
/* taps01[n,6,3] = ((clamp(base[n,3] + eps*e_k, -radius, radius)) + radius) / (2 radius), e_k = +x,-x,+y,-y,+z,-z */
int32_t ia_fd_taps_fwd(const float *base, int64_t n, float eps, float radius, float *taps01, void *stream);
int32_t ia_fd_taps_bwd(const float *base, int64_t n, float eps, float radius, const float *dtaps01, float *dbase,
                       void *stream);
/* grad[n,3] = 0.5 * (sdf6[n,2i] - sdf6[n,2i+1]) / eps */
int32_t ia_fd_grad_fwd(const float *sdf6, int64_t n, float eps, float *grad, void *stream);
int32_t ia_fd_grad_bwd(const float *dgrad, int64_t n, float eps, float *dsdf6, void *stream);
/* normals = normalize(grad); shifted = pts01 + cross(normals, normalize(rnd)) * eps  (reference quirks kept) */
int32_t ia_curv_shift_fwd(const float *grad, const float *rnd, const float *pts01, int64_t n, float eps, float *normals,
                          float *shifted, void *stream);
int32_t ia_curv_shift_bwd(const float *grad, const float *rnd, int64_t n, float eps, const float *dnormals,
                          const float *dshifted, float *dgrad, void *stream);
/* laplace[n] = acos(clamp(normals . normalize(gshift), -1+1e-6, 1-1e-6)) / pi */
int32_t ia_curv_angle_fwd(const float *normals, const float *gshift, int64_t n, float *laplace, void *stream);
int32_t ia_curv_angle_bwd(const float *normals, const float *gshift, int64_t n, const float *dlaplace, float *dnormals,
                          float *dgshift, void *stream);

/* Sample points of the marched intervals (models/neus.py:153-157 background, :218-223 foreground):
 *   midpoints = (t_starts + t_ends) / 2, positions = rays_o[ri] + rays_d[ri] * midpoints, t_dirs = rays_d[ri],
 *   dists = t_ends - t_starts; each operation rounded as the tensor expression (no fused multiply-add): the background
 * marcher prunes on a density evaluated at these positions.  t_dirs / midpoints / dists may be NULL. */
int32_t ia_ray_samples(const float *rays_o, const float *rays_d, const int32_t *ray_indices, const float *t_starts,
                       const float *t_ends, int64_t n, float *positions, float *t_dirs, float *midpoints, float *dists,
                       void *stream);
/* contract_to_unisphere (models/geometry.py:19-31) of positions that carry no gradient: IA_AABB -> (x + r) / 2r;
 * IA_UN_BOUNDED_SPHERE -> that, then y = 2x - 1, y <- (2 - 1/|y|) y/|y| where |y| > 1, y/4 + 0.5. */
int32_t ia_contract(const float *x, int64_t n, float radius, int32_t contraction_type, float *out, void *stream);
/* F.normalize(x, p=2, dim=-1, eps) for x[n,3] (models/neus.py:229, 247; systems/neus.py:182-183) and its adjoint. */
int32_t ia_normalize3_fwd(const float *x, int64_t n, float eps, float *out, void *stream);
int32_t ia_normalize3_bwd(const float *x, const float *dout, int64_t n, float eps, float *dx, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Occupancy grid + ray marching        replaces nerfacc.OccupancyGrid / ray_aabb_intersect /
 *                                      ray_marching as called at models/neus.py:64-74, 108-111, 153,
 *                                      159-169, 209-220
 * ------------------------------------------------------------------------------------------------ */
typedef enum ia_contraction { IA_AABB = 0, IA_UN_BOUNDED_TANH = 1, IA_UN_BOUNDED_SPHERE = 2 } ia_contraction;

typedef struct ia_grid_desc {
    float roi[6];             /* xmin ymin zmin xmax ymax zmax */
    int32_t res[3];
    int32_t contraction;      /* ia_contraction */
} ia_grid_desc;

/* occs[idx[i]] = max(occs[idx[i]]*decay, max over duplicates occ[i]); then
 * binary = occs > min(mean(occs), thre) written both as bool bytes (nerfacc `_binary`, C order
 * x*ry*rz + y*rz + z) and as a packed bitfield (bit c of word c/32).  `idx` int64, NULL = all cells in
 * order.  workspace: >= ia_occ_workspace_bytes(num_cells) bytes. */
int64_t ia_occ_workspace_bytes(int64_t num_cells);
int32_t ia_occ_update(const int64_t *idx, const float *occ, int64_t n, float *occs, int64_t num_cells,
                      float ema_decay, float occ_thre, uint8_t *binary, uint32_t *bitfield, void *workspace,
                      void *stream);
/* bool bytes -> bitfield (for grids loaded from a checkpoint or set by hand). */
int32_t ia_occ_pack(const uint8_t *binary, int64_t num_cells, uint32_t *bitfield, void *stream);

/* nerfacc ray_aabb_intersect: miss => (1e10,1e10); clamp_zero applies t_min = max(t_min, 0). */
int32_t ia_aabb(const float *rays_o, const float *rays_d, int64_t n_rays, const float *aabb_host6,
                int32_t clamp_zero, float *t_min, float *t_max, void *stream);

/* Pass 1: num_steps[r] = number of marched samples of ray r.  bitfield may be NULL (all occupied). */
int32_t ia_march_count(const float *rays_o, const float *rays_d, const float *t_min, const float *t_max,
                       int64_t n_rays, const ia_grid_desc *grid_host, const uint32_t *bitfield,
                       float step_size, float cone_angle, int32_t *num_steps, void *stream);
/* Exclusive scan: packed_info[r] = (offset, count); *total_dev (int64) = sum.  workspace >=
 * ia_march_scan_workspace_bytes(n_rays). */
int64_t ia_march_scan_workspace_bytes(int64_t n_rays);
int32_t ia_march_scan(const int32_t *num_steps, int64_t n_rays, int32_t *packed_info, int64_t *total_dev,
                      void *workspace, void *stream);
/* Synchronises `stream` and returns *total_dev on the host (the one D2H sync nerfacc also performs). */
int32_t ia_march_total(const int64_t *total_dev, int64_t *total_host, void *stream);
/* Pass 2: writes ray_indices[S] (int32), t_starts[S], t_ends[S]. */
int32_t ia_march_write(const float *rays_o, const float *rays_d, const float *t_min, const float *t_max,
                       int64_t n_rays, const ia_grid_desc *grid_host, const uint32_t *bitfield,
                       float step_size, float cone_angle, const int32_t *packed_info, int32_t *ray_indices,
                       float *t_starts, float *t_ends, void *stream);
/* nerfacc render_visibility on packed samples (sequential per-ray transmittance). */
/* Two marches of the same rays in one launch pair (the foreground march through the AABB grid and the background march through
 * the contracted grid, models/neus.py:209-220 and :159-169): ia_march_pair(write = 0) counts both, two ia_march_scan calls and
 * ONE ia_march_totals read-back follow, ia_march_pair(write = 1) emits both.  Same per-ray function as ia_march_count / _write. */
typedef struct ia_march_set {
    const float *t_min, *t_max;
    const ia_grid_desc *grid;       /* host */
    const uint32_t *bitfield;
    float step_size, cone_angle;
    const int32_t *packed_info;     /* write pass */
    int32_t *num_steps;             /* count pass */
    int32_t *ray_indices;           /* write pass */
    float *t_starts, *t_ends;       /* write pass */
} ia_march_set;
int32_t ia_march_pair(const float *rays_o, const float *rays_d, int64_t n_rays, const ia_march_set *a_host, const ia_march_set *b_host,
                      int32_t write, void *stream);
/* totals_host[count] <- totals_dev[count]; synchronises the stream (like ia_march_total, for several totals at once). */
int32_t ia_march_totals(const int64_t *totals_dev, int32_t count, int64_t *totals_host, void *stream);
int32_t ia_visibility(const float *alphas, const int32_t *packed_info, int64_t n_rays, float early_stop_eps,
                      float alpha_thre, uint8_t *visible, void *stream);
/* The visibility pruning of nerfacc.ray_marching with a sigma_fn (models/neus.py:144-149, 159-169) together with the
 * compaction it is followed by: alpha = 1 - exp(-sigma (t_end - t_start)) per sample, the transmittance test of
 * ia_visibility, and -- count -> ia_march_scan -> ia_march_total -> write, the marcher's structure -- the kept samples'
 * ray_indices / t_starts / t_ends with their packed_info.  visible [S] and num_kept [n_rays] are scratch between the calls. */
int32_t ia_prune_count(const float *sigmas, const float *t_starts, const float *t_ends, const int32_t *packed_info,
                       int64_t n_rays, float early_stop_eps, float alpha_thre, uint8_t *visible, int32_t *num_kept, void *stream);
int32_t ia_prune_write(const uint8_t *visible, const int32_t *packed_info, const int32_t *packed_info_kept,
                       const float *t_starts, const float *t_ends, int64_t n_rays, int32_t *ray_indices_kept,
                       float *t_starts_kept, float *t_ends_kept, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Per-ray segmented compositing        replaces NeuSModel.get_alpha (models/neus.py:117-139) +
 *                                      nerfacc.render_weight_from_alpha/_density + accumulate_along_rays
 *                                      (models/neus.py:181-184, 234-239)
 * ------------------------------------------------------------------------------------------------ */
typedef enum ia_alpha_mode { IA_ALPHA_GIVEN = 0, IA_ALPHA_NEUS = 1, IA_ALPHA_DENSITY = 2 } ia_alpha_mode;

typedef struct ia_composite_args {
    int32_t mode;               /* ia_alpha_mode */
    int64_t n_rays, n_samples;
    const int32_t *packed_info; /* [n_rays,2] (offset,count), samples sorted by ray */
    /* IA_ALPHA_GIVEN */
    const float *alpha_in;      /* [S] */
    /* IA_ALPHA_NEUS */
    const float *sdf;           /* [S] */
    const float *normal;        /* [S,3] unit normals */
    const float *dirs;          /* [S,3] */
    const float *dists;         /* [S] */
    const float *inv_s;         /* device scalar, already clipped to [1e-6,1e6] */
    float cos_anneal_ratio;
    /* IA_ALPHA_DENSITY */
    const float *sigma;         /* [S] */
    const float *t_starts, *t_ends; /* [S] */
    /* values to accumulate (any may be NULL) */
    const float *t_mid;         /* [S]   -> depth   */
    const float *rgb;           /* [S,3] -> comp_rgb */
    const float *nrm;           /* [S,3] -> comp_normal (un-normalised sum) */
} ia_composite_args;

/* Forward: alpha[S], trans[S] (T_i = prod_{j<i}(1-alpha_j), saved for backward), weights[S],
 * opacity[R], depth[R], comp_rgb[R,3], comp_normal[R,3] (outputs for absent values may be NULL). */
int32_t ia_composite_fwd(const ia_composite_args *args_host, float *alpha, float *trans, float *weights,
                         float *opacity, float *depth, float *comp_rgb, float *comp_normal, void *stream);
/* Backward: upstream grads (NULL = zero) -> grads of the per-sample inputs (NULL = not wanted).
 * d_inv_s (device scalar) is ACCUMULATED. */
int32_t ia_composite_bwd(const ia_composite_args *args_host, const float *alpha, const float *trans,
                         const float *g_weights, const float *g_opacity, const float *g_depth,
                         const float *g_comp_rgb, const float *g_comp_normal,
                         float *d_alpha_in, float *d_sdf, float *d_normal, float *d_inv_s, float *d_sigma,
                         float *d_rgb, float *d_nrm, void *stream);

/* Per-ray mix of the foreground and background renders (models/neus.py:186, 272-276):
 *   comp_rgb_bg = comp_rgb_bg_raw + background_color (1 - opacity_bg), comp_rgb_full = comp_rgb + comp_rgb_bg (1 - opacity),
 *   rays_valid = opacity > 0, rays_valid_bg = opacity_bg > 0, rays_valid_full = rays_valid | rays_valid_bg  (one byte per ray).
 * Backward: g_* may be NULL (zero), d_* may be NULL (not wanted); background_color [3] is a device array without gradient. */
int32_t ia_ray_mix_fwd(const float *comp_rgb, const float *opacity, const float *comp_rgb_bg_raw, const float *opacity_bg,
                       const float *background_color, int64_t n_rays, float *comp_rgb_bg, float *comp_rgb_full,
                       uint8_t *rays_valid, uint8_t *rays_valid_bg, uint8_t *rays_valid_full, void *stream);
int32_t ia_ray_mix_bwd(const float *opacity, const float *opacity_bg, const float *background_color, const float *comp_rgb_bg,
                       const float *g_comp_rgb_bg, const float *g_comp_rgb_full, int64_t n_rays, float *d_comp_rgb,
                       float *d_opacity, float *d_comp_rgb_bg_raw, float *d_opacity_bg, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Loss terms of the training step      replaces the tensor expressions of systems/neus.py:132-160 and
 *                                      systems/criterions.py:155-159 (binary_cross_entropy)
 * ------------------------------------------------------------------------------------------------ */
typedef struct ia_loss_args {
    int64_t n_rays, n_samples;
    float lambda_rgb_mse, lambda_rgb_l1, lambda_eikonal, lambda_mask, lambda_opaque, lambda_sparsity, lambda_curvature;
    float sparsity_scale;
} ia_loss_args;
/* Forward: terms[8] = (rgb_mse, rgb_l1, eikonal, mask, opaque, sparsity, curvature, n_valid) and
 * loss[1] = sum lambda_k terms_k in the reference's order of additions.
 *   rgb_*   : mean over the channels of the rays with valid[r] != 0 of (comp_rgb - rgb_gt)^2 / |.|      (:134-138)
 *   eikonal : mean_s (|sdf_grad_s| - 1)^2                                                                (:140-141)
 *   opaque  : BCE(o, o), mask: BCE(o, fg_mask) on o = clamp(opacity, 1e-3, 1 - 1e-3); fg_mask NULL = no mask term (:143-150)
 *   sparsity: mean_s exp(-sparsity_scale |sdf_s|)                                                        (:152-153)
 *   curvature: mean_s |laplace_s|, only when laplace != NULL and lambda_curvature > 0                    (:155-159)
 * Means over empty sets are nan, as in the reference.  valid is one byte per ray.  workspace: device memory of
 * ia_neus_losses_workspace_bytes() (zeroed by the call). */
int64_t ia_neus_losses_workspace_bytes(void);
int32_t ia_neus_losses_fwd(const ia_loss_args *args_host, const float *comp_rgb, const float *rgb_gt, const uint8_t *valid,
                           const float *opacity, const float *fg_mask, const float *sdf_grad, const float *sdf,
                           const float *laplace, void *workspace, float *terms, float *loss, void *stream);
/* Backward of loss[0]: dloss is a DEVICE scalar (the upstream gradient); each d_* may be NULL (not wanted) and is
 * overwritten.  The gradient of BCE(o, o) flows through both of its arguments, as autograd's does in the reference. */
int32_t ia_neus_losses_bwd(const ia_loss_args *args_host, const float *comp_rgb, const float *rgb_gt, const uint8_t *valid,
                           const float *opacity, const float *fg_mask, const float *sdf_grad, const float *sdf,
                           const float *laplace, const float *terms, const float *dloss, float *d_comp_rgb,
                           float *d_opacity, float *d_sdf_grad, float *d_sdf, float *d_laplace, void *stream);

/* The sparse-point terms (systems/neus.py:173-186) from the SDF and its gradient at the n SfM points:
 *   out4 = (sdf_l1, normal_cos, lambda_sdf_l1 sdf_l1 + lambda_normal normal_cos, mean(weights)),
 *   sdf_l1 = mean|sdf| * mean(weights) (the reference multiplies the SCALAR l1 loss by the weights: Appendix C-11),
 *   normal_cos = mean(1 - normalize(grad) . normalize(normal_gt)).  Backward of out4[2] w.r.t. sdf and grad (dloss: device
 * scalar; weights and normal_gt carry no gradient).  workspace: ia_neus_losses_workspace_bytes(). */
int32_t ia_point_losses_fwd(const float *sdf, const float *grad, const float *normal_gt, const float *weights, int64_t n,
                            float lambda_sdf_l1, float lambda_normal, void *workspace, float *out4, void *stream);
int32_t ia_point_losses_bwd(const float *sdf, const float *grad, const float *normal_gt, int64_t n, float lambda_sdf_l1,
                            float lambda_normal, const float *out4, const float *dloss, float *d_sdf, float *d_grad, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Fused AdamW over a flat arena        replaces torch.optim.AdamW as configured at
 *                                      configs/neuralangelo-colmap_sparse.yaml:134-139 (systems/utils.py:314-325)
 * ------------------------------------------------------------------------------------------------ */
int32_t ia_adamw_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                      float grad_scale, void *stream);

/* ------------------------------------------------------------------------------------------------
 * L2 residency of the hash tables      (no reference counterpart: tcnn leaves the tables to the L2's LRU;
 *                                      the tables gathered at models/network_utils.py:56-59 are re-fetched
 *                                      from HBM after every [N, L*F] activation pass on a 126 MB L2)
 * ------------------------------------------------------------------------------------------------ */
/* Marks [base, base + bytes) as a persisting access-policy window of `stream` (every kernel enqueued on it
 * afterwards) and reserves the matching L2 set-aside; bytes is clipped to the device's limits, hit_ratio is
 * the fraction of the window that keeps the persisting property.  bytes == 0 removes the window and resets
 * the persisting lines.  info_host (NULL or 3 int64): L2 size, set-aside granted, window bytes granted. */
int32_t ia_l2_persist(const void *base, int64_t bytes, float hit_ratio, int64_t *info_host, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Measurement aids (not part of the drop-in surface; used by bench.py and tools/)
 * ------------------------------------------------------------------------------------------------ */
/* Random 32-byte-sector gather micro-benchmark (the L2 peak "must be measured on the box"): n_threads
 * (multiple of 256) threads issue `iters` independent 8-byte loads each at pseudo-random sector-aligned addresses inside
 * [table, table + 32 n_sectors); out[n_threads] is practically never written. */
int32_t ia_debug_sector_gather(const float *table, int64_t n_sectors, int64_t n_threads, int32_t iters, float *out,
                               void *stream);
/* A/B switch of ia_hashgrid_fwd: on = 1 launches the one-loop kernel (dense / hashed index chosen per level under a
 * predicate), on = 0 the split-loop kernel (default when the plan's dense levels come first), on = -1 lets the environment
 * variable IA_HASHGRID_FWD_GENERIC decide.  Both compute the same indices and the same interpolation expression. */
int32_t ia_debug_hashgrid_fwd_generic(int32_t on);
/* Cycle accounting of the tensor-core MLP kernels (thread 0 of every CTA): enable != 0 starts it; out8_host (may be NULL)
 * receives {barrier wait, MMA issue, MMA completion wait, tile total, tiles, wait m0, wait m1, wait m2} and clears them. */
int32_t ia_debug_tc_timing(int32_t enable, unsigned long long *out8_host);

#ifdef __cplusplus
}
#endif
#endif /* IA_B200_H */
