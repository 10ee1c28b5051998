// ia_mlp_fwd / ia_mlp_bwd entry points: select the arithmetic the descriptor asks for.
//   IA_MLP_FP32   -> mlp_fp32.cu  (FFMA, the VanillaMLP-exact path; reference models/network_utils.py:96-113)
//   IA_MLP_TC_F16 -> mlp_tc.cu    (tcgen05 tensor cores, fp16 operands / fp32 accumulate; the
//                                  "FullyFusedMLP" otype of reference models/network_utils.py:181-184)
#include "ia_common.cuh"

int ia_mlp_fwd_fp32(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, int32_t, float *, int64_t, void *);
int ia_mlp_bwd_fp32(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, const float *, int32_t, int64_t,
                    float *, float *, float *, void *);
int ia_mlp_fwd_tc(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, int32_t, float *, int64_t, void *);
int ia_mlp_bwd_tc(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, const float *, int32_t, int64_t,
                  float *, float *, float *, void *);

extern "C" int32_t ia_mlp_fwd(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                              int32_t n_out_used, float *out, int64_t ld_out, void *stream)
{
    IA_REQUIRE(desc != nullptr, "mlp_fwd: desc is NULL");
    if (desc->precision == IA_MLP_FP32) return ia_mlp_fwd_fp32(desc, in0, in1, n, params, n_out_used, out, ld_out, stream);
    if (desc->precision == IA_MLP_TC_F16) return ia_mlp_fwd_tc(desc, in0, in1, n, params, n_out_used, out, ld_out, stream);
    ia_set_error("mlp_fwd: unknown precision %d", desc->precision);
    return IA_ERR_INVALID_ARG;
}

extern "C" int32_t ia_mlp_bwd(const ia_mlp_desc *desc, const float *in0, const float *in1, int64_t n, const float *params,
                              const float *dout, int32_t n_out_used, int64_t ld_dout, float *din0, float *din1,
                              float *dparams, void *stream)
{
    IA_REQUIRE(desc != nullptr, "mlp_bwd: desc is NULL");
    if (desc->precision == IA_MLP_FP32)
        return ia_mlp_bwd_fp32(desc, in0, in1, n, params, dout, n_out_used, ld_dout, din0, din1, dparams, stream);
    if (desc->precision == IA_MLP_TC_F16)
        return ia_mlp_bwd_tc(desc, in0, in1, n, params, dout, n_out_used, ld_dout, din0, din1, dparams, stream);
    ia_set_error("mlp_bwd: unknown precision %d", desc->precision);
    return IA_ERR_INVALID_ARG;
}
