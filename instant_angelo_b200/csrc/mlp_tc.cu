// tcgen05 tensor-core variant of the width-64 fused MLP (placeholder until the UMMA kernel lands):
// the entry points exist so the ABI is complete and fail loudly instead of silently running fp32.
#include "ia_common.cuh"

int ia_mlp_fwd_tc(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, int32_t, float *, int64_t, void *)
{
    ia_set_error("mlp: IA_MLP_TC_F16 is not available in this build");
    return IA_ERR_UNSUPPORTED;
}

int ia_mlp_bwd_tc(const ia_mlp_desc *, const float *, const float *, int64_t, const float *, const float *, int32_t, int64_t,
                  float *, float *, float *, void *)
{
    ia_set_error("mlp: IA_MLP_TC_F16 is not available in this build");
    return IA_ERR_UNSUPPORTED;
}
