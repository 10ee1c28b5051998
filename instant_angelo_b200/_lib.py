"""ctypes binding of the C ABI declared in include/ia_b200.h (libia_b200.so, built by build.py).

There is NO fallback: if the shared library is missing or a call returns a non-zero status a
RuntimeError is raised (mirroring tcnn's runtime errors / nerfacc's refusal of non-CUDA inputs).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IA_LIB_PATH") or os.path.join(_HERE, "_build", "libia_b200.so")   # IA_LIB_PATH: A/B builds (build.py)
IA_MAX_LEVELS = 32

IA_ACT_NONE, IA_ACT_RELU, IA_ACT_SOFTPLUS100, IA_ACT_SIGMOID = 0, 1, 2, 3
IA_MLP_FP32, IA_MLP_TC_F16 = 0, 1
IA_AABB, IA_UN_BOUNDED_TANH, IA_UN_BOUNDED_SPHERE = 0, 1, 2
IA_ALPHA_GIVEN, IA_ALPHA_NEUS, IA_ALPHA_DENSITY = 0, 1, 2


class GridPlan(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("n_features", C.c_int32), ("log2_hashmap_size", C.c_int32),
                ("base_resolution", C.c_int32), ("per_level_scale", C.c_float),
                ("scale", C.c_float * IA_MAX_LEVELS), ("res", C.c_uint32 * IA_MAX_LEVELS),
                ("size", C.c_uint32 * IA_MAX_LEVELS), ("offset", C.c_uint32 * (IA_MAX_LEVELS + 1)),
                ("hashed", C.c_uint32 * IA_MAX_LEVELS)]

    @property
    def n_entries(self) -> int:
        return int(self.offset[self.n_levels])

    @property
    def n_params(self) -> int:
        return self.n_entries * self.n_features


class MlpDesc(C.Structure):
    _fields_ = [("n_in0", C.c_int32), ("in0_scale", C.c_float), ("in0_offset", C.c_float), ("n_in1", C.c_int32),
                ("n_hidden_layers", C.c_int32), ("width", C.c_int32), ("n_out", C.c_int32), ("hidden_act", C.c_int32),
                ("out_act", C.c_int32), ("precision", C.c_int32)]


class GridDesc(C.Structure):
    _fields_ = [("roi", C.c_float * 6), ("res", C.c_int32 * 3), ("contraction", C.c_int32)]


class MarchSet(C.Structure):
    _fields_ = [("t_min", C.c_void_p), ("t_max", C.c_void_p), ("grid", C.POINTER(GridDesc)), ("bitfield", C.c_void_p),
                ("step_size", C.c_float), ("cone_angle", C.c_float), ("packed_info", C.c_void_p), ("num_steps", C.c_void_p),
                ("ray_indices", C.c_void_p), ("t_starts", C.c_void_p), ("t_ends", C.c_void_p)]


class CompositeArgs(C.Structure):
    _fields_ = [("mode", C.c_int32), ("n_rays", C.c_int64), ("n_samples", C.c_int64), ("packed_info", C.c_void_p),
                ("alpha_in", C.c_void_p), ("sdf", C.c_void_p), ("normal", C.c_void_p), ("dirs", C.c_void_p),
                ("dists", C.c_void_p), ("inv_s", C.c_void_p), ("cos_anneal_ratio", C.c_float), ("sigma", C.c_void_p),
                ("t_starts", C.c_void_p), ("t_ends", C.c_void_p), ("t_mid", C.c_void_p), ("rgb", C.c_void_p),
                ("nrm", C.c_void_p)]


IA_WN_MAX_LAYERS = 4


class WnDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("n_out", C.c_int32 * IA_WN_MAX_LAYERS), ("n_in", C.c_int32 * IA_WN_MAX_LAYERS),
                ("g", C.c_void_p * IA_WN_MAX_LAYERS), ("v", C.c_void_p * IA_WN_MAX_LAYERS), ("b", C.c_void_p * IA_WN_MAX_LAYERS),
                ("dg", C.c_void_p * IA_WN_MAX_LAYERS), ("dv", C.c_void_p * IA_WN_MAX_LAYERS), ("db", C.c_void_p * IA_WN_MAX_LAYERS)]


class LossArgs(C.Structure):
    _fields_ = [("n_rays", C.c_int64), ("n_samples", C.c_int64), ("lambda_rgb_mse", C.c_float), ("lambda_rgb_l1", C.c_float),
                ("lambda_eikonal", C.c_float), ("lambda_mask", C.c_float), ("lambda_opaque", C.c_float),
                ("lambda_sparsity", C.c_float), ("lambda_curvature", C.c_float), ("sparsity_scale", C.c_float)]


_P = C.c_void_p
_I32, _I64, _F = C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); kept in sync with include/ia_b200.h (tests/test_abi.py parses the header)
SIGNATURES = {
    "ia_last_error_string": (C.c_char_p, []),
    "ia_abi_version": (_I32, []),
    "ia_device_arch": (_I32, []),
    "ia_hashgrid_plan": (_I32, [_I32, _I32, _I32, _I32, _F, C.POINTER(GridPlan)]),
    "ia_hashgrid_fwd": (_I32, [_P, _I64, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_hashgrid_fwd_grouped": (_I32, [_P, _I64, _P, C.POINTER(GridPlan), _I32, _I32, _P, _P]),
    "ia_hashgrid_bwd_table": (_I32, [_P, _I64, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_hashgrid_bwd_input": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_hashgrid_bwd": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _P, _P, _P]),
    "ia_hashgrid_bwd_grouped": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _I32, _P, _P, _P]),
    "ia_hashgrid_jvp": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_hashgrid_bwd_input_bwd_table": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_table_to_half": (_I32, [_P, _I64, _P, _P]),
    "ia_hashgrid_fwd_h": (_I32, [_P, _I64, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_hashgrid_bwd_h": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _I32, _P, _P, _P]),
    "ia_hashgrid_jvp_h": (_I32, [_P, _I64, _P, _P, C.POINTER(GridPlan), _I32, _P, _P]),
    "ia_sh_fwd": (_I32, [_P, _I64, _I32, _P, _P]),
    "ia_sh_bwd": (_I32, [_P, _I64, _I32, _P, _P, _P]),
    "ia_mlp_param_count": (_I64, [C.POINTER(MlpDesc)]),
    "ia_mlp_fwd": (_I32, [C.POINTER(MlpDesc), _P, _P, _I64, _P, _I32, _P, _I64, _P]),
    "ia_mlp_bwd": (_I32, [C.POINTER(MlpDesc), _P, _P, _I64, _P, _P, _I32, _I64, _P, _P, _P, _P]),
    "ia_mlp_fwd_grad": (_I32, [C.POINTER(MlpDesc), _P, _P, _I64, _P, _P, _P, _P, _P]),
    "ia_mlp_fwd_grad_bwd": (_I32, [C.POINTER(MlpDesc), _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ia_sdf_taps_fused_fwd": (_I32, [C.POINTER(MlpDesc), C.POINTER(GridPlan), _I32, _P, _I64, _P, _P, _I32, _P, _I64, _P]),
    "ia_sdf_taps_fused_bwd": (_I32, [C.POINTER(MlpDesc), C.POINTER(GridPlan), _I32, _P, _I64, _P, _P, _P, _I32, _I64, _I32, _P, _P, _P,
                                     _P, _P, _P]),
    "ia_weightnorm_flat_fwd": (_I32, [C.POINTER(WnDesc), _P, _P]),
    "ia_weightnorm_flat_bwd": (_I32, [C.POINTER(WnDesc), _P, _P]),
    "ia_weightnorm_flat_bwd_acc": (_I32, [C.POINTER(WnDesc), _P, _P]),
    "ia_linear64_fwd": (_I32, [_P, _I64, _P, _P, _I32, _P, _I64, _P]),
    "ia_linear64_bwd": (_I32, [_P, _I64, _P, _P, _I64, _I32, _P, _I32, _P, _P, _P, _P]),
    "ia_sdf_head_fwd": (_I32, [_P, _I64, _P, _P, _I32, _P, _P, _I32, _P, _P, _I64, _P, _P, _P]),
    "ia_sdf_head_bwd": (_I32, [_P, _I64, _P, _P, _I64, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P, _P, _P]),
    "ia_colour_in_fwd": (_I32, [_P, _I64, _P, _P, _P, _P, _I32, _P, _P, _I64, _P, _P, _P]),
    "ia_colour_in_bwd": (_I32, [_P, _I64, _P, _P, _I64, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ia_fold_head_fwd": (_I32, [_P, _P, _P, _I32, _I32, _I32, _I64, _P, _P]),
    "ia_fold_head_bwd": (_I32, [_P, _P, _P, _P, _I32, _I32, _I32, _I64, _P, _P, _P, _P]),
    "ia_fd_taps_fwd": (_I32, [_P, _I64, _F, _F, _P, _P]),
    "ia_fd_taps_bwd": (_I32, [_P, _I64, _F, _F, _P, _P, _P]),
    "ia_fd_grad_fwd": (_I32, [_P, _I64, _F, _P, _P]),
    "ia_fd_grad_bwd": (_I32, [_P, _I64, _F, _P, _P]),
    "ia_curv_shift_fwd": (_I32, [_P, _P, _P, _I64, _F, _P, _P, _P]),
    "ia_curv_shift_bwd": (_I32, [_P, _P, _I64, _F, _P, _P, _P, _P]),
    "ia_curv_angle_fwd": (_I32, [_P, _P, _I64, _P, _P]),
    "ia_curv_angle_bwd": (_I32, [_P, _P, _I64, _P, _P, _P, _P]),
    "ia_ray_samples": (_I32, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "ia_contract": (_I32, [_P, _I64, _F, _I32, _P, _P]),
    "ia_normalize3_fwd": (_I32, [_P, _I64, _F, _P, _P]),
    "ia_normalize3_bwd": (_I32, [_P, _P, _I64, _F, _P, _P]),
    "ia_occ_workspace_bytes": (_I64, [_I64]),
    "ia_occ_update": (_I32, [_P, _P, _I64, _P, _I64, _F, _F, _P, _P, _P, _P]),
    "ia_occ_pack": (_I32, [_P, _I64, _P, _P]),
    "ia_aabb": (_I32, [_P, _P, _I64, C.POINTER(C.c_float * 6), _I32, _P, _P, _P]),
    "ia_march_count": (_I32, [_P, _P, _P, _P, _I64, C.POINTER(GridDesc), _P, _F, _F, _P, _P]),
    "ia_march_scan_workspace_bytes": (_I64, [_I64]),
    "ia_march_scan": (_I32, [_P, _I64, _P, _P, _P, _P]),
    "ia_march_total": (_I32, [_P, C.POINTER(C.c_int64), _P]),
    "ia_march_write": (_I32, [_P, _P, _P, _P, _I64, C.POINTER(GridDesc), _P, _F, _F, _P, _P, _P, _P, _P]),
    "ia_march_pair": (_I32, [_P, _P, _I64, C.POINTER(MarchSet), C.POINTER(MarchSet), _I32, _P]),
    "ia_march_totals": (_I32, [_P, _I32, C.POINTER(C.c_int64 * 2), _P]),
    "ia_visibility": (_I32, [_P, _P, _I64, _F, _F, _P, _P]),
    "ia_prune_count": (_I32, [_P, _P, _P, _P, _I64, _F, _F, _P, _P, _P]),
    "ia_prune_write": (_I32, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P]),
    "ia_composite_fwd": (_I32, [C.POINTER(CompositeArgs), _P, _P, _P, _P, _P, _P, _P, _P]),
    "ia_composite_bwd": (_I32, [C.POINTER(CompositeArgs), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ia_ray_mix_fwd": (_I32, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "ia_ray_mix_bwd": (_I32, [_P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "ia_neus_losses_workspace_bytes": (_I64, []),
    "ia_neus_losses_fwd": (_I32, [C.POINTER(LossArgs), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ia_neus_losses_bwd": (_I32, [C.POINTER(LossArgs), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ia_point_losses_fwd": (_I32, [_P, _P, _P, _P, _I64, _F, _F, _P, _P, _P]),
    "ia_point_losses_bwd": (_I32, [_P, _P, _P, _I64, _F, _F, _P, _P, _P, _P, _P]),
    "ia_adamw_step": (_I32, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I32, _F, _P]),
    "ia_l2_persist": (_I32, [_P, _I64, _F, C.POINTER(C.c_int64 * 3), _P]),
    "ia_debug_sector_gather": (_I32, [_P, _I64, _I64, _I32, _P, _P]),
    "ia_debug_hashgrid_fwd_generic": (_I32, [_I32]),
    "ia_debug_tc_timing": (_I32, [_I32, C.POINTER(C.c_ulonglong)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libia_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"instant_angelo_b200: CUDA library {LIB_PATH} is missing. Build it with "
                f"`python -m instant_angelo_b200.build` (or __graft_entry__.build()). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().ia_last_error_string().decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"instant_angelo_b200 {what} failed (status {rc}): {last_error()}")


def ptr(t) -> int | None:
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream() -> int:
    """cudaStream_t of torch's current stream on the current device.  (torch.cuda.current_stream() builds a Stream object
    through three Python layers, ~18 us a call, 80 calls a step: 1.5 ms of a launch-bound 7 ms step at 256-1024 rays.)"""
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")


def f32c(t: torch.Tensor) -> torch.Tensor:
    """contiguous fp32 view/copy."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
