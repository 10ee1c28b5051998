"""torch.autograd wrappers over the C ABI (include/ia_b200.h).  PyTorch supplies device memory,
streams and the autograd tape; every arithmetic step runs in libia_b200.so.

Each function documents the reference call it stands in for.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib as L

# kernels launched per ABI call (for bench.py's gpu_launches claim)
_LAUNCHES = {"ia_hashgrid_fwd": 1, "ia_hashgrid_fwd_grouped": 1, "ia_hashgrid_bwd": 1, "ia_hashgrid_bwd_grouped": 1, "ia_hashgrid_bwd_table": 1, "ia_hashgrid_bwd_input": 1, "ia_sh_fwd": 1,
             "ia_sh_bwd": 1, "ia_mlp_fwd": 1, "ia_mlp_bwd": 1, "ia_linear64_fwd": 1, "ia_linear64_bwd": 2, "ia_aabb": 1, "ia_march_count": 1, "ia_march_scan": 1,
             "ia_march_total": 0, "ia_march_write": 1, "ia_visibility": 1, "ia_occ_update": 4, "ia_occ_pack": 1,
             "ia_composite_fwd": 1, "ia_composite_bwd": 1, "ia_adamw_step": 1, "ia_hashgrid_plan": 0,
             "ia_sdf_taps_fused_fwd": 1, "ia_sdf_taps_fused_bwd": 2, "ia_mlp_fwd_grad": 1, "ia_mlp_fwd_grad_bwd": 4,
             "ia_colour_in_fwd": 1, "ia_colour_in_bwd": 1, "ia_march_totals": 0, "ia_l2_persist": 0}


class Profiler:
    """Optional per-call CUDA-event timing + launch counting (bench.py turns it on inside its timed region to
    get the dominant kernel's live launch durations on the launching stream)."""

    def __init__(self):
        self.enabled = False
        self.timing = False
        self.reset()

    def reset(self):
        self.launches = 0
        self.calls = {}
        self._events = []

    def summary(self):
        """name -> (calls, total_ms, total algorithmic work: bytes for gather kernels, FLOP for the MLP);
        call after torch.cuda.synchronize()."""
        out = {}
        for name, tag, a, b, work in self._events:
            key = name if tag is None else f"{name}[{tag}]"
            c, t, w = out.get(key, (0, 0.0, 0.0))
            out[key] = (c + 1, t + a.elapsed_time(b), w + work)
        return out


PROFILER = Profiler()


def _run(name: str, *args, tag=None, work: float = 0.0) -> None:
    fn = getattr(L.load(), name)
    prof = PROFILER
    if prof.enabled:
        prof.launches += _LAUNCHES.get(name, 1)
        prof.calls[name] = prof.calls.get(name, 0) + 1
        if prof.timing:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            prof._events.append((name, tag, a, b, work))
        else:
            rc = fn(*args)
    else:
        rc = fn(*args)
    if rc != 0:
        raise RuntimeError(f"instant_angelo_b200 {name} failed (status {rc}): {L.last_error()}")


# ---------------------------------------------------------------------------------------------
# hash grid  (tcnn.Encoding(HashGrid).forward, reference models/network_utils.py:57)
# ---------------------------------------------------------------------------------------------

def make_grid_plan(n_levels: int, n_features: int, log2_hashmap_size: int, base_resolution: int,
                   per_level_scale: float) -> L.GridPlan:
    plan = L.GridPlan()
    _run("ia_hashgrid_plan", n_levels, n_features, log2_hashmap_size, base_resolution,
                                      C.c_float(per_level_scale), C.byref(plan))
    return plan


class _HashGridFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, plan, active_levels, group=1, table_h=None):
        L.require_cuda(x, table)
        x = L.f32c(x)
        n = x.shape[0]
        out = torch.empty(n, plan.n_levels * plan.n_features, device=x.device, dtype=torch.float32)
        ctx.table_h = table_h
        if table_h is not None:
            _check_shadow(table, table_h)
            _run("ia_hashgrid_fwd_h", L.ptr(x), n, L.ptr(table_h), C.byref(plan), active_levels, L.ptr(out), L.stream(),
                 tag="f16 table", work=n * hashgrid_bytes_per_point(plan, active_levels, "fwd", param_bytes=2))
        elif group == 6 and _FWD_GROUPED:
            _run("ia_hashgrid_fwd_grouped", L.ptr(x), n, L.ptr(table), C.byref(plan), active_levels, group, L.ptr(out), L.stream(),
                 tag="g6", work=n * hashgrid_bytes_per_point(plan, active_levels, "fwd"))
        else:
            _run("ia_hashgrid_fwd", L.ptr(x), n, L.ptr(table), C.byref(plan), active_levels, L.ptr(out), L.stream(),
                 work=n * hashgrid_bytes_per_point(plan, active_levels, "fwd"))
        ctx.save_for_backward(x, table)
        ctx.plan, ctx.active, ctx.group = plan, active_levels, group
        ctx.sink_param = table if getattr(table, "_ia_grad_inplace", False) else None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, table = ctx.saved_tensors
        need_x, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dy = L.f32c(dy)
        n = x.shape[0]
        # table gradient: scattered (atomic adds) straight into the parameter's slice of the gradient arena when there is
        # one -- no table-sized zero-filled temporary and no autograd accumulation pass per hash-grid call
        sink = _grad_sink(ctx.sink_param) if need_t else None
        dtable = (sink if sink is not None else torch.zeros_like(table)) if need_t else None
        dx = torch.empty_like(x) if need_x else None
        if n > 0 and (need_t or need_x):
            work = (hashgrid_bytes_per_point(ctx.plan, ctx.active, "bwd_table") if need_t else 0) + \
                   (hashgrid_bytes_per_point(ctx.plan, ctx.active, "bwd_input") if need_x else 0)
            tag = ("table+input" if need_t and need_x else "table" if need_t else "input") + (f",g{ctx.group}" if ctx.group > 1 else "")
            if ctx.table_h is not None:
                _run("ia_hashgrid_bwd_h", L.ptr(x), n, L.ptr(ctx.table_h), L.ptr(dy), C.byref(ctx.plan), ctx.active, ctx.group,
                     L.ptr(dtable), L.ptr(dx), L.stream(), tag=tag + ",f16 table", work=n * work)
            elif ctx.group > 1:
                _run("ia_hashgrid_bwd_grouped", L.ptr(x), n, L.ptr(table), L.ptr(dy), C.byref(ctx.plan), ctx.active, ctx.group,
                     L.ptr(dtable), L.ptr(dx), L.stream(), tag=tag, work=n * work)
            else:
                _run("ia_hashgrid_bwd", L.ptr(x), n, L.ptr(table), L.ptr(dy), C.byref(ctx.plan), ctx.active,
                     L.ptr(dtable), L.ptr(dx), L.stream(), tag=tag, work=n * work)
        elif need_x:
            dx.zero_()
        return dx, (None if sink is not None else dtable), None, None, None, None


def _check_shadow(table: torch.Tensor, table_h: torch.Tensor) -> None:
    if table_h.dtype != torch.float16 or table_h.numel() != table.numel() or not table_h.is_contiguous() or table_h.device != table.device:
        raise ValueError("fp16 shadow table must be a contiguous float16 tensor with the table's element count on its device")


def table_to_half(table: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp16 shadow of a hash table (round to nearest even): what the gathers read under `table_precision: fp16`."""
    L.require_cuda(table)
    table = table.detach()
    if out is None:
        out = torch.empty(table.numel(), device=table.device, dtype=torch.float16)
    _check_shadow(table, out)
    _run("ia_table_to_half", L.ptr(table), table.numel(), L.ptr(out), L.stream())
    return out


# ia_hashgrid_fwd_grouped (one thread walks the six taps of a (group, level) and re-fetches corners only on a cell change) was
# measured on a B200 against the lane-pair kernel on the tap rows of a training step: 2.68 vs 1.85 ms per step (profiles/
# r02_ab_hashgrid_fwd_grouped.md) -- the cell-change branch diverges between the 16 groups of a warp, so the gathers are
# issued anyway, and the walk is a chain of dependent gathers.  Off unless IA_HASHGRID_FWD_GROUPED=1.
_FWD_GROUPED = os.environ.get("IA_HASHGRID_FWD_GROUPED", "0") not in ("0", "")
_NO_GRAD_SINK = os.environ.get("IA_NO_GRAD_SINK") is not None     # A/B switch (tools / bench experiments)


def _grad_sink(param) -> Optional[torch.Tensor]:
    """Called inside a backward: the tensor a hash-grid backward may accumulate d(table) into directly -- `param.grad` when
    `param` is a leaf that opted in (dp.ParamArena sets `_ia_grad_inplace` on the parameters it re-homes; their .grad is a
    persistent, zeroed view of the gradient arena).  Autograd then receives None for that input: the adds have already
    happened.  Not taken under create_graph (grad mode is on inside such a backward), where autograd must see the value;
    parameters that opt in must be differentiated with .backward(), not torch.autograd.grad()."""
    if param is None or not param.is_leaf or torch.is_grad_enabled() or _NO_GRAD_SINK:
        return None
    g = param.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != param.shape or g.device != param.device:
        return None
    return g


def hashgrid_bytes_per_point(plan: L.GridPlan, active_levels: int, which: str = "fwd", param_bytes: int = 4,
                             out_bytes: int = 4) -> int:
    """ALGORITHMIC bytes per point-evaluation (BASELINE.md section 3 / SURVEY.md section 8d):
    fwd = 12 + La*8*F*P + L*F*O; bwd_table the same; bwd_input = fwd + 12."""
    Lv, F = plan.n_levels, plan.n_features
    base = 12 + active_levels * 8 * F * param_bytes + Lv * F * out_bytes
    return base + (12 if which == "bwd_input" else 0)


def mlp_flops_per_row(desc: L.MlpDesc, n_out_used: int) -> int:
    """2*MAC of one forward evaluation (unpadded)."""
    din = desc.n_in0 + desc.n_in1
    macs = din * desc.width + (desc.width * desc.width if desc.n_hidden_layers == 2 else 0) + desc.width * n_out_used
    return 2 * macs


def hashgrid_encode(x: torch.Tensor, table: torch.Tensor, plan: L.GridPlan, active_levels: Optional[int] = None,
                    group: int = 1, table_h: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,3] in [0,1] -> [N, L*F]; levels >= active_levels are exact zeros (the progressive mask).
    group=6: rows come in groups of 6 spatially close points (finite-difference taps) -- a hint for the backward scatter.
    table_h: fp16 shadow of `table` (table_to_half) that the gathers read instead; gradients still go to `table`."""
    if active_levels is None:
        active_levels = plan.n_levels
    return _HashGridFn.apply(x, table, plan, int(active_levels), int(group), table_h)


class _HashGridInputGradFn(torch.autograd.Function):
    """dx = J(x; table)^T dy as a differentiable operator (the analytic SDF normal of reference models/geometry.py:214-218,
    where autograd runs tcnn's backward with create_graph=True).  Backward: d(dy) = J v (ia_hashgrid_jvp) and
    d(table) (ia_hashgrid_bwd_input_bwd_table); no gradient w.r.t. x (sample positions are not trainable)."""

    @staticmethod
    def forward(ctx, x, table, dy, plan, active_levels, table_h=None):
        L.require_cuda(x, table, dy)
        x, dy = L.f32c(x), L.f32c(dy)
        n = x.shape[0]
        dx = torch.empty(n, 3, device=x.device, dtype=torch.float32)
        ctx.table_h = table_h
        if table_h is not None:
            _check_shadow(table, table_h)
            _run("ia_hashgrid_bwd_h", L.ptr(x), n, L.ptr(table_h), L.ptr(dy), C.byref(plan), active_levels, 1, None, L.ptr(dx),
                 L.stream(), tag="analytic normal,f16 table", work=n * hashgrid_bytes_per_point(plan, active_levels, "bwd_input", 2))
        else:
            _run("ia_hashgrid_bwd_input", L.ptr(x), n, L.ptr(table), L.ptr(dy), C.byref(plan), active_levels, L.ptr(dx), L.stream(),
                 tag="analytic normal", work=n * hashgrid_bytes_per_point(plan, active_levels, "bwd_input"))
        ctx.save_for_backward(x, table, dy)
        ctx.plan, ctx.active = plan, active_levels
        ctx.sink_param = table if getattr(table, "_ia_grad_inplace", False) else None
        return dx

    @staticmethod
    def backward(ctx, v):
        x, table, dy = ctx.saved_tensors
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("second derivative of the hash grid w.r.t. the sample positions is not on the B200 path")
        v = L.f32c(v)
        n = x.shape[0]
        dtable = ddy = None
        if ctx.needs_input_grad[1]:
            sink = _grad_sink(ctx.sink_param)
            dtable = sink if sink is not None else torch.zeros_like(table)
            _run("ia_hashgrid_bwd_input_bwd_table", L.ptr(x), n, L.ptr(v), L.ptr(dy), C.byref(ctx.plan), ctx.active, L.ptr(dtable),
                 L.stream(), work=n * hashgrid_bytes_per_point(ctx.plan, ctx.active, "bwd_table"))
            if sink is not None:
                dtable = None
        if ctx.needs_input_grad[2]:
            ddy = torch.empty_like(dy)
            if ctx.table_h is not None:
                _run("ia_hashgrid_jvp_h", L.ptr(x), n, L.ptr(ctx.table_h), L.ptr(v), C.byref(ctx.plan), ctx.active, L.ptr(ddy),
                     L.stream(), work=n * hashgrid_bytes_per_point(ctx.plan, ctx.active, "fwd", 2))
            else:
                _run("ia_hashgrid_jvp", L.ptr(x), n, L.ptr(table), L.ptr(v), C.byref(ctx.plan), ctx.active, L.ptr(ddy), L.stream(),
                     work=n * hashgrid_bytes_per_point(ctx.plan, ctx.active, "fwd"))
        return None, dtable, ddy, None, None, None


def hashgrid_input_grad(x: torch.Tensor, table: torch.Tensor, dy: torch.Tensor, plan: L.GridPlan,
                        active_levels: Optional[int] = None, table_h: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[N,3] = J(x; table)^T dy, differentiable w.r.t. table and dy (see _HashGridInputGradFn)."""
    if active_levels is None:
        active_levels = plan.n_levels
    return _HashGridInputGradFn.apply(x, table, dy, plan, int(active_levels), table_h)


# ---------------------------------------------------------------------------------------------
# spherical harmonics  (tcnn.Encoding(SphericalHarmonics), reference models/texture.py:25)
# ---------------------------------------------------------------------------------------------

class _SHFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d01, degree):
        L.require_cuda(d01)
        d01 = L.f32c(d01)
        n = d01.shape[0]
        out = torch.empty(n, degree * degree, device=d01.device, dtype=torch.float32)
        _run("ia_sh_fwd", L.ptr(d01), n, degree, L.ptr(out), L.stream())
        ctx.save_for_backward(d01)
        ctx.degree = degree
        return out

    @staticmethod
    def backward(ctx, dout):
        (d01,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None
        dout = L.f32c(dout)
        dd = torch.empty_like(d01)
        _run("ia_sh_bwd", L.ptr(d01), d01.shape[0], ctx.degree, L.ptr(dout), L.ptr(dd), L.stream())
        return dd, None


def sh_encode(d01: torch.Tensor, degree: int) -> torch.Tensor:
    return _SHFn.apply(d01, int(degree))


# ---------------------------------------------------------------------------------------------
# fused MLP  (VanillaMLP.forward, reference models/network_utils.py:108-113)
# ---------------------------------------------------------------------------------------------

def make_mlp_desc(n_in0: int, n_in1: int, n_hidden_layers: int, n_out: int, hidden_act: int, in0_scale: float = 1.0,
                  in0_offset: float = 0.0, precision: int = L.IA_MLP_FP32) -> L.MlpDesc:
    return L.MlpDesc(n_in0, in0_scale, in0_offset, n_in1, n_hidden_layers, 64, n_out, hidden_act, L.IA_ACT_NONE, precision)


class _MLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in0, in1, params, desc, n_out_used):
        L.require_cuda(in1 if in1 is not None else in0, params)
        in0 = L.f32c(in0) if in0 is not None else None
        in1 = L.f32c(in1) if in1 is not None else None
        params = L.f32c(params)
        n = (in1 if in1 is not None else in0).shape[0]
        width_out = n_out_used if n_out_used > 0 else desc.width       # 0: last hidden layer ("feature mode")
        out = torch.empty(n, width_out, device=params.device, dtype=torch.float32)
        _run("ia_mlp_fwd", C.byref(desc), L.ptr(in0), L.ptr(in1), n, L.ptr(params), n_out_used, L.ptr(out),
             width_out, L.stream(), work=n * mlp_flops_per_row(desc, n_out_used),
             tag=f"{desc.n_in0 + desc.n_in1}>{n_out_used}/{desc.n_out}")
        ctx.save_for_backward(in0, in1, params)
        ctx.desc, ctx.nou = desc, n_out_used
        ctx.acc = getattr(params, "_ia_acc", None)
        return out

    @staticmethod
    def backward(ctx, dout):
        in0, in1, params = ctx.saved_tensors
        dout = L.f32c(dout)
        n = dout.shape[0]
        need0 = in0 is not None and ctx.needs_input_grad[0]
        need1 = in1 is not None and ctx.needs_input_grad[1]
        needp = ctx.needs_input_grad[2]
        d0 = torch.empty_like(in0) if need0 else None
        d1 = torch.empty_like(in1) if need1 else None
        sink = _flat_sink(ctx.acc) if needp else None
        dp = (sink if sink is not None else torch.zeros_like(params)) if needp else None
        if n > 0:
            _run("ia_mlp_bwd", C.byref(ctx.desc), L.ptr(in0), L.ptr(in1), n, L.ptr(params), L.ptr(dout), ctx.nou,
                 dout.shape[1], L.ptr(d0), L.ptr(d1), L.ptr(dp), L.stream(), work=2 * n * mlp_flops_per_row(ctx.desc, ctx.nou),
                 tag=f"{ctx.desc.n_in0 + ctx.desc.n_in1}>{ctx.nou}/{ctx.desc.n_out}")
        return d0, d1, (None if sink is not None else dp), None, None


class _MLPGradFn(torch.autograd.Function):
    """(h, g0, g1) = (last hidden layer, d out0 / d in0, d out0 / d in1) of the SDF network in one kernel, with the adjoint
    of that triple as its backward (ia_mlp_fwd_grad / ia_mlp_fwd_grad_bwd): reference models/geometry.py:206 + :214-218,
    `torch.autograd.grad(sdf, points_, create_graph=True)` followed by a backward through the resulting normals."""

    @staticmethod
    def forward(ctx, in0, in1, params, desc):
        L.require_cuda(in0, in1, params)
        in0, in1, params = L.f32c(in0), L.f32c(in1), L.f32c(params)
        n = in1.shape[0]
        h = torch.empty(n, desc.width, device=params.device, dtype=torch.float32)
        g0 = torch.empty(n, desc.n_in0, device=params.device, dtype=torch.float32)
        g1 = torch.empty(n, desc.n_in1, device=params.device, dtype=torch.float32)
        # 2 MAC per weight: forward (W0, W1) + the gradient chain (W1^T, W0^T)
        flops = 4 * n * (desc.width * (desc.n_in0 + desc.n_in1) + desc.width * desc.width)
        _run("ia_mlp_fwd_grad", C.byref(desc), L.ptr(in0), L.ptr(in1), n, L.ptr(params), L.ptr(h), L.ptr(g0), L.ptr(g1), L.stream(),
             work=flops, tag=f"{desc.n_in0 + desc.n_in1}>h+g")
        ctx.save_for_backward(in0, in1, params)
        ctx.desc = desc
        ctx.acc = getattr(params, "_ia_acc", None)
        return h, g0, g1

    @staticmethod
    def backward(ctx, dh, dg0, dg1):
        in0, in1, params = ctx.saved_tensors
        desc = ctx.desc
        n = in1.shape[0]
        dh = L.f32c(dh) if dh is not None else None
        dg0 = L.f32c(dg0) if dg0 is not None else None
        dg1 = L.f32c(dg1) if dg1 is not None else None
        need0, need1, needp = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        d0 = torch.empty_like(in0) if need0 else None
        d1 = torch.empty_like(in1) if need1 else None
        sink = _flat_sink(ctx.acc) if needp else None
        dp = (sink if sink is not None else torch.zeros_like(params)) if needp else None
        if n > 0:
            flops = 2 * n * (7 * desc.width * desc.width + 4 * desc.width * (desc.n_in0 + desc.n_in1))     # 4 + 2 W0-shaped, 4 + 3 W1-shaped GEMMs... per row
            _run("ia_mlp_fwd_grad_bwd", C.byref(desc), L.ptr(in0), L.ptr(in1), n, L.ptr(params), L.ptr(dh), L.ptr(dg0), L.ptr(dg1),
                 L.ptr(d0), L.ptr(d1), L.ptr(dp), L.stream(), work=flops, tag=f"{desc.n_in0 + desc.n_in1}>h+g")
        return d0, d1, (None if sink is not None else dp), None


def mlp_fwd_grad_supported(desc: L.MlpDesc) -> bool:
    return (desc.precision == L.IA_MLP_TC_F16 and desc.n_in0 == 3 and desc.n_in1 == 32 and desc.n_hidden_layers == 2
            and desc.width == 64 and desc.hidden_act == L.IA_ACT_SOFTPLUS100 and os.environ.get("IA_ANALYTIC_TORCH") is None)


def mlp_fwd_grad(in0: torch.Tensor, in1: torch.Tensor, params: torch.Tensor, desc: L.MlpDesc):
    """-> (h [N,64] last hidden layer, g0 [N,3] = d out0 / d in0, g1 [N,32] = d out0 / d in1); differentiable once more."""
    return _MLPGradFn.apply(in0, in1, params, desc)


class _SdfFusedFn(torch.autograd.Function):
    """network(cat[x*s+o, hashgrid(x)]) in one kernel (ia_sdf_taps_fused_fwd / _bwd): reference models/geometry.py:206, 233, 266
    `self.network(self.encoding(points))` without the [N, L*F] encoding in HBM."""

    @staticmethod
    def forward(ctx, x, table, params, desc, plan, active_levels, n_out_used, group):
        L.require_cuda(x, table, params)
        x, params = L.f32c(x), L.f32c(params)
        n = x.shape[0]
        width_out = n_out_used if n_out_used > 0 else desc.width
        out = torch.empty(n, width_out, device=x.device, dtype=torch.float32)
        _run("ia_sdf_taps_fused_fwd", C.byref(desc), C.byref(plan), active_levels, L.ptr(x), n, L.ptr(table), L.ptr(params), n_out_used,
             L.ptr(out), width_out, L.stream(), tag=f"{desc.n_in0 + desc.n_in1}>{n_out_used}/{desc.n_out}",
             work=n * mlp_flops_per_row(desc, n_out_used))
        ctx.save_for_backward(x, table, params)
        ctx.cfg = (desc, plan, active_levels, n_out_used, group)
        ctx.sink_param = table if getattr(table, "_ia_grad_inplace", False) else None
        return out

    @staticmethod
    def backward(ctx, dout):
        x, table, params = ctx.saved_tensors
        desc, plan, active, nou, group = ctx.cfg
        dout = L.f32c(dout)
        n = x.shape[0]
        need_x, need_t, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        sink = _grad_sink(ctx.sink_param) if need_t else None
        dtable = (sink if sink is not None else torch.zeros_like(table)) if need_t else None
        dp = torch.zeros_like(params) if need_p else None
        dx_enc = torch.empty_like(x) if need_x else None
        dx_dir = torch.empty_like(x) if need_x else None
        if n > 0:
            ws = torch.empty(n, desc.n_in1, device=x.device, dtype=torch.float32) if (need_t or need_x) else None
            tag = f"{desc.n_in0 + desc.n_in1}>{nou}/{desc.n_out}" + (f",g{group}" if group > 1 else "") + (",dx" if need_x else "")
            _run("ia_sdf_taps_fused_bwd", C.byref(desc), C.byref(plan), active, L.ptr(x), n, L.ptr(table), L.ptr(params), L.ptr(dout), nou,
                 dout.shape[1], group, L.ptr(dtable), L.ptr(dx_enc), L.ptr(dx_dir), L.ptr(dp), L.ptr(ws), L.stream(), tag=tag,
                 work=2 * n * mlp_flops_per_row(desc, nou))
            dx = dx_enc + dx_dir if need_x else None
        else:
            dx = torch.zeros_like(x) if need_x else None
        return dx, (None if sink is not None else dtable), dp, None, None, None, None, None


def sdf_fused(x: torch.Tensor, table: torch.Tensor, params: torch.Tensor, desc: L.MlpDesc, plan: L.GridPlan,
              active_levels: Optional[int] = None, n_out_used: Optional[int] = None, group: int = 1) -> torch.Tensor:
    """Fused hash-grid encode + tensor-core MLP on cat[x*scale+offset, encode(x)]; x [N,3] in the encoder's [0,1] coordinates.
    n_out_used as for mlp_apply (None: all outputs; wide output layers go through the last hidden layer + linear64)."""
    if active_levels is None:
        active_levels = plan.n_levels
    nou = int(desc.n_out if n_out_used is None else n_out_used)
    if nou > 8:
        h = _SdfFusedFn.apply(x, table, params, desc, plan, int(active_levels), 0, int(group))
        n_hidden = params.numel() - (desc.n_out * desc.width + desc.n_out)
        w_last = params[n_hidden:n_hidden + desc.n_out * desc.width].view(desc.n_out, desc.width)
        b_last = params[n_hidden + desc.n_out * desc.width:]
        return linear64(h, w_last[:nou], b_last[:nou])
    return _SdfFusedFn.apply(x, table, params, desc, plan, int(active_levels), nou, int(group))


def sdf_fused_enabled() -> bool:
    """Whether the module layer (network_utils.fused_encode_mlp) routes VolumeSDF through the fused kernels.  Measured on a
    B200 (profiles/r02_fused_encoder_ab.md): the in-kernel gather trades 256 B/row of HBM traffic for gather latency inside a
    kernel that runs 16-32 warps per SM instead of the stand-alone encoder's 64, and the step is 12 % SLOWER with it (30.7 vs
    27.2 ms), so the default stays the unfused pair; IA_FUSED_ENCODER=1 selects the fused path (e.g. when the [N, L*F]
    activations do not fit)."""
    return os.environ.get("IA_FUSED_ENCODER", "0") not in ("0", "")


def sdf_fused_supported(desc: L.MlpDesc, plan: L.GridPlan, needs_grad: bool) -> bool:
    """Shapes the fused kernels cover (anything else takes hashgrid_encode + mlp_apply)."""
    ok = (desc.precision == L.IA_MLP_TC_F16 and desc.n_in0 == 3 and plan.n_features == 2 and plan.n_levels % 4 == 0
          and desc.n_in1 == plan.n_levels * plan.n_features and desc.width == 64)
    if not ok:
        return False
    if desc.hidden_act == L.IA_ACT_SOFTPLUS100:
        return desc.n_hidden_layers == 2 or not needs_grad
    return False


class _WeightNormFlatFn(torch.autograd.Function):
    """Flat effective parameters [W0 b0 W1 b1 ...] of a VanillaMLP from its per-layer (weight_g | None, weight_v | weight,
    bias) tensors in one launch, with the weight-norm adjoint in backward (reference models/network_utils.py:115-134).

    Gradient plumbing without autograd accumulation passes: `acc` is a zeroed buffer of the flat vector's shape that the
    consumers of `flat` (the MLP backward kernels, which flush their weight gradients with atomic adds anyway) add into
    directly, returning None to autograd (_flat_sink); this backward then runs once every consumer is done -- autograd
    calls it with None, or with the sum of what arrived through ordinary operators (slices of `flat`) -- and adds the
    weight-norm adjoint straight into the parameters' slices of the gradient arena when they have one (_grad_sink).
    A step with 13 network evaluations and 30 small parameter tensors otherwise spends ~60 add / fill launches here."""

    @staticmethod
    def forward(ctx, has_g, acc, *tensors):
        n_layers = len(has_g)
        assert n_layers <= L.IA_WN_MAX_LAYERS
        d = L.WnDesc()
        d.n_layers = n_layers
        it = iter(tensors)
        keep, total = [], 0
        for i, hg in enumerate(has_g):
            g = L.f32c(next(it)) if hg else None
            v, b = L.f32c(next(it)), L.f32c(next(it))
            L.require_cuda(g, v, b)
            d.n_out[i], d.n_in[i] = v.shape[0], v.shape[1]
            d.g[i], d.v[i], d.b[i] = L.ptr(g), L.ptr(v), L.ptr(b)
            total += v.numel() + b.numel()
            keep += ([g] if hg else []) + [v, b]
        flat = torch.empty(total, device=tensors[0].device, dtype=torch.float32)
        _run("ia_weightnorm_flat_fwd", C.byref(d), L.ptr(flat), L.stream())
        ctx.save_for_backward(*keep)
        ctx.has_g = has_g
        ctx.acc = acc
        ctx.leaves = tensors                       # the parameter objects themselves (their .grad may be an arena slice)
        ctx.set_materialize_grads(False)
        return flat

    @staticmethod
    def backward(ctx, dflat):
        acc = ctx.acc
        if acc is not None and getattr(acc, "_ia_used", False):
            dflat = acc if dflat is None else dflat + acc
        if dflat is None:
            return (None, None) + (None,) * len(ctx.leaves)
        dflat = L.f32c(dflat)
        saved = list(ctx.saved_tensors)
        sinks = [_grad_sink(t if getattr(t, "_ia_grad_inplace", False) else None) for t in ctx.leaves]
        in_place = all(s is not None for s in sinks)
        d = L.WnDesc()
        d.n_layers = len(ctx.has_g)
        grads, k = [], 0
        for i, hg in enumerate(ctx.has_g):
            g = saved[k] if hg else None
            v, b = saved[k + hg], saved[k + hg + 1]
            if in_place:
                dg = sinks[k] if hg else None
                dv, db = sinks[k + hg], sinks[k + hg + 1]
            else:
                dg = torch.empty_like(g) if hg else None
                dv, db = torch.empty_like(v), torch.empty_like(b)
            k += 2 + hg
            d.n_out[i], d.n_in[i] = v.shape[0], v.shape[1]
            d.g[i], d.v[i], d.b[i] = L.ptr(g), L.ptr(v), L.ptr(b)
            d.dg[i], d.dv[i], d.db[i] = L.ptr(dg), L.ptr(dv), L.ptr(db)
            grads += ([dg] if hg else []) + [dv, db]
        _run("ia_weightnorm_flat_bwd_acc" if in_place else "ia_weightnorm_flat_bwd", C.byref(d), L.ptr(dflat), L.stream())
        if acc is not None and getattr(acc, "_ia_used", False):
            acc.zero_()                            # a second backward through a retained graph starts from zero again
            acc._ia_used = False
        if in_place:
            return (None, None) + (None,) * len(ctx.leaves)
        return (None, None, *grads)


def weightnorm_flat(layers) -> torch.Tensor:
    """layers: [(weight_g or None, weight_v / weight [out, in], bias [out]), ...] -> flat [sum(out*in + out)]."""
    has_g = tuple(int(g is not None) for g, _, _ in layers)
    tensors = [t for g, v, b in layers for t in ((g, v, b) if g is not None else (v, b))]
    acc = None
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors) and not _NO_GRAD_SINK:
        total = sum(v.numel() + b.numel() for _, v, b in layers)
        acc = torch.zeros(total, device=tensors[0].device, dtype=torch.float32)
    flat = _WeightNormFlatFn.apply(has_g, acc, *tensors)
    if acc is not None:
        flat._ia_acc = acc
    return flat


def _flat_sink(acc) -> Optional[torch.Tensor]:
    """Called inside a backward: the accumulator of a flat parameter vector made by weightnorm_flat() that this backward
    may add its parameter gradient into (it then returns None for that input).  Not under create_graph."""
    if acc is None or torch.is_grad_enabled() or _NO_GRAD_SINK:
        return None
    acc._ia_used = True
    return acc


class _Linear64Fn(torch.autograd.Function):
    """out = h @ W.T + b for h [N,64], W [n_out,64] (the wide output layer behind the tensor-core feature mode)."""

    @staticmethod
    def forward(ctx, h, W, b):
        L.require_cuda(h, W, b)
        h, W, b = L.f32c(h), L.f32c(W), L.f32c(b)
        n, n_out = h.shape[0], W.shape[0]
        out = torch.empty(n, n_out, device=h.device, dtype=torch.float32)
        _run("ia_linear64_fwd", L.ptr(h), n, L.ptr(W), L.ptr(b), n_out, L.ptr(out), n_out, L.stream(), work=2.0 * n * n_out * 64,
             tag=f"64>{n_out}")
        ctx.save_for_backward(h, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        h, W = ctx.saved_tensors
        dout = L.f32c(dout)
        n, n_out = h.shape[0], W.shape[0]
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dW = torch.zeros_like(W) if ctx.needs_input_grad[1] else None
        db = torch.zeros(n_out, device=h.device) if ctx.needs_input_grad[2] else None
        if dW is None and db is not None:
            dW = torch.zeros_like(W)
        _run("ia_linear64_bwd", L.ptr(h), n, L.ptr(W), L.ptr(dout), dout.shape[1], n_out, None, 0, L.ptr(dh), L.ptr(dW), L.ptr(db),
             L.stream(), work=4.0 * n * n_out * 64, tag=f"64>{n_out}")
        return dh, (dW if ctx.needs_input_grad[1] else None), db


def linear64(h: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return _Linear64Fn.apply(h, W, b)


class _SdfHeadFn(torch.autograd.Function):
    """Output layer of the SDF network fused with the assembly of the colour head's input row
    (reference models/geometry.py:206-207 + models/texture.py:24-27, 58):

        out   = h @ W.T + b                                   [N, n_feat]   (written straight into tin[:, :n_feat])
        tin   = cat[out, pts01*2-1, enc, normal]              [N, n_feat + 3 + n_enc + 3]
        sdf   = out[:, 0];  rgb_raw = out[:, 1:4]

    One buffer is written once instead of out / feature / network_inp being materialised one after the other, and in
    backward the gradients that reach out[:, 0:4] through sdf / rgb_raw are added inside the linear64 kernels."""

    @staticmethod
    def forward(ctx, h, W, b, pts01, enc, normal):
        L.require_cuda(h, W, b, pts01, enc, normal)
        h, W, b, pts01, enc, normal = L.f32c(h), L.f32c(W), L.f32c(b), L.f32c(pts01), L.f32c(enc), L.f32c(normal)
        n, n_feat, n_enc = h.shape[0], W.shape[0], enc.shape[1]
        ld = n_feat + 3 + n_enc + 3
        tin = torch.empty(n, ld, device=h.device, dtype=torch.float32)
        sdf = torch.empty(n, device=h.device, dtype=torch.float32)
        rgb_raw = torch.empty(n, 3, device=h.device, dtype=torch.float32)
        _run("ia_sdf_head_fwd", L.ptr(h), n, L.ptr(W), L.ptr(b), n_feat, L.ptr(pts01), L.ptr(enc), n_enc, L.ptr(normal), L.ptr(tin), ld,
             L.ptr(sdf), L.ptr(rgb_raw), L.stream(), work=2.0 * n * n_feat * 64, tag=f"64>{n_feat}")
        ctx.save_for_backward(h, W)
        ctx.dims = (n_feat, n_enc)
        return tin, sdf, rgb_raw

    @staticmethod
    def backward(ctx, dtin, dsdf, drgb):
        h, W = ctx.saved_tensors
        n_feat, n_enc = ctx.dims
        n = h.shape[0]
        dtin = L.f32c(dtin)
        dextra = torch.cat([dsdf.reshape(n, 1), drgb.reshape(n, 3)], dim=1)
        need = ctx.needs_input_grad
        dh = torch.empty_like(h)
        dW = torch.zeros_like(W) if (need[1] or need[2]) else None
        db = torch.zeros(n_feat, device=h.device) if need[2] else None
        dpts = torch.empty(n, 3, device=h.device) if need[3] else None
        denc = torch.empty(n, n_enc, device=h.device) if need[4] else None
        dnrm = torch.empty(n, 3, device=h.device) if need[5] else None
        _run("ia_sdf_head_bwd", L.ptr(h), n, L.ptr(W), L.ptr(dtin), dtin.shape[1], n_feat, n_enc, L.ptr(dextra), 4, L.ptr(dh), L.ptr(dW),
             L.ptr(db), L.ptr(dpts), L.ptr(denc), L.ptr(dnrm), L.stream(), work=4.0 * n * n_feat * 64, tag=f"64>{n_feat}")
        return (dh if need[0] else None), (dW if need[1] else None), db, dpts, denc, dnrm


def sdf_head(h, W, b, pts01, enc, normal):
    """-> (tin [N, n_feat+3+n_enc+3], sdf [N], rgb_raw [N,3]); see _SdfHeadFn."""
    return _SdfHeadFn.apply(h, W, b, pts01, enc, normal)


class _ColourInFn(torch.autograd.Function):
    """Input row of the colour head with the SDF network's output layer folded into the colour network's first layer
    (ia_colour_in_fwd / _bwd): tin = [h | pts01*2-1 | enc | normal | 0-pad], sdf = h.W4[0] + b4[0], rgb_raw = h W4[1:4]^T + b4[1:4]."""

    @staticmethod
    def forward(ctx, h, W4, b4, pts01, enc, normal, ld):
        L.require_cuda(h, W4, b4, pts01, enc, normal)
        h, W4, b4, pts01, enc, normal = L.f32c(h), L.f32c(W4), L.f32c(b4), L.f32c(pts01), L.f32c(enc), L.f32c(normal)
        n, n_enc = h.shape[0], enc.shape[1]
        tin = torch.empty(n, ld, device=h.device, dtype=torch.float32)
        sdf = torch.empty(n, device=h.device, dtype=torch.float32)
        rgb_raw = torch.empty(n, 3, device=h.device, dtype=torch.float32)
        _run("ia_colour_in_fwd", L.ptr(h), n, L.ptr(W4), L.ptr(b4), L.ptr(pts01), L.ptr(enc), n_enc, L.ptr(normal), L.ptr(tin), ld,
             L.ptr(sdf), L.ptr(rgb_raw), L.stream(), work=2.0 * n * 4 * 64)
        ctx.save_for_backward(h, W4)
        ctx.dims = (n_enc, ld)
        return tin, sdf, rgb_raw

    @staticmethod
    def backward(ctx, dtin, dsdf, drgb):
        h, W4 = ctx.saved_tensors
        n_enc, ld = ctx.dims
        n = h.shape[0]
        dtin = L.f32c(dtin)
        dsdf = L.f32c(dsdf) if dsdf is not None else None
        drgb = L.f32c(drgb) if drgb is not None else None
        need = ctx.needs_input_grad
        dh = torch.empty_like(h) if need[0] else None
        dW = torch.zeros_like(W4) if (need[1] or need[2]) else None
        db = torch.zeros(4, device=h.device) if need[2] else None
        dpts = torch.empty(n, 3, device=h.device) if need[3] else None
        denc = torch.empty(n, n_enc, device=h.device) if need[4] else None
        dnrm = torch.empty(n, 3, device=h.device) if need[5] else None
        if n > 0:
            _run("ia_colour_in_bwd", L.ptr(h), n, L.ptr(W4), L.ptr(dtin), ld, n_enc, L.ptr(dsdf), L.ptr(drgb), L.ptr(dh), L.ptr(dW), L.ptr(db),
                 L.ptr(dpts), L.ptr(denc), L.ptr(dnrm), L.stream(), work=4.0 * n * 4 * 64)
        elif dh is not None:
            dh.zero_()
        return dh, (dW if need[1] else None), db, dpts, denc, dnrm, None


class _FoldHeadFn(torch.autograd.Function):
    """flat_eff of texture.forward_fused_head (ia_fold_head_fwd / _bwd): the geometry output layer composed into the colour
    network's first layer, in parameter space, one launch each way.  The gradient w.r.t. the colour network's flat vector goes
    into the accumulator behind it when there is one (_flat_sink)."""

    @staticmethod
    def forward(ctx, flat, w_last, b_last, n_in, n_feat, ld):
        L.require_cuda(flat, w_last, b_last)
        acc = getattr(flat, "_ia_acc", None)
        flat, w_last, b_last = L.f32c(flat), L.f32c(w_last), L.f32c(b_last)
        n_rest = flat.numel() - 64 * n_in - 64
        out = torch.empty(64 * ld + 64 + n_rest, device=flat.device, dtype=torch.float32)
        _run("ia_fold_head_fwd", L.ptr(flat), L.ptr(w_last), L.ptr(b_last), n_in, n_feat, ld, n_rest, L.ptr(out), L.stream())
        ctx.save_for_backward(flat, w_last, b_last)
        ctx.dims, ctx.acc = (n_in, n_feat, ld, n_rest), acc
        return out

    @staticmethod
    def backward(ctx, g):
        flat, w_last, b_last = ctx.saved_tensors
        n_in, n_feat, ld, n_rest = ctx.dims
        need = ctx.needs_input_grad
        g = L.f32c(g)
        sink = _flat_sink(ctx.acc) if need[0] else None
        dflat = (sink if sink is not None else torch.zeros_like(flat)) if need[0] else None
        dwl = torch.empty_like(w_last) if need[1] else None
        dbl = torch.empty_like(b_last) if need[2] else None
        _run("ia_fold_head_bwd", L.ptr(g), L.ptr(flat), L.ptr(w_last), L.ptr(b_last), n_in, n_feat, ld, n_rest, L.ptr(dflat),
             L.ptr(dwl), L.ptr(dbl), L.stream())
        return (None if sink is not None else dflat), dwl, dbl, None, None, None


def fold_head(flat, w_last, b_last, n_in: int, n_feat: int, ld: int) -> torch.Tensor:
    """-> flat_eff [64*ld + 64 + rest]; see _FoldHeadFn."""
    return _FoldHeadFn.apply(flat, w_last, b_last, int(n_in), int(n_feat), int(ld))


def colour_in(h, W4, b4, pts01, enc, normal, ld: int):
    """-> (tin [N, ld], sdf [N], rgb_raw [N,3]); see _ColourInFn."""
    return _ColourInFn.apply(h, W4, b4, pts01, enc, normal, int(ld))


def mlp_apply(in0: Optional[torch.Tensor], in1: Optional[torch.Tensor], params: torch.Tensor, desc: L.MlpDesc,
              n_out_used: Optional[int] = None) -> torch.Tensor:
    """Network on cat[in0*scale+offset, in1]; returns the first n_out_used outputs (n_out_used=None: all;
    n_out_used=0, tensor-core precision only: the last hidden layer's activations [N, 64])."""
    nou = int(desc.n_out if n_out_used is None else n_out_used)
    if desc.precision == L.IA_MLP_TC_F16 and nou > 8:
        # wide output layer (the 65-feature centre evaluation): hidden layers on the tensor-core kernel (feature mode), the
        # 64 -> nou projection by the streaming fp32 kernels of linear64.cu
        h = _MLPFn.apply(in0, in1, params, desc, 0)
        n_hidden = params.numel() - (desc.n_out * desc.width + desc.n_out)
        w_last = params[n_hidden:n_hidden + desc.n_out * desc.width].view(desc.n_out, desc.width)
        b_last = params[n_hidden + desc.n_out * desc.width:]
        return linear64(h, w_last[:nou], b_last[:nou])
    return _MLPFn.apply(in0, in1, params, desc, nou)


# ---------------------------------------------------------------------------------------------
# finite-difference / curvature stages of VolumeSDF.forward (reference models/geometry.py:219-275)
# ---------------------------------------------------------------------------------------------

class _FDTapsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, base, eps, radius):
        L.require_cuda(base)
        base = L.f32c(base)
        n = base.shape[0]
        taps = torch.empty(n, 6, 3, device=base.device, dtype=torch.float32)
        _run("ia_fd_taps_fwd", L.ptr(base), n, C.c_float(eps), C.c_float(radius), L.ptr(taps), L.stream())
        ctx.save_for_backward(base)
        ctx.eps, ctx.radius = eps, radius
        return taps

    @staticmethod
    def backward(ctx, dtaps):
        (base,) = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None
        dtaps = L.f32c(dtaps)
        dbase = torch.empty_like(base)
        _run("ia_fd_taps_bwd", L.ptr(base), base.shape[0], C.c_float(ctx.eps), C.c_float(ctx.radius), L.ptr(dtaps), L.ptr(dbase),
             L.stream())
        return dbase, None, None


def fd_taps(base: torch.Tensor, eps: float, radius: float) -> torch.Tensor:
    """[S,3] -> [S,6,3] normalised tap positions: ((base + eps*e_k).clamp(-r, r) + r) / (2r)."""
    return _FDTapsFn.apply(base, float(eps), float(radius))


class _FDGradFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf6, eps):
        L.require_cuda(sdf6)
        sdf6 = L.f32c(sdf6)
        n = sdf6.shape[0]
        grad = torch.empty(n, 3, device=sdf6.device, dtype=torch.float32)
        _run("ia_fd_grad_fwd", L.ptr(sdf6), n, C.c_float(eps), L.ptr(grad), L.stream())
        ctx.eps = eps
        return grad

    @staticmethod
    def backward(ctx, dgrad):
        dgrad = L.f32c(dgrad)
        n = dgrad.shape[0]
        d6 = torch.empty(n, 6, device=dgrad.device, dtype=torch.float32)
        _run("ia_fd_grad_bwd", L.ptr(dgrad), n, C.c_float(ctx.eps), L.ptr(d6), L.stream())
        return d6, None


def fd_grad(sdf6: torch.Tensor, eps: float) -> torch.Tensor:
    """[S,6] tap SDFs -> [S,3] central differences 0.5*(s+ - s-)/eps."""
    return _FDGradFn.apply(sdf6, float(eps))


class _CurvShiftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grad, rnd, pts01, eps):
        L.require_cuda(grad, rnd, pts01)
        grad, rnd, pts01 = L.f32c(grad), L.f32c(rnd), L.f32c(pts01)
        n = grad.shape[0]
        normals = torch.empty_like(grad)
        shifted = torch.empty_like(grad)
        _run("ia_curv_shift_fwd", L.ptr(grad), L.ptr(rnd), L.ptr(pts01), n, C.c_float(eps), L.ptr(normals), L.ptr(shifted), L.stream())
        ctx.save_for_backward(grad, rnd)
        ctx.eps = eps
        return normals, shifted

    @staticmethod
    def backward(ctx, dnormals, dshifted):
        grad, rnd = ctx.saved_tensors
        if not ctx.needs_input_grad[0]:
            return None, None, None, None
        dgrad = torch.empty_like(grad)
        _run("ia_curv_shift_bwd", L.ptr(grad), L.ptr(rnd), grad.shape[0], C.c_float(ctx.eps), L.ptr(L.f32c(dnormals)),
             L.ptr(L.f32c(dshifted)), L.ptr(dgrad), L.stream())
        return dgrad, None, None, None


def curv_shift(grad: torch.Tensor, rnd: torch.Tensor, pts01: torch.Tensor, eps: float):
    """normals = normalize(grad); shifted = pts01 + cross(normals, normalize(rnd)) * eps."""
    return _CurvShiftFn.apply(grad, rnd, pts01, float(eps))


class _CurvAngleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, normals, gshift):
        L.require_cuda(normals, gshift)
        normals, gshift = L.f32c(normals), L.f32c(gshift)
        n = normals.shape[0]
        lap = torch.empty(n, 1, device=normals.device, dtype=torch.float32)
        _run("ia_curv_angle_fwd", L.ptr(normals), L.ptr(gshift), n, L.ptr(lap), L.stream())
        ctx.save_for_backward(normals, gshift)
        return lap

    @staticmethod
    def backward(ctx, dlap):
        normals, gshift = ctx.saved_tensors
        dn, dg = torch.empty_like(normals), torch.empty_like(gshift)
        _run("ia_curv_angle_bwd", L.ptr(normals), L.ptr(gshift), normals.shape[0], L.ptr(L.f32c(dlap)), L.ptr(dn), L.ptr(dg), L.stream())
        return dn, dg


def curv_angle(normals: torch.Tensor, gshift: torch.Tensor) -> torch.Tensor:
    """laplace[S,1] = acos(clamp(normals . normalize(gshift), -1+1e-6, 1-1e-6)) / pi."""
    return _CurvAngleFn.apply(normals, gshift)


# ---------------------------------------------------------------------------------------------
# marching  (nerfacc.ray_aabb_intersect / ray_marching kernels, reference models/neus.py:153,159-169,209-220)
# ---------------------------------------------------------------------------------------------

def make_grid_desc(roi, res, contraction: int) -> L.GridDesc:
    g = L.GridDesc()
    for i in range(6):
        g.roi[i] = float(roi[i])
    for i in range(3):
        g.res[i] = int(res[i])
    g.contraction = int(contraction)
    return g


@torch.no_grad()
def aabb_intersect(rays_o: torch.Tensor, rays_d: torch.Tensor, aabb6, clamp_zero: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    L.require_cuda(rays_o, rays_d)
    rays_o, rays_d = L.f32c(rays_o), L.f32c(rays_d)
    n = rays_o.shape[0]
    t_min = torch.empty(n, device=rays_o.device, dtype=torch.float32)
    t_max = torch.empty_like(t_min)
    bb = (C.c_float * 6)(*[float(v) for v in aabb6])
    _run("ia_aabb", L.ptr(rays_o), L.ptr(rays_d), n, C.byref(bb), int(clamp_zero), L.ptr(t_min), L.ptr(t_max),
                             L.stream())
    return t_min, t_max


@torch.no_grad()
def march(rays_o, rays_d, t_min, t_max, grid: L.GridDesc, bitfield: Optional[torch.Tensor], step_size: float,
          cone_angle: float):
    """Two-pass marching.  Returns (packed_info [R,2] i32, ray_indices [S] i32, t_starts [S], t_ends [S])."""
    L.require_cuda(rays_o, rays_d, t_min, t_max)
    rays_o, rays_d, t_min, t_max = L.f32c(rays_o), L.f32c(rays_d), L.f32c(t_min), L.f32c(t_max)
    n = rays_o.shape[0]
    dev = rays_o.device
    num = torch.empty(n, device=dev, dtype=torch.int32)
    packed = torch.empty(n, 2, device=dev, dtype=torch.int32)
    total_dev = torch.zeros(1, device=dev, dtype=torch.int64)
    s = L.stream()
    _run("ia_march_count", L.ptr(rays_o), L.ptr(rays_d), L.ptr(t_min), L.ptr(t_max), n, C.byref(grid), L.ptr(bitfield),
                               C.c_float(step_size), C.c_float(cone_angle), L.ptr(num), s)
    _run("ia_march_scan", L.ptr(num), n, L.ptr(packed), L.ptr(total_dev), None, s)
    total = C.c_int64(0)
    _run("ia_march_total", L.ptr(total_dev), C.byref(total), s)
    S = int(total.value)
    ray_indices = torch.empty(S, device=dev, dtype=torch.int32)
    t_starts = torch.empty(S, device=dev, dtype=torch.float32)
    t_ends = torch.empty(S, device=dev, dtype=torch.float32)
    if S > 0:
        _run("ia_march_write", L.ptr(rays_o), L.ptr(rays_d), L.ptr(t_min), L.ptr(t_max), n, C.byref(grid),
                                   L.ptr(bitfield), C.c_float(step_size), C.c_float(cone_angle), L.ptr(packed),
                                   L.ptr(ray_indices), L.ptr(t_starts), L.ptr(t_ends), s)
    return packed, ray_indices, t_starts, t_ends


@torch.no_grad()
def march_pair(rays_o, rays_d, set_a, set_b):
    """Two marches of the same rays (set = (t_min, t_max, grid desc, bitfield, step_size, cone_angle)) with one count launch,
    one read-back of both totals and one write launch.  Returns ((packed_info, ray_indices, t_starts, t_ends), (...)) exactly
    as two march() calls would."""
    L.require_cuda(rays_o, rays_d, set_a[0], set_a[1], set_b[0], set_b[1])
    rays_o, rays_d = L.f32c(rays_o), L.f32c(rays_d)
    n, dev, s = rays_o.shape[0], rays_o.device, L.stream()
    totals_dev = torch.zeros(2, device=dev, dtype=torch.int64)
    sets, keep = [], []
    for k, (t_min, t_max, grid, bitfield, step, cone) in enumerate((set_a, set_b)):
        t_min, t_max = L.f32c(t_min), L.f32c(t_max)
        num = torch.empty(n, device=dev, dtype=torch.int32)
        packed = torch.empty(n, 2, device=dev, dtype=torch.int32)
        ms = L.MarchSet(L.ptr(t_min), L.ptr(t_max), C.pointer(grid), L.ptr(bitfield), float(step), float(cone), L.ptr(packed), L.ptr(num),
                        None, None, None)
        sets.append(ms)
        keep.append((t_min, t_max, num, packed, grid, bitfield))
    _run("ia_march_pair", L.ptr(rays_o), L.ptr(rays_d), n, C.byref(sets[0]), C.byref(sets[1]), 0, s)
    for k in range(2):
        _run("ia_march_scan", L.ptr(keep[k][2]), n, L.ptr(keep[k][3]), totals_dev.data_ptr() + 8 * k, None, s)
    totals = (C.c_int64 * 2)()
    _run("ia_march_totals", L.ptr(totals_dev), 2, C.byref(totals), s)
    outs = []
    for k in range(2):
        S = int(totals[k])
        ri = torch.empty(S, device=dev, dtype=torch.int32)
        t0 = torch.empty(S, device=dev, dtype=torch.float32)
        t1 = torch.empty(S, device=dev, dtype=torch.float32)
        sets[k].ray_indices, sets[k].t_starts, sets[k].t_ends = L.ptr(ri), L.ptr(t0), L.ptr(t1)
        outs.append((keep[k][3], ri, t0, t1))
    if int(totals[0]) + int(totals[1]) > 0:
        _run("ia_march_pair", L.ptr(rays_o), L.ptr(rays_d), n, C.byref(sets[0]), C.byref(sets[1]), 1, s)
    return outs[0], outs[1]


@torch.no_grad()
def prune_samples(sigmas, t_starts, t_ends, packed_info, early_stop_eps: float, alpha_thre: float):
    """nerfacc.ray_marching's sigma_fn pruning and the compaction after it (reference models/neus.py:144-149, 159-169) as
    count -> scan -> write: (ray_indices [S'] i32, t_starts [S',1], t_ends [S',1], packed_info [R,2] i32) of the kept samples."""
    L.require_cuda(sigmas, t_starts, t_ends, packed_info)
    sigmas, t0, t1 = L.f32c(sigmas.reshape(-1)), L.f32c(t_starts.reshape(-1)), L.f32c(t_ends.reshape(-1))
    packed_info = packed_info.contiguous()
    n, S = packed_info.shape[0], sigmas.shape[0]
    dev = sigmas.device
    vis = torch.empty(S, device=dev, dtype=torch.uint8)
    num = torch.empty(n, device=dev, dtype=torch.int32)
    packed = torch.empty(n, 2, device=dev, dtype=torch.int32)
    total_dev = torch.zeros(1, device=dev, dtype=torch.int64)
    s = L.stream()
    _run("ia_prune_count", L.ptr(sigmas), L.ptr(t0), L.ptr(t1), L.ptr(packed_info), n, C.c_float(early_stop_eps), C.c_float(alpha_thre),
         L.ptr(vis), L.ptr(num), s)
    _run("ia_march_scan", L.ptr(num), n, L.ptr(packed), L.ptr(total_dev), None, s)
    total = C.c_int64(0)
    _run("ia_march_total", L.ptr(total_dev), C.byref(total), s)
    K = int(total.value)
    ri = torch.empty(K, device=dev, dtype=torch.int32)
    k0 = torch.empty(K, 1, device=dev, dtype=torch.float32)
    k1 = torch.empty(K, 1, device=dev, dtype=torch.float32)
    if K > 0:
        _run("ia_prune_write", L.ptr(vis), L.ptr(packed_info), L.ptr(packed), L.ptr(t0), L.ptr(t1), n, L.ptr(ri), L.ptr(k0), L.ptr(k1), s)
    return ri, k0, k1, packed


@torch.no_grad()
def visibility(alphas: torch.Tensor, packed_info: torch.Tensor, early_stop_eps: float, alpha_thre: float) -> torch.Tensor:
    L.require_cuda(alphas, packed_info)
    alphas = L.f32c(alphas.reshape(-1))
    vis = torch.empty(alphas.shape[0], device=alphas.device, dtype=torch.uint8)
    _run("ia_visibility", L.ptr(alphas), L.ptr(packed_info), packed_info.shape[0], C.c_float(early_stop_eps),
                                   C.c_float(alpha_thre), L.ptr(vis), L.stream())
    return vis.bool()


@torch.no_grad()
def occ_update(idx: Optional[torch.Tensor], occ: torch.Tensor, occs: torch.Tensor, ema_decay: float, occ_thre: float,
               binary_u8: torch.Tensor, bitfield: torch.Tensor, workspace: torch.Tensor) -> None:
    L.require_cuda(occ, occs)
    occ = L.f32c(occ.reshape(-1))
    n = occ.shape[0]
    _run("ia_occ_update", L.ptr(idx), L.ptr(occ), n, L.ptr(occs), occs.numel(), C.c_float(ema_decay),
                                   C.c_float(occ_thre), L.ptr(binary_u8), L.ptr(bitfield), L.ptr(workspace), L.stream())


@torch.no_grad()
def occ_pack(binary_u8: torch.Tensor, bitfield: torch.Tensor) -> None:
    _run("ia_occ_pack", L.ptr(binary_u8), binary_u8.numel(), L.ptr(bitfield), L.stream())


# ---------------------------------------------------------------------------------------------
# compositing  (get_alpha + render_weight_from_alpha/_density + accumulate_along_rays,
#               reference models/neus.py:117-139, 181-184, 234-239)
# ---------------------------------------------------------------------------------------------

class _CompositeFn(torch.autograd.Function):
    """inputs: mode, packed_info, n_rays, cos_anneal, then tensors
       a (alpha | sdf | sigma), normal, dirs, dists | t_starts, t_ends, inv_s, t_mid, rgb, nrm
       outputs: weights[S], opacity[R], depth[R], comp_rgb[R,3], comp_nrm[R,3], alpha[S]"""

    @staticmethod
    def forward(ctx, mode, packed_info, cos_anneal, a, normal, dirs, dists, t_starts, t_ends, inv_s, t_mid, rgb, nrm):
        L.require_cuda(a, packed_info)
        R, S = packed_info.shape[0], a.shape[0]
        dev = a.device
        cz = lambda t: L.f32c(t) if t is not None else None
        a, normal, dirs, dists, t_starts, t_ends, inv_s, t_mid, rgb, nrm = map(
            cz, (a, normal, dirs, dists, t_starts, t_ends, inv_s, t_mid, rgb, nrm))
        args = L.CompositeArgs()
        args.mode, args.n_rays, args.n_samples = mode, R, S
        args.packed_info = L.ptr(packed_info)
        if mode == L.IA_ALPHA_GIVEN:
            args.alpha_in = L.ptr(a)
        elif mode == L.IA_ALPHA_NEUS:
            args.sdf, args.normal, args.dirs, args.dists, args.inv_s = L.ptr(a), L.ptr(normal), L.ptr(dirs), L.ptr(dists), L.ptr(inv_s)
            args.cos_anneal_ratio = float(cos_anneal)
        else:
            args.sigma, args.t_starts, args.t_ends = L.ptr(a), L.ptr(t_starts), L.ptr(t_ends)
        args.t_mid, args.rgb, args.nrm = L.ptr(t_mid), L.ptr(rgb), L.ptr(nrm)
        alpha = torch.empty(S, device=dev)
        trans = torch.empty(S, device=dev)
        weights = torch.empty(S, device=dev)
        opacity = torch.empty(R, device=dev)
        depth = torch.empty(R, device=dev) if t_mid is not None else None
        comp_rgb = torch.empty(R, 3, device=dev) if rgb is not None else None
        comp_nrm = torch.empty(R, 3, device=dev) if nrm is not None else None
        _run("ia_composite_fwd", C.byref(args), L.ptr(alpha), L.ptr(trans), L.ptr(weights), L.ptr(opacity),
                                          L.ptr(depth), L.ptr(comp_rgb), L.ptr(comp_nrm), L.stream())
        ctx.args = args
        ctx.keep = (packed_info, a, normal, dirs, dists, t_starts, t_ends, inv_s, t_mid, rgb, nrm, alpha, trans)
        ctx.mode = mode
        ctx.mark_non_differentiable(alpha)
        z = lambda t, *shape: t if t is not None else torch.zeros(*shape, device=dev)
        return weights, opacity, z(depth, R), z(comp_rgb, R, 3), z(comp_nrm, R, 3), alpha

    @staticmethod
    def backward(ctx, g_w, g_o, g_d, g_c, g_n, _g_alpha):
        (packed_info, a, normal, dirs, dists, t_starts, t_ends, inv_s, t_mid, rgb, nrm, alpha, trans) = ctx.keep
        dev = a.device
        S = a.shape[0]
        cz = lambda t: L.f32c(t) if t is not None else None
        g_w, g_o, g_d, g_c, g_n = map(cz, (g_w, g_o, g_d, g_c, g_n))
        mode = ctx.mode
        d_a = torch.empty(S, device=dev)
        d_normal = torch.empty(S, 3, device=dev) if mode == L.IA_ALPHA_NEUS else None
        d_inv_s = torch.zeros(1, device=dev) if mode == L.IA_ALPHA_NEUS else None
        d_rgb = torch.empty(S, 3, device=dev) if rgb is not None else None
        d_nrm = torch.empty(S, 3, device=dev) if nrm is not None else None
        _run("ia_composite_bwd", 
            C.byref(ctx.args), L.ptr(alpha), L.ptr(trans), L.ptr(g_w), L.ptr(g_o), L.ptr(g_d), L.ptr(g_c), L.ptr(g_n),
            L.ptr(d_a) if mode == L.IA_ALPHA_GIVEN else None, L.ptr(d_a) if mode == L.IA_ALPHA_NEUS else None,
            L.ptr(d_normal), L.ptr(d_inv_s), L.ptr(d_a) if mode == L.IA_ALPHA_DENSITY else None, L.ptr(d_rgb), L.ptr(d_nrm),
            L.stream())
        if d_inv_s is not None and inv_s is not None:
            d_inv_s = d_inv_s.reshape(inv_s.shape)
        return (None, None, None, d_a, d_normal, None, None, None, None, d_inv_s, None, d_rgb, d_nrm)


def composite_neus(sdf, normal, dirs, dists, inv_s, cos_anneal: float, packed_info, t_mid=None, rgb=None, nrm=None):
    """NeuS alpha + per-ray compositing.  inv_s: 1-element CUDA tensor (already clipped).
    Returns (weights[S], opacity[R], depth[R], comp_rgb[R,3], comp_normal_sum[R,3], alpha[S])."""
    return _CompositeFn.apply(L.IA_ALPHA_NEUS, packed_info, cos_anneal, sdf, normal, dirs, dists, None, None, inv_s, t_mid, rgb, nrm)


def composite_density(sigma, t_starts, t_ends, packed_info, t_mid=None, rgb=None):
    return _CompositeFn.apply(L.IA_ALPHA_DENSITY, packed_info, 0.0, sigma, None, None, None, t_starts, t_ends, None, t_mid, rgb, None)


def composite_alpha(alpha, packed_info, t_mid=None, rgb=None, nrm=None):
    return _CompositeFn.apply(L.IA_ALPHA_GIVEN, packed_info, 0.0, alpha, None, None, None, None, None, None, t_mid, rgb, nrm)


# ---------------------------------------------------------------------------------------------
# optimizer
# ---------------------------------------------------------------------------------------------

_PARAM_EPOCH = 0      # bumped by every ia_adamw_step: raw-pointer parameter writes that torch's version counters do not see


def param_epoch() -> int:
    """Number of fused optimizer steps issued by this process (Encoding.shadow() re-derives its fp16 copy when it moves)."""
    return _PARAM_EPOCH


@torch.no_grad()
def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0) -> None:
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1
    _run("ia_adamw_step", L.ptr(param), L.ptr(grad), L.ptr(exp_avg), L.ptr(exp_avg_sq), param.numel(),
                                   C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                   C.c_float(weight_decay), int(step), C.c_float(grad_scale), L.stream())


# ---------------------------------------------------------------------------------------------
# L2 residency of the hash tables
# ---------------------------------------------------------------------------------------------

def l2_persist(t: Optional[torch.Tensor], hit_ratio: float = 1.0) -> dict:
    """Persisting access-policy window over `t` on the current stream (None removes it): the hash tables stay in the L2's
    set-aside across the [N, L*F] activation passes that would otherwise evict them (ia_l2_persist)."""
    info = (C.c_int64 * 3)()
    L.require_cuda(t)
    if t is None:
        _run("ia_l2_persist", None, 0, C.c_float(0.0), C.byref(info), L.stream())
    else:
        _run("ia_l2_persist", L.ptr(t), t.numel() * t.element_size(), C.c_float(hit_ratio), C.byref(info), L.stream())
    return {"l2_bytes": int(info[0]), "set_aside_bytes": int(info[1]), "window_bytes": int(info[2])}


# ---------------------------------------------------------------------------------------------
# loss terms of the training step  (reference systems/neus.py:132-160)
# ---------------------------------------------------------------------------------------------

LOSS_TERMS = ("rgb_mse", "rgb_l1", "eikonal", "mask", "opaque", "sparsity", "curvature")


class _NeusLossesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, args, comp_rgb, rgb_gt, valid, opacity, fg_mask, sdf_grad, sdf, laplace):
        L.require_cuda(comp_rgb, rgb_gt, valid, opacity, sdf_grad, sdf)
        comp_rgb, rgb_gt, opacity = L.f32c(comp_rgb), L.f32c(rgb_gt), L.f32c(opacity)
        sdf_grad, sdf = L.f32c(sdf_grad), L.f32c(sdf)
        fg_mask = L.f32c(fg_mask) if fg_mask is not None else None
        laplace = L.f32c(laplace) if laplace is not None else None
        valid = valid.contiguous()
        if valid.dtype not in (torch.bool, torch.uint8):
            raise ValueError("valid must be a bool / uint8 tensor")
        dev = comp_rgb.device
        ws = torch.empty(int(L.load().ia_neus_losses_workspace_bytes()), device=dev, dtype=torch.uint8)
        out = torch.empty(9, device=dev, dtype=torch.float32)            # terms[8] | loss
        _run("ia_neus_losses_fwd", C.byref(args), L.ptr(comp_rgb), L.ptr(rgb_gt), L.ptr(valid), L.ptr(opacity), L.ptr(fg_mask),
             L.ptr(sdf_grad), L.ptr(sdf), L.ptr(laplace), L.ptr(ws), L.ptr(out), out.data_ptr() + 32, L.stream())
        ctx.save_for_backward(comp_rgb, rgb_gt, valid, opacity, fg_mask, sdf_grad, sdf, laplace, out)
        ctx.args = args
        loss, terms = out[8], out[:8]
        ctx.mark_non_differentiable(terms)
        return loss, terms

    @staticmethod
    def backward(ctx, dloss, _dterms):
        comp_rgb, rgb_gt, valid, opacity, fg_mask, sdf_grad, sdf, laplace, out = ctx.saved_tensors
        need = ctx.needs_input_grad
        dloss = L.f32c(dloss).reshape(1)
        d_rgb = torch.empty_like(comp_rgb) if need[1] else None
        d_op = torch.empty_like(opacity) if need[4] else None
        d_g = torch.empty_like(sdf_grad) if need[6] else None
        d_s = torch.empty_like(sdf) if need[7] else None
        d_l = torch.empty_like(laplace) if (laplace is not None and need[8]) else None
        _run("ia_neus_losses_bwd", C.byref(ctx.args), L.ptr(comp_rgb), L.ptr(rgb_gt), L.ptr(valid), L.ptr(opacity), L.ptr(fg_mask),
             L.ptr(sdf_grad), L.ptr(sdf), L.ptr(laplace), L.ptr(out), L.ptr(dloss), L.ptr(d_rgb), L.ptr(d_op), L.ptr(d_g), L.ptr(d_s),
             L.ptr(d_l), L.stream())
        return None, d_rgb, None, None, d_op, None, d_g, d_s, d_l


def neus_losses(comp_rgb, rgb_gt, valid, opacity, fg_mask, sdf_grad, sdf, laplace, lambdas: dict, sparsity_scale: float):
    """The per-ray / per-sample loss terms of reference systems/neus.py:132-160 and their weighted sum in one launch (one more
    in backward).  comp_rgb [R,3], rgb_gt [R,3], valid [R] bool, opacity [R], fg_mask [R] or None, sdf_grad [S,3], sdf [S],
    laplace [S] or None; lambdas: {'rgb_mse', 'rgb_l1', 'eikonal', 'mask', 'opaque', 'sparsity', 'curvature'} -> float.
    Returns (loss scalar, {term: 0-d tensor}); the terms are values for logging (not differentiable), the scalar is."""
    n_rays, n_samples = comp_rgb.shape[0], sdf.reshape(-1).shape[0]
    args = L.LossArgs(n_rays, n_samples, float(lambdas["rgb_mse"]), float(lambdas["rgb_l1"]), float(lambdas["eikonal"]),
                      float(lambdas.get("mask", 0.0)), float(lambdas["opaque"]), float(lambdas["sparsity"]),
                      float(lambdas.get("curvature", 0.0)), float(sparsity_scale))
    loss, terms = _NeusLossesFn.apply(args, comp_rgb.reshape(-1, 3), rgb_gt.reshape(-1, 3), valid.reshape(-1), opacity.reshape(-1),
                                      None if fg_mask is None else fg_mask.reshape(-1), sdf_grad.reshape(-1, 3), sdf.reshape(-1),
                                      None if laplace is None else laplace.reshape(-1))
    named = {k: terms[i] for i, k in enumerate(LOSS_TERMS)}
    if fg_mask is None:
        named.pop("mask")
    if laplace is None or float(lambdas.get("curvature", 0.0)) <= 0:
        named.pop("curvature")
    return loss, named


# ---------------------------------------------------------------------------------------------
# glue between the marcher and the networks  (reference models/neus.py:153-157, 218-223, 229)
# ---------------------------------------------------------------------------------------------

@torch.no_grad()
def ray_samples(rays_o, rays_d, ray_indices, t_starts, t_ends, want_dirs=True, want_mid=True, want_dists=True):
    """(positions [S,3], t_dirs [S,3], midpoints [S,1], dists [S,1]) of marched samples in one launch; rays carry no gradient
    on the training path (callers with differentiable rays keep the tensor expression)."""
    L.require_cuda(rays_o, rays_d, ray_indices, t_starts, t_ends)
    rays_o, rays_d, t_starts, t_ends = L.f32c(rays_o), L.f32c(rays_d), L.f32c(t_starts), L.f32c(t_ends)
    ri = ray_indices.contiguous()
    if ri.dtype != torch.int32:
        ri = ri.int()
    n = ri.shape[0]
    dev = rays_o.device
    pos = torch.empty(n, 3, device=dev, dtype=torch.float32)
    dirs = torch.empty(n, 3, device=dev, dtype=torch.float32) if want_dirs else None
    mid = torch.empty(n, 1, device=dev, dtype=torch.float32) if want_mid else None
    dist = torch.empty(n, 1, device=dev, dtype=torch.float32) if want_dists else None
    _run("ia_ray_samples", L.ptr(rays_o), L.ptr(rays_d), L.ptr(ri), L.ptr(t_starts), L.ptr(t_ends), n, L.ptr(pos), L.ptr(dirs),
         L.ptr(mid), L.ptr(dist), L.stream())
    return pos, dirs, mid, dist


class _Normalize3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        L.require_cuda(x)
        x = L.f32c(x)
        out = torch.empty_like(x)
        _run("ia_normalize3_fwd", L.ptr(x), x.shape[0], C.c_float(eps), L.ptr(out), L.stream())
        ctx.save_for_backward(x)
        ctx.eps = eps
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = L.f32c(g)
        dx = torch.empty_like(x)
        _run("ia_normalize3_bwd", L.ptr(x), L.ptr(g), x.shape[0], C.c_float(ctx.eps), L.ptr(dx), L.stream())
        return dx, None


def normalize3(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, p=2, dim=-1, eps) for [..., 3] CUDA tensors: one launch forward, one backward."""
    return _Normalize3Fn.apply(x.reshape(-1, 3), float(eps)).view(x.shape)


@torch.no_grad()
def contract(x: torch.Tensor, radius: float, contraction_type: int) -> torch.Tensor:
    """contract_to_unisphere (reference models/geometry.py:19-31) in one launch, for positions without gradient."""
    L.require_cuda(x)
    x = L.f32c(x)
    out = torch.empty_like(x)
    _run("ia_contract", L.ptr(x), x.numel() // 3, C.c_float(radius), int(contraction_type), L.ptr(out), L.stream())
    return out


class _RayMixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, op, rgb_bg_raw, op_bg, bg_color):
        L.require_cuda(rgb, op, rgb_bg_raw, op_bg, bg_color)
        rgb, op, rgb_bg_raw, op_bg, bg_color = L.f32c(rgb), L.f32c(op), L.f32c(rgb_bg_raw), L.f32c(op_bg), L.f32c(bg_color)
        n = rgb.shape[0]
        dev = rgb.device
        rgb_bg, rgb_full = torch.empty_like(rgb), torch.empty_like(rgb)
        valid = torch.empty(3, n, 1, device=dev, dtype=torch.bool)
        _run("ia_ray_mix_fwd", L.ptr(rgb), L.ptr(op), L.ptr(rgb_bg_raw), L.ptr(op_bg), L.ptr(bg_color), n, L.ptr(rgb_bg), L.ptr(rgb_full),
             valid.data_ptr(), valid.data_ptr() + n, valid.data_ptr() + 2 * n, L.stream())
        ctx.save_for_backward(op, op_bg, bg_color, rgb_bg)
        v0, v1, v2 = valid[0], valid[1], valid[2]
        ctx.mark_non_differentiable(v0, v1, v2)
        return rgb_bg, rgb_full, v0, v1, v2

    @staticmethod
    def backward(ctx, g_bg, g_full, _a, _b, _c):
        op, op_bg, bg_color, rgb_bg = ctx.saved_tensors
        need = ctx.needs_input_grad
        n = op.shape[0]
        g_bg = L.f32c(g_bg) if g_bg is not None else None
        g_full = L.f32c(g_full) if g_full is not None else None
        d_rgb = torch.empty(n, 3, device=op.device) if need[0] else None
        d_op = torch.empty_like(op) if need[1] else None
        d_raw = torch.empty(n, 3, device=op.device) if need[2] else None
        d_opb = torch.empty_like(op_bg) if need[3] else None
        _run("ia_ray_mix_bwd", L.ptr(op), L.ptr(op_bg), L.ptr(bg_color), L.ptr(rgb_bg), L.ptr(g_bg), L.ptr(g_full), n, L.ptr(d_rgb),
             L.ptr(d_op), L.ptr(d_raw), L.ptr(d_opb), L.stream())
        return d_rgb, d_op, d_raw, d_opb, None


def ray_mix(comp_rgb, opacity, comp_rgb_bg_raw, opacity_bg, background_color):
    """Foreground / background mix of reference models/neus.py:186, 272-276 in one launch each way.
    comp_rgb [R,3], opacity [R,1], comp_rgb_bg_raw [R,3], opacity_bg [R,1], background_color [3] ->
    (comp_rgb_bg [R,3], comp_rgb_full [R,3], rays_valid [R,1], rays_valid_bg [R,1], rays_valid_full [R,1])."""
    return _RayMixFn.apply(comp_rgb, opacity, comp_rgb_bg_raw, opacity_bg, background_color)


class _PointLossesFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sdf, grad, normal_gt, weights, lam_sdf, lam_normal):
        L.require_cuda(sdf, grad, normal_gt, weights)
        sdf, grad, normal_gt, weights = L.f32c(sdf), L.f32c(grad), L.f32c(normal_gt), L.f32c(weights)
        n = sdf.shape[0]
        ws = torch.empty(int(L.load().ia_neus_losses_workspace_bytes()), device=sdf.device, dtype=torch.uint8)
        out = torch.empty(4, device=sdf.device, dtype=torch.float32)
        _run("ia_point_losses_fwd", L.ptr(sdf), L.ptr(grad), L.ptr(normal_gt), L.ptr(weights), n, C.c_float(lam_sdf), C.c_float(lam_normal),
             L.ptr(ws), L.ptr(out), L.stream())
        ctx.save_for_backward(sdf, grad, normal_gt, out)
        ctx.lams = (lam_sdf, lam_normal)
        terms = out[:2]
        ctx.mark_non_differentiable(terms)
        return out[2], terms

    @staticmethod
    def backward(ctx, dloss, _dterms):
        sdf, grad, normal_gt, out = ctx.saved_tensors
        need = ctx.needs_input_grad
        dloss = L.f32c(dloss).reshape(1)
        d_sdf = torch.empty_like(sdf) if need[0] else None
        d_grad = torch.empty_like(grad) if need[1] else None
        _run("ia_point_losses_bwd", L.ptr(sdf), L.ptr(grad), L.ptr(normal_gt), sdf.shape[0], C.c_float(ctx.lams[0]), C.c_float(ctx.lams[1]),
             L.ptr(out), L.ptr(dloss), L.ptr(d_sdf), L.ptr(d_grad), L.stream())
        return d_sdf, d_grad, None, None, None, None


def point_losses(sdf, grad, normal_gt, weights, lambda_sdf_l1: float, lambda_normal: float):
    """Sparse-point terms of reference systems/neus.py:173-186 and their weighted sum in one launch each way.
    -> (weighted sum, {'sdf_l1', 'normal_cos'}); the terms are values for logging (not differentiable)."""
    total, terms = _PointLossesFn.apply(sdf.reshape(-1), grad.reshape(-1, 3), normal_gt.reshape(-1, 3), weights.reshape(-1),
                                        float(lambda_sdf_l1), float(lambda_normal))
    return total, {"sdf_l1": terms[0], "normal_cos": terms[1]}
