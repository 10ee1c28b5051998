"""GPU parity tests, operator level: every kernel of libia_b200.so (called through the C ABI via
instant_angelo_b200.ops) against the CPU oracle on identical seeded inputs.

Bars (BASELINE.json north_star): occupancy bits, packed_info, ray_indices, t_starts/t_ends: BIT-EXACT;
floating-point outputs and gradients: within 1e-3 relative (fp32-accumulate) -- the fp32 kernels are held to
much tighter tolerances here so that regressions show."""
import math

import numpy as np
import pytest
import torch

from oracle import model_ref as mr
from oracle import nerfacc_ref as nf
from oracle import tcnn_ref as tc
from tests.helpers import assert_close, grad_tol

pytestmark = pytest.mark.gpu

PLS = 1.3195079107728942


def _dev(t):
    return t.cuda()


# ---------------------------------------------------------------------------------------------
# hash grid
# ---------------------------------------------------------------------------------------------
GRID_CFGS = {
    "sparse_2p19": dict(n_levels=16, n_features=2, log2_hashmap_size=19, base_resolution=32, per_level_scale=PLS),
    "dense_2p21": dict(n_levels=16, n_features=2, log2_hashmap_size=21, base_resolution=32, per_level_scale=PLS),   # BASELINE config 3
    "small_mixed": dict(n_levels=8, n_features=2, log2_hashmap_size=12, base_resolution=4, per_level_scale=1.5),
    "tiny_hash": dict(n_levels=5, n_features=2, log2_hashmap_size=8, base_resolution=16, per_level_scale=2.0),
}


def _points(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, generator=g)
    # edge cases: corners of the unit cube, exact cell boundaries, tap-like clusters
    x[0] = 0.0
    x[1] = 1.0
    x[2] = torch.tensor([0.5, 0.25, 0.75])
    x[3] = torch.tensor([1.0, 0.0, 1.0])
    if n > 64:
        x[8:15] = x[7] + torch.tensor([[0, 0, 0], [1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]) * 1e-3
    return x.clamp(0, 1)


@pytest.mark.parametrize("cfg_name", list(GRID_CFGS))
@pytest.mark.parametrize("active", [None, 4])
def test_hashgrid_forward_backward(cuda_lib, cfg_name, active):
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS[cfg_name]
    plan_ref = tc.grid_plan(**cfg)
    plan = ops.make_grid_plan(**cfg)
    assert plan.n_params == plan_ref.n_params
    n = 1000 if cfg_name in ("sparse_2p19", "dense_2p21") else 777
    act = active if active is None else min(active, cfg["n_levels"])
    g = torch.Generator().manual_seed(7)
    x = _points(n, 11).requires_grad_(True)
    table = (torch.randn(plan_ref.n_params, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(n, plan_ref.n_output_dims, generator=g)

    y_ref = tc.hashgrid_forward(x, table, plan_ref, act)
    y_ref.backward(dy)

    xg = x.detach().cuda().requires_grad_(True)
    tg = table.detach().cuda().requires_grad_(True)
    y = ops.hashgrid_encode(xg, tg, plan, act)
    y.backward(dy.cuda())
    torch.cuda.synchronize()

    assert_close(y, y_ref, rtol=1e-5, atol=1e-6, name="enc")
    if act is not None:
        assert torch.count_nonzero(y[:, act * 2:]) == 0, "masked levels must be exact zeros"
    assert_close(tg.grad, table.grad, rtol=1e-4, atol=1e-5, name="dtable")
    # d enc / d x is discontinuous at cell faces; points 0..3 sit exactly on faces -> compare the rest
    rt, at = grad_tol(x.grad[4:], 1e-4)
    assert_close(xg.grad[4:], x.grad[4:], rtol=rt, atol=at, name="dx")


@pytest.mark.parametrize("cfg_name", ["sparse_2p19", "small_mixed"])
@pytest.mark.parametrize("active", [None, 5])
def test_hashgrid_grouped_backward(cuda_lib, cfg_name, active):
    """ia_hashgrid_bwd_grouped (6 finite-difference taps per group, same-cell contributions merged before the scatter)
    against the oracle's autograd, for tap clusters at several scales (same cell at coarse levels, different cells at
    fine ones) and a row count that is not a multiple of the CTA tile."""
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS[cfg_name]
    plan_ref, plan = tc.grid_plan(**cfg), ops.make_grid_plan(**cfg)
    g = torch.Generator().manual_seed(17)
    S = 173
    centre = torch.rand(S, 1, 3, generator=g) * 0.98 + 0.01
    eps = torch.where(torch.arange(S) % 3 == 0, 5e-4, torch.where(torch.arange(S) % 3 == 1, 4e-3, 3e-2))[:, None, None]
    signs = torch.tensor([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, 1.0], [0, 0, -1.0]])
    x = (centre + signs * eps).clamp(0, 1).reshape(-1, 3).contiguous().requires_grad_(True)
    n = x.shape[0]
    table = (torch.randn(plan_ref.n_params, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(n, plan_ref.n_output_dims, generator=g)
    dy[5::7] = 0.0                                            # some rows with exactly-zero gradient
    act = active if active is None else min(active, cfg["n_levels"])
    y_ref = tc.hashgrid_forward(x, table, plan_ref, act)
    y_ref.backward(dy)
    xg, tg = x.detach().cuda().requires_grad_(True), table.detach().cuda().requires_grad_(True)
    y = ops.hashgrid_encode(xg, tg, plan, act, group=6)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    # ia_hashgrid_fwd_grouped (one thread walks the 6 taps; opt-in, see ops._FWD_GROUPED) through the raw ABI
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    y = torch.empty(n, plan_ref.n_output_dims, device="cuda")
    L.check(cuda_lib.ia_hashgrid_fwd_grouped(xg.data_ptr(), n, tg.data_ptr(), C.byref(plan), plan.n_levels if act is None else act, 6,
                                             y.data_ptr(), L.stream()))
    torch.cuda.synchronize()
    assert_close(y, y_ref.detach(), rtol=1e-5, atol=1e-6, name="enc (grouped forward)")
    if act is not None:
        assert torch.count_nonzero(y[:, act * 2:]) == 0, "masked levels must be exact zeros"
    y_plain = ops.hashgrid_encode(x.detach().cuda(), table.detach().cuda(), plan, act)
    assert_close(y, y_plain, rtol=1e-5, atol=1e-6, name="enc (grouped vs plain forward)")
    assert_close(tg.grad, table.grad, rtol=1e-4, atol=1e-5, name="dtable (grouped)")
    rt, at = grad_tol(x.grad, 1e-4)
    assert_close(xg.grad, x.grad, rtol=rt, atol=at, name="dx (grouped)")
    # table-only variant (taps that do not need d/dx)
    tg2 = table.detach().cuda().requires_grad_(True)
    ops.hashgrid_encode(x.detach().cuda(), tg2, plan, act, group=6).backward(dy.cuda())
    assert_close(tg2.grad, table.grad, rtol=1e-4, atol=1e-5, name="dtable (grouped, table only)")


@pytest.mark.parametrize("cfg_name", ["sparse_2p19", "small_mixed"])
@pytest.mark.parametrize("group", [1, 6])
def test_hashgrid_fp16_shadow_table(cuda_lib, cfg_name, group):
    """`table_precision: fp16`: the kernels that read an fp16 shadow (ia_hashgrid_fwd_h / _bwd_h / _jvp_h) compute, bit for bit,
    what the fp32 kernels compute on the table rounded to fp16 (the shadow only changes what a gather returns), the table
    gradient lands on the fp32 table, and against the CPU oracle on the ROUNDED table the usual tolerances hold -- the
    deviation from the un-rounded oracle is the rounding of the entries (2^-11 relative), which is why the mode is opt-in."""
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS[cfg_name]
    plan_ref, plan = tc.grid_plan(**cfg), ops.make_grid_plan(**cfg)
    g = torch.Generator().manual_seed(23)
    S = 211
    centre = torch.rand(S, 1, 3, generator=g) * 0.98 + 0.01
    signs = torch.tensor([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, 1.0], [0, 0, -1.0]])
    x = (centre + signs * 2e-3).clamp(0, 1).reshape(-1, 3).contiguous()
    n = x.shape[0]
    table = torch.randn(plan_ref.n_params, generator=g) * 0.1
    dy = torch.randn(n, plan_ref.n_output_dims, generator=g)
    v = torch.randn(n, 3, generator=g)
    act = min(11, cfg["n_levels"])

    tg = table.cuda().requires_grad_(True)
    th = ops.table_to_half(tg)
    assert th.dtype == torch.float16 and torch.equal(th, table.cuda().half())             # round to nearest even, as torch
    rounded = th.float().requires_grad_(True)                                              # fp32 kernels on the rounded table
    xg, xr = x.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    y_h = ops.hashgrid_encode(xg, tg, plan, act, group, table_h=th)
    y_r = ops.hashgrid_encode(xr, rounded, plan, act, group)
    assert torch.equal(y_h, y_r)
    y_h.backward(dy.cuda())
    y_r.backward(dy.cuda())
    assert torch.equal(xg.grad, xr.grad)
    rt, at = grad_tol(rounded.grad, 1e-6)               # same kernel, same addends: only the order of the atomic adds differs
    assert_close(tg.grad, rounded.grad, rtol=rt, atol=at, name="dtable (atomic order only)")
    # second-order pair of the analytic normal
    tg2, rounded2 = table.cuda().requires_grad_(True), th.float().requires_grad_(True)
    dyg, dyr = dy.cuda().requires_grad_(True), dy.cuda().requires_grad_(True)
    dx_h = ops.hashgrid_input_grad(x.cuda(), tg2, dyg, plan, act, th)
    dx_r = ops.hashgrid_input_grad(x.cuda(), rounded2, dyr, plan, act)
    assert torch.equal(dx_h, dx_r)
    (dx_h * v.cuda()).sum().backward()
    (dx_r * v.cuda()).sum().backward()
    assert torch.equal(dyg.grad, dyr.grad)
    rt, at = grad_tol(rounded2.grad, 1e-6)
    assert_close(tg2.grad, rounded2.grad, rtol=rt, atol=at, name="dtable 2nd order (atomic order only)")
    # oracle on the rounded table
    xo, to = x.clone().requires_grad_(True), th.float().cpu().requires_grad_(True)
    y_ref = tc.hashgrid_forward(xo, to, plan_ref, act)
    y_ref.backward(dy)
    assert_close(y_h, y_ref, rtol=1e-5, atol=1e-6, name="enc vs oracle(rounded table)")
    assert_close(tg.grad, to.grad, rtol=1e-4, atol=1e-5, name="dtable vs oracle")
    # ... and the distance to the un-rounded oracle is the rounding of the entries
    y_full = tc.hashgrid_forward(x, table, plan_ref, act)
    assert float((y_h.cpu() - y_full).abs().max()) <= 2.0 ** -11 * float(table.abs().max()) * 1.01
    # argument checks
    with pytest.raises(ValueError):
        ops.hashgrid_encode(xg, tg, plan, act, group, table_h=th[:-2])
    with pytest.raises(ValueError):
        ops.hashgrid_encode(xg, tg, plan, act, group, table_h=th.float())


def test_encoding_fp16_shadow_follows_the_parameters(cuda_lib):
    """Encoding(table_precision='fp16'): the shadow is re-derived after update_step() (the hook that follows the fused
    optimizer's raw-pointer writes) and after in-place torch writes; state-dict keys are those of the fp32 module."""
    from instant_angelo_b200 import ops
    from instant_angelo_b200.network_utils import get_encoding
    cfg = {"otype": "ProgressiveBandHashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 12, "base_resolution": 8,
           "per_level_scale": 1.5, "start_level": 4, "start_step": 0, "update_steps": 1, "include_xyz": True}
    enc_h = get_encoding(3, {**cfg, "table_precision": "fp16"}).cuda()
    enc_f = get_encoding(3, cfg).cuda()
    assert list(enc_h.state_dict().keys()) == list(enc_f.state_dict().keys())
    grid = enc_h.encoding.encoding
    with torch.no_grad():
        grid.params.mul_(1e3)
    enc_h.update_step(0, 10)
    enc_f.update_step(0, 10)
    x = torch.rand(500, 3, generator=torch.Generator().manual_seed(1)).cuda()
    enc_f.load_state_dict({k: v.half().float() for k, v in enc_h.state_dict().items()})
    assert torch.equal(enc_h(x), enc_f(x))
    # fused optimizer step (raw-pointer writes, invisible to torch's version counter): ops.param_epoch() moves
    ops.adamw_step(grid.params.data, torch.ones_like(grid.params), torch.zeros_like(grid.params), torch.zeros_like(grid.params),
                   0.05, 0.9, 0.99, 1e-15, 0.0, 1)
    fresh = enc_h(x)                                # no update_step in between: the optimizer step itself marked the copy stale
    enc_f.load_state_dict({k: v.half().float() for k, v in enc_h.state_dict().items()})
    assert torch.equal(fresh, enc_f(x))
    # a raw write that bypasses ops.adamw_step (the bare ABI call) is invisible to both counters: picked up at the next update_step
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    m, v, g1 = torch.zeros_like(grid.params), torch.zeros_like(grid.params), torch.ones_like(grid.params)
    L.check(L.load().ia_adamw_step(L.ptr(grid.params), L.ptr(g1), L.ptr(m), L.ptr(v), grid.params.numel(), C.c_float(0.05), C.c_float(0.9),
                                   C.c_float(0.99), C.c_float(1e-15), C.c_float(0.0), 1, C.c_float(1.0), L.stream()))
    stale = enc_h(x)
    enc_h.update_step(0, 11)
    fresh = enc_h(x)
    enc_f.load_state_dict({k: v_.half().float() for k, v_ in enc_h.state_dict().items()})
    assert torch.equal(fresh, enc_f(x)) and not torch.equal(stale, fresh)
    # in-place torch write: seen through the version counter without update_step
    with torch.no_grad():
        grid.params.add_(0.01)
    enc_f.load_state_dict({k: v.half().float() for k, v in enc_h.state_dict().items()})
    assert torch.equal(enc_h(x), enc_f(x))


@pytest.mark.parametrize("cfg_name", ["sparse_2p19", "small_mixed"])
def test_hashgrid_points_outside_unit_cube(cuda_lib, cfg_name):
    """tcnn's index is total: for x outside [0,1] the cell coordinates are negative / beyond the grid, the uint32 stride
    sum wraps and `% size` (dense levels) / the hash mask (hashed levels) keep it inside the level.  The sparse-point
    losses hand un-clamped COLMAP points to VolumeSDF (reference systems/neus.py:178-186 -> models/geometry.py:200-206),
    so the kernels must neither read nor scatter outside the table: forward, table/input gradients, grouped backward and
    the second-order adjoints against the oracle on x in [-2, 3], with canaries around the table gradient."""
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS[cfg_name]
    plan_ref, plan = tc.grid_plan(**cfg), ops.make_grid_plan(**cfg)
    g = torch.Generator().manual_seed(23)
    n = 6 * 211
    x = (torch.rand(n, 3, generator=g) * 5.0 - 2.0)
    x[::5] = torch.rand(x[::5].shape, generator=g)            # a share of in-range points
    x[1] = torch.tensor([-2.0, 3.0, -0.5])
    x[2] = torch.tensor([1.0 + 1e-6, -1e-6, 0.37])
    x = x.contiguous().requires_grad_(True)
    table = (torch.randn(plan_ref.n_params, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(n, plan_ref.n_output_dims, generator=g)
    y_ref = tc.hashgrid_forward(x, table, plan_ref)
    y_ref.backward(dy)
    for group in (1, 6):
        xg = x.detach().cuda().requires_grad_(True)
        pad = 4096
        guard = torch.zeros(plan.n_params + 2 * pad, device="cuda")
        guard[pad:pad + plan.n_params] = table.detach().cuda()
        tg = guard[pad:pad + plan.n_params].requires_grad_(True)
        y = ops.hashgrid_encode(xg, tg, plan, None, group=group)
        y.backward(dy.cuda())
        torch.cuda.synchronize()
        assert_close(y, y_ref, rtol=1e-5, atol=1e-6, name=f"enc outside cube (group {group})")
        assert_close(tg.grad, table.grad, rtol=1e-4, atol=1e-5, name=f"dtable outside cube (group {group})")
        rt, at = grad_tol(x.grad, 1e-4)
        assert_close(xg.grad, x.grad, rtol=rt, atol=at, name=f"dx outside cube (group {group})")
    # raw ABI with canaries around dtable: nothing may be scattered outside [0, n_params)
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    pad = 1 << 16
    buf = torch.zeros(plan.n_params + 2 * pad, device="cuda")
    dt = buf[pad:pad + plan.n_params]
    xc, tcu, dyc = x.detach().cuda(), table.detach().cuda(), dy.cuda()
    v = torch.randn(n, 3, generator=g).cuda()
    s = L.stream()
    yg = torch.empty(n, plan_ref.n_output_dims, device="cuda")
    L.check(cuda_lib.ia_hashgrid_fwd_grouped(xc.data_ptr(), n, tcu.data_ptr(), C.byref(plan), plan.n_levels, 6, yg.data_ptr(), s))
    assert_close(yg, y_ref.detach(), rtol=1e-5, atol=1e-6, name="enc outside cube (grouped forward kernel)")
    L.check(cuda_lib.ia_hashgrid_bwd(xc.data_ptr(), n, tcu.data_ptr(), dyc.data_ptr(), C.byref(plan), plan.n_levels, dt.data_ptr(), None, s))
    L.check(cuda_lib.ia_hashgrid_bwd_grouped(xc.data_ptr(), n, tcu.data_ptr(), dyc.data_ptr(), C.byref(plan), plan.n_levels, 6, dt.data_ptr(), None, s))
    L.check(cuda_lib.ia_hashgrid_bwd_input_bwd_table(xc.data_ptr(), n, v.data_ptr(), dyc.data_ptr(), C.byref(plan), plan.n_levels, dt.data_ptr(), s))
    torch.cuda.synchronize()
    assert float(buf[:pad].abs().max()) == 0.0 and float(buf[pad + plan.n_params:].abs().max()) == 0.0, "scatter outside the table"
    # second-order adjoints (grad_type analytic) on the same points
    xr = x.detach().clone()
    tr = table.detach().clone().requires_grad_(True)
    dyr = dy.clone().requires_grad_(True)
    xr.requires_grad_(True)
    (dx_ref,) = torch.autograd.grad(tc.hashgrid_forward(xr, tr, plan_ref), xr, dyr, create_graph=True)
    (dx_ref * v.cpu()).sum().backward()
    tg2 = table.detach().cuda().requires_grad_(True)
    dyg = dy.cuda().requires_grad_(True)
    dxg = ops.hashgrid_input_grad(xc, tg2, dyg, plan)
    (dxg * v).sum().backward()
    torch.cuda.synchronize()
    rt, at = grad_tol(dx_ref, 1e-4)
    assert_close(dxg, dx_ref, rtol=rt, atol=at, name="J^T dy outside cube")
    rt, at = grad_tol(dyr.grad, 1e-4)
    assert_close(dyg.grad, dyr.grad, rtol=rt, atol=at, name="jvp outside cube")
    rt, at = grad_tol(tr.grad, 1e-4)
    assert_close(tg2.grad, tr.grad, rtol=rt, atol=at, name="second-order dtable outside cube")


def test_hashgrid_scatters_into_the_gradient_arena(cuda_lib):
    """Parameters re-homed by dp.ParamArena receive their table gradient by atomic adds straight into the arena slice
    (ops._grad_sink): same values as autograd's zero-filled temporary + accumulate, added on top of what is already there,
    for the plain, the grouped and the second-order backward; create_graph falls back to the autograd path."""
    from instant_angelo_b200 import ops
    from instant_angelo_b200.dp import ParamArena
    cfg = GRID_CFGS["small_mixed"]
    plan = ops.make_grid_plan(**cfg)
    g = torch.Generator().manual_seed(41)
    n = 6 * 100
    x = torch.rand(n, 3, generator=g).cuda()
    t0 = (torch.randn(plan.n_params, generator=g) * 0.1).cuda()
    dy = torch.randn(n, 16, generator=g).cuda()
    dy2 = torch.randn(n, 16, generator=g).cuda()
    v = torch.randn(n, 3, generator=g).cuda()

    def run(table):
        ops.hashgrid_encode(x, table, plan).backward(dy)
        ops.hashgrid_encode(x, table, plan, 5, group=6).backward(dy2)
        (ops.hashgrid_input_grad(x, table, dy, plan) * v).sum().backward()

    plain = torch.nn.Parameter(t0.clone())
    run(plain)
    homed = torch.nn.Parameter(t0.clone())
    other = torch.nn.Parameter(torch.ones(7, device="cuda"))
    arena = ParamArena([other, homed])
    assert homed.grad.data_ptr() == arena.grad[arena.offsets[1]:].data_ptr()
    arena.grad.fill_(0.25)                      # pre-existing gradient: the kernels must ADD
    grad_view = homed.grad
    run(homed)
    torch.cuda.synchronize()
    assert homed.grad is grad_view and homed.grad.data_ptr() == arena.grad[arena.offsets[1]:].data_ptr(), "autograd replaced .grad"
    rt, at = grad_tol(plain.grad, 1e-5)          # same kernels, different atomic order
    assert_close(homed.grad - 0.25, plain.grad, rtol=rt, atol=at, name="arena-resident table gradient")
    assert float((arena.grad[:7] - 0.25).abs().max()) == 0.0
    # create_graph: autograd needs the value -> no in-place accumulation
    arena.zero_grad()
    y = ops.hashgrid_input_grad(x, homed, dy, plan)
    (gt,) = torch.autograd.grad((y * v).sum(), homed, create_graph=True)
    assert gt is not None and float(arena.grad.abs().max()) == 0.0
    assert float(gt.abs().max()) > 0.0


def test_l2_persist_window_grant_and_reset(cuda_lib):
    """ia_l2_persist: the window is clipped to the device limits, results do not change, bytes == 0 resets, bad args fail."""
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS["small_mixed"]
    plan = ops.make_grid_plan(**cfg)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(1000, 3, generator=g).cuda()
    table = (torch.randn(plan.n_params, generator=g) * 0.1).cuda()
    before = ops.hashgrid_encode(x, table, plan, plan.n_levels).clone()
    info = ops.l2_persist(table, 1.0)
    assert info["l2_bytes"] > 0 and 0 < info["window_bytes"] <= table.numel() * 4
    assert 0 < info["set_aside_bytes"] <= info["window_bytes"]
    assert torch.equal(ops.hashgrid_encode(x, table, plan, plan.n_levels), before)
    info = ops.l2_persist(None)
    assert info["window_bytes"] == 0 and info["set_aside_bytes"] == 0
    assert torch.equal(ops.hashgrid_encode(x, table, plan, plan.n_levels), before)
    lib = L.load()
    assert lib.ia_l2_persist(L.ptr(table), 1024, C.c_float(1.5), None, L.stream()) == -1      # IA_ERR_INVALID_ARG
    assert lib.ia_l2_persist(None, 1024, C.c_float(1.0), None, L.stream()) == -1
    assert b"l2_persist" in lib.ia_last_error_string()


def test_hashgrid_abi_entry_points_and_errors(cuda_lib):
    """ia_hashgrid_bwd_table / ia_hashgrid_bwd_input agree with the fused ia_hashgrid_bwd; bad args fail loudly."""
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS["small_mixed"]
    plan = ops.make_grid_plan(**cfg)
    g = torch.Generator().manual_seed(3)
    n = 300
    x = torch.rand(n, 3, generator=g).cuda()
    table = (torch.randn(plan.n_params, generator=g) * 0.1).cuda()
    dy = torch.randn(n, 16, generator=g).cuda()
    s = L.stream()
    dt_a, dx_a = torch.zeros_like(table), torch.empty_like(x)
    dt_b, dx_b = torch.zeros_like(table), torch.empty_like(x)
    L.check(cuda_lib.ia_hashgrid_bwd(x.data_ptr(), n, table.data_ptr(), dy.data_ptr(), C.byref(plan), 8, dt_a.data_ptr(), dx_a.data_ptr(), s))
    L.check(cuda_lib.ia_hashgrid_bwd_table(x.data_ptr(), n, dy.data_ptr(), C.byref(plan), 8, dt_b.data_ptr(), s))
    L.check(cuda_lib.ia_hashgrid_bwd_input(x.data_ptr(), n, table.data_ptr(), dy.data_ptr(), C.byref(plan), 8, dx_b.data_ptr(), s))
    torch.cuda.synchronize()
    assert_close(dt_b, dt_a, rtol=1e-5, atol=1e-6, name="bwd_table")
    assert_close(dx_b, dx_a, rtol=1e-6, atol=1e-7, name="bwd_input")
    # errors: status code + message, no exception across the ABI
    rc = cuda_lib.ia_hashgrid_fwd(x.data_ptr(), n, table.data_ptr(), C.byref(plan), 99, dx_a.data_ptr(), s)
    assert rc == -1 and b"active_levels" in cuda_lib.ia_last_error_string()
    with pytest.raises(NotImplementedError):
        ops.hashgrid_encode(x.cpu(), table.cpu(), plan)
    # empty input
    out = ops.hashgrid_encode(x[:0], table, plan)
    assert out.shape == (0, 16)


def test_hashgrid_full_size_properties(cuda_lib):
    """BASELINE size (2^22 points, 2^19-entry tables): size-independent properties.
    (1) partition of unity: an all-ones table encodes to all ones; (2) linearity in the table;
    (3) checksum: per level/feature, sum(dtable) == sum(dy) because the 8 weights sum to 1."""
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS["sparse_2p19"]
    plan = ops.make_grid_plan(**cfg)
    n = 1 << 22
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(n, 3, device="cuda", generator=g)
    ones = torch.ones(plan.n_params, device="cuda")
    y1 = ops.hashgrid_encode(x, ones, plan)
    assert float((y1 - 1).abs().max()) < 2e-6
    t1 = torch.randn(plan.n_params, device="cuda", generator=g)
    t2 = torch.randn(plan.n_params, device="cuda", generator=g)
    ya, yb = ops.hashgrid_encode(x, t1, plan), ops.hashgrid_encode(x, t2, plan)
    yc = ops.hashgrid_encode(x, 2.0 * t1 - 0.5 * t2, plan)
    assert float((yc - (2.0 * ya - 0.5 * yb)).abs().max()) < 2e-5
    n2 = 1 << 20
    tg = t1.clone().requires_grad_(True)
    dy = torch.rand(n2, 32, device="cuda", generator=g)
    ops.hashgrid_encode(x[:n2], tg, plan).backward(dy)
    got = []
    for l in range(16):
        lo, hi = plan.offset[l] * 2, plan.offset[l + 1] * 2
        got.append(tg.grad[lo:hi].view(-1, 2).double().sum(0))
    got = torch.stack(got).reshape(-1)
    want = dy.double().sum(0)
    assert_close(got, want, rtol=1e-4, atol=1e-2, name="dtable checksum")


# ---------------------------------------------------------------------------------------------
# spherical harmonics
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("degree", [1, 2, 3, 4])
def test_sh(cuda_lib, degree):
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(degree)
    d = torch.nn.functional.normalize(torch.randn(513, 3, generator=g), dim=-1)
    d01 = ((d + 1) / 2).requires_grad_(True)
    y_ref = tc.sh_forward(d01, degree)
    go = torch.randn_like(y_ref)
    dg = d01.detach().cuda().requires_grad_(True)
    y = ops.sh_encode(dg, degree)
    y.backward(go.cuda())
    assert_close(y, y_ref, rtol=1e-5, atol=1e-6, name="sh")
    if degree == 1:      # constant basis function: zero input gradient
        assert float(dg.grad.abs().max()) == 0.0
    else:
        y_ref.backward(go)
        assert_close(dg.grad, d01.grad, rtol=1e-4, atol=1e-5, name="dsh")


# ---------------------------------------------------------------------------------------------
# fused MLP (fp32)
# ---------------------------------------------------------------------------------------------
MLP_CASES = {
    #                 n_in0 n_in1 hidden out  softplus weight_norm n_out_used
    "geometry_full": (3, 32, 2, 65, True, True, 65),
    "geometry_sdf": (3, 32, 2, 65, True, True, 1),
    "texture": (0, 87, 2, 3, False, False, 3),
    "texture_fold": (0, 88, 2, 3, False, False, 3),      # colour head after the output-layer fold: mlp_tc_bwd_duo96_kernel
    "v3_weight": (0, 71, 2, 1, False, False, 1),
    "bg_geometry": (3, 32, 1, 8, False, False, 8),
    "bg_texture": (0, 24, 2, 3, False, False, 3),
    "neus_colmap_geo": (3, 32, 1, 13, True, True, 13),
}


@pytest.mark.parametrize("case", list(MLP_CASES))
@pytest.mark.parametrize("n", [1, 1000])
@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_mlp(cuda_lib, case, n, precision):
    """fp32: FFMA kernel; tc: tcgen05 3xF16-split kernel (wide output layers route to the fp32 kernel inside)."""
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    n0, n1, nh, nout, softplus, wn, nou = MLP_CASES[case]
    prec = L.IA_MLP_FP32 if precision == "fp32" else L.IA_MLP_TC_F16
    torch.manual_seed(sum(map(ord, case)))
    cfg = {"n_neurons": 64, "n_hidden_layers": nh, "sphere_init": softplus, "weight_norm": wn, "output_activation": "none"}
    ref = mr.RefVanillaMLP(n0 + n1, nout, cfg)
    with torch.no_grad():  # make every weight matter (sphere init zeroes the non-xyz columns)
        for p in ref.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    g = torch.Generator().manual_seed(n)
    a = torch.rand(n, n0, generator=g).requires_grad_(True) if n0 else None
    b = (torch.randn(n, n1, generator=g) * 0.3).requires_grad_(True)
    inp = torch.cat([a * 2 - 1, b], dim=1) if n0 else b
    y_ref = ref(inp)[:, :nou]
    go = torch.randn(n, nou, generator=g)
    y_ref.backward(go)

    # product: same effective weights in the ABI's flat layout
    lin = [m for m in ref.layers if isinstance(m, mr.RefWNLinear)]
    flat = torch.cat([t for m in lin for t in (m.effective_weight().reshape(-1), m.bias)]).detach().cuda().requires_grad_(True)
    desc = ops.make_mlp_desc(n0, n1, nh, nout, L.IA_ACT_SOFTPLUS100 if softplus else L.IA_ACT_RELU, 2.0, -1.0, prec)
    assert cuda_lib.ia_mlp_param_count(desc) == flat.numel()
    ag = a.detach().cuda().requires_grad_(True) if n0 else None
    bg = b.detach().cuda().requires_grad_(True)
    y = ops.mlp_apply(ag, bg, flat, desc, nou)
    y.backward(go.cuda())
    torch.cuda.synchronize()
    assert_close(y, y_ref, rtol=1e-4, atol=1e-5, name="mlp out")
    rt, at = grad_tol(b.grad, 2e-4)
    assert_close(bg.grad, b.grad, rtol=rt, atol=at, name="d in1")
    if n0:
        rt, at = grad_tol(a.grad, 2e-4)
        assert_close(ag.grad, a.grad, rtol=rt, atol=at, name="d in0")
    # parameter grads: reference grads w.r.t. effective weights via autograd on a functional copy
    eff = [(m.effective_weight().detach().requires_grad_(True), m.bias.detach().clone().requires_grad_(True)) for m in lin]
    h = inp.detach()
    for i, (w, bias) in enumerate(eff):
        h = torch.nn.functional.linear(h, w, bias)
        if i < len(eff) - 1:
            h = torch.nn.functional.softplus(h, beta=100) if softplus else torch.relu(h)
    h[:, :nou].backward(go)
    want = torch.cat([t.grad.reshape(-1) if t.grad is not None else torch.zeros_like(t).reshape(-1) for pair in eff for t in pair])
    rt, at = grad_tol(want, 2e-4)
    assert_close(flat.grad, want, rtol=rt, atol=at, name="d params")


MLP_CASES_LARGE = dict(MLP_CASES, v3_cam=(0, 80, 2, 3, False, False, 3))


def _mlp_ref64(a, b, Ws, bs, softplus, nou):
    """float64 torch reference of VanillaMLP.forward (reference models/network_utils.py:96-113) on cat[a*2-1, b];
    also returns, per row, the smallest |pre-activation| over the hidden units (distance to the nearest ReLU kink)."""
    h = (torch.cat([a * 2 - 1, b], 1) if a is not None else b).double()
    zmin = torch.full((h.shape[0],), float("inf"), device=h.device, dtype=torch.float64)
    for i, (w, bias) in enumerate(zip(Ws, bs)):
        h = h @ w.double().t() + bias.double()
        if i < len(Ws) - 1:
            zmin = torch.minimum(zmin, h.detach().abs().min(dim=1).values)
            h = torch.nn.functional.softplus(h, beta=100) if softplus else torch.relu(h)
    return h[:, :nou], zmin


def _assert_rel(got, want, name, rtol=1e-3, floor=1e-2, l2=1e-4):
    """north_star tolerance: |got - want| <= rtol * |want| on every entry above floor * max|want|; smaller entries are held to
    rtol * floor * max|want| = 1e-5 of the tensor's scale absolute; and the relative L2 error is below `l2`.
    Why a floor at all: an fp32-accumulated 64..96-term dot product carries an absolute rounding error proportional to the
    size of its TERMS (~5e-7 of the tensor's scale here, for the fp32 FFMA kernel and the tcgen05 kernel alike: measured
    1e-6 and 2e-6 of max over 3e7 entries), so an entry that is small through cancellation cannot be relatively accurate
    in any fp32 implementation, the reference's included."""
    got, want = got.detach().double(), want.detach().double()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    if want.numel() == 0:
        return
    scale = float(want.abs().max())
    if scale == 0.0:
        assert float(got.abs().max()) == 0.0, f"{name}: expected exact zeros"
        return
    err = (got - want).abs()
    tol = rtol * torch.clamp(want.abs(), min=floor * scale)
    bad = err > tol
    if bool(bad.any()):
        i = int(torch.argmax(err - tol))
        raise AssertionError(f"{name}: {int(bad.sum())}/{want.numel()} entries out of tolerance; worst flat index {i}: got "
                             f"{float(got.reshape(-1)[i]):.8g} want {float(want.reshape(-1)[i]):.8g} (max|want| {scale:.3g})")
    rel_l2 = float((got - want).norm() / want.norm().clamp_min(1e-300))
    assert rel_l2 < l2, f"{name}: relative L2 error {rel_l2:.3g} >= {l2}"


@pytest.mark.parametrize("case", list(MLP_CASES_LARGE))
@pytest.mark.parametrize("n", [50_000, 1_000_003])
@pytest.mark.parametrize("dout_scale", [1.0, 1e-4])
@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_mlp_large_n(cuda_lib, case, n, dout_scale, precision):
    """The regime bench.py runs: hundreds of 128-row tiles per persistent CTA (cross-tile dW accumulation, per-tile
    power-of-two gradient rescale, the software pipeline's steady state, a ragged last tile), against a float64 torch
    reference of the same network.  Tolerance 1e-3 relative per entry (entries > 1e-3 of the tensor's max) and 1e-4
    relative L2.

    ReLU networks: the gradient is discontinuous where a hidden pre-activation crosses zero.  With ~1e8 hidden units
    per case a handful lie within rounding distance (1e-6) of the kink, and there fp32 (either kernel) and float64
    legitimately pick different sides -- a whole row of d(input) then differs by O(1).  (This is what the round-1
    probe's `v3_weight n=50000 din1 9.8e-02` line was: one row, on the kink.)  Rows whose float64 pre-activations come
    within 1e-5 of zero get a zero incoming gradient, so they take part in the forward comparison only."""
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    n0, n1, nh, nout, softplus, _wn, nou = MLP_CASES_LARGE[case]
    prec = L.IA_MLP_FP32 if precision == "fp32" else L.IA_MLP_TC_F16
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(sum(map(ord, case)) + n)
    din = n0 + n1
    dims = [din] + [64] * nh + [nout]
    Ws = [(torch.randn(dims[i + 1], dims[i], device=dev, generator=g) * (1.5 / dims[i] ** 0.5)).requires_grad_(True)
          for i in range(len(dims) - 1)]
    bs = [(torch.randn(dims[i + 1], device=dev, generator=g) * 0.1).requires_grad_(True) for i in range(len(dims) - 1)]
    a = torch.rand(n, n0, device=dev, generator=g).requires_grad_(True) if n0 else None
    b = (torch.randn(n, n1, device=dev, generator=g) * 0.3).requires_grad_(True)
    go = torch.randn(n, nou, device=dev, generator=g) * dout_scale
    y64, zmin = _mlp_ref64(a, b, Ws, bs, softplus, nou)
    if not softplus:
        on_kink = zmin < 1e-5
        assert int(on_kink.sum()) < max(8, n // 100), "kink mask should only remove a small share of the rows"
        go = torch.where(on_kink[:, None], torch.zeros_like(go), go)
    y64.backward(go.double())
    want_p = torch.cat([t.grad.reshape(-1) for pair in zip(Ws, bs) for t in pair])

    flat = torch.cat([t.detach().reshape(-1) for pair in zip(Ws, bs) for t in pair]).requires_grad_(True)
    desc = ops.make_mlp_desc(n0, n1, nh, nout, L.IA_ACT_SOFTPLUS100 if softplus else L.IA_ACT_RELU, 2.0, -1.0, prec)
    ag = a.detach().clone().requires_grad_(True) if n0 else None
    bg = b.detach().clone().requires_grad_(True)
    y = ops.mlp_apply(ag, bg, flat, desc, nou)
    y.backward(go)
    torch.cuda.synchronize()
    _assert_rel(y, y64, "mlp out")
    _assert_rel(bg.grad, b.grad, "d in1")
    if n0:
        _assert_rel(ag.grad, a.grad, "d in0")
    _assert_rel(flat.grad, want_p, "d params")


@pytest.mark.parametrize("n", [1, 777, 50_003, 300_007])
@pytest.mark.parametrize("cot_scale", [1.0, 1e-4])
@pytest.mark.parametrize("regime", ["geometric", "soft"])
def test_mlp_fwd_grad_second_order(cuda_lib, n, cot_scale, regime):
    """grad_type 'analytic' on the tensor cores (ia_mlp_fwd_grad / ia_mlp_fwd_grad_bwd): the SDF network's last hidden
    layer and d sdf / d(input) in one kernel, and the adjoint of that pair, against torch float64 autograd with
    create_graph=True -- exactly what the reference does (models/geometry.py:206, :214-218) -- at 1e-3 per entry / 1e-4
    relative L2.  'geometric': weights at the scale of the sphere initialisation (Softplus(100) mostly saturated);
    'soft': small pre-activations, where the beta s (1 - s) second-derivative terms dominate the adjoint.  n up to 2345
    tiles (16 per persistent CTA) with a ragged last tile; cotangents scaled by 1 and 1e-4 (launch-wide power-of-two scale)."""
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(1234 + n)
    din, nout = 35, 65
    wscale = 1.5 if regime == "geometric" else 0.08
    dims = [din, 64, 64, nout]
    Ws = [(torch.randn(dims[i + 1], dims[i], device=dev, generator=g) * (wscale / dims[i] ** 0.5)) for i in range(3)]
    Ws[2] = torch.randn(nout, 64, device=dev, generator=g) * 0.3          # output layer: its own scale
    bs = [(torch.randn(dims[i + 1], device=dev, generator=g) * (0.1 if regime == "geometric" else 0.01)) for i in range(3)]
    a = torch.rand(n, 3, device=dev, generator=g)
    b = torch.randn(n, 32, device=dev, generator=g) * 0.3
    ch = torch.randn(n, 64, device=dev, generator=g) * cot_scale
    cg0 = torch.randn(n, 3, device=dev, generator=g) * cot_scale
    cg1 = torch.randn(n, 32, device=dev, generator=g) * cot_scale

    # float64 reference: autograd through autograd
    W64 = [w.double().requires_grad_(True) for w in Ws]
    b64 = [t.double().requires_grad_(True) for t in bs]
    a64, bb64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
    x = torch.cat([a64 * 2 - 1, bb64], 1)
    h1 = torch.nn.functional.softplus(x @ W64[0].t() + b64[0], beta=100)
    h2 = torch.nn.functional.softplus(h1 @ W64[1].t() + b64[1], beta=100)
    y = h2 @ W64[2][0] + b64[2][0]
    g0_ref, g1_ref = torch.autograd.grad(y.sum(), [a64, bb64], create_graph=True)
    loss = (h2 * ch.double()).sum() + (g0_ref * cg0.double()).sum() + (g1_ref * cg1.double()).sum()
    loss.backward()
    want_p = torch.cat([t.grad.reshape(-1) if t.grad is not None else torch.zeros_like(t).reshape(-1)
                        for pair in zip(W64, b64) for t in pair])

    flat = torch.cat([t.reshape(-1) for pair in zip(Ws, bs) for t in pair]).requires_grad_(True)
    desc = ops.make_mlp_desc(3, 32, 2, nout, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_TC_F16)
    assert ops.mlp_fwd_grad_supported(desc)
    ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    h, g0, g1 = ops.mlp_fwd_grad(ag, bg, flat, desc)
    ((h * ch).sum() + (g0 * cg0).sum() + (g1 * cg1).sum()).backward()
    torch.cuda.synchronize()
    _assert_rel(h, h2, "last hidden layer")
    _assert_rel(g0, g0_ref, "d sdf / d in0")
    _assert_rel(g1, g1_ref, "d sdf / d in1")
    # adjoints: Softplus(beta = 100) multiplies the rounding error of a pre-activation (2^-23 relative on z ~ 1) by beta inside
    # sigmoid(beta z) and beta s (1 - s); torch's own fp32 double backward is 9e-6 of the tensor's scale away from float64
    # on these cases (tools/probe_mlp_grad2.py, profiles/r02_mlp_grad2_probe.txt), so entries below 5 % of the maximum are
    # held to 5e-5 of the scale absolute instead of helpers' 1e-5; relative L2 stays at 1e-4
    _assert_rel(ag.grad, a64.grad, "adjoint d in0", floor=5e-2)
    _assert_rel(bg.grad, bb64.grad, "adjoint d in1", floor=5e-2)
    _assert_rel(flat.grad, want_p, "adjoint d params", floor=5e-2)


def test_mlp_fwd_grad_rejects_other_shapes(cuda_lib):
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    desc = ops.make_mlp_desc(3, 32, 1, 13, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_TC_F16)      # one hidden layer
    assert not ops.mlp_fwd_grad_supported(desc)
    x = torch.zeros(4, 3, device="cuda")
    e = torch.zeros(4, 32, device="cuda")
    p = torch.zeros(int(cuda_lib.ia_mlp_param_count(desc)), device="cuda")
    h = torch.zeros(4, 64, device="cuda")
    import ctypes as C
    rc = cuda_lib.ia_mlp_fwd_grad(C.byref(desc), x.data_ptr(), e.data_ptr(), 4, p.data_ptr(), h.data_ptr(), None, None, None)
    assert rc == -2 and "SDF network shape" in L.last_error()          # IA_ERR_UNSUPPORTED


@pytest.mark.parametrize("cfg_name,active", [("sparse_2p19", None), ("sparse_2p19", 6), ("small_mixed", None), ("small_mixed", 4)])
@pytest.mark.parametrize("nou,group,n", [(1, 6, 6 * 20011), (1, 1, 1000), (0, 1, 50_003), (65, 1, 3000), (1, 6, 6)])
def test_sdf_taps_fused_matches_unfused_and_oracle(cuda_lib, cfg_name, active, nou, group, n):
    """ia_sdf_taps_fused_fwd / _bwd (hash-grid gather inside the tcgen05 MLP kernel, no [N, L*F] tensor) against
    (a) the unfused product path hashgrid_encode -> mlp_apply on the same tensor-core arithmetic and (b) the CPU oracle
    (tcnn_ref hash grid + float64 MLP): outputs, d table, d params, d x; tap-like clusters, points outside the unit cube,
    masked levels, ragged last tile, multi-tile persistent CTAs."""
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS[cfg_name]
    plan_ref, plan = tc.grid_plan(**cfg), ops.make_grid_plan(**cfg)
    nl = cfg["n_levels"]
    act = nl if active is None else active
    g = torch.Generator().manual_seed(1000 * nou + n % 997)
    if group == 6:
        centre = torch.rand(n // 6, 1, 3, generator=g)
        signs = torch.tensor([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, 1.0], [0, 0, -1.0]])
        x = (centre + signs * 1.5e-3).clamp(0, 1).reshape(-1, 3)
    else:
        x = torch.rand(n, 3, generator=g)
        x[::17] = x[::17] * 3.0 - 1.0                       # some points outside the unit cube
    x = x.contiguous()
    table = torch.randn(plan_ref.n_params, generator=g) * 0.1
    n_out = 65
    dims = [3 + 2 * nl, 64, 64, n_out]
    Ws = [torch.randn(dims[i + 1], dims[i], generator=g) * (1.0 / dims[i] ** 0.5) for i in range(3)]
    bs = [torch.randn(dims[i + 1], generator=g) * 0.1 for i in range(3)]
    flat = torch.cat([t.reshape(-1) for pair in zip(Ws, bs) for t in pair])
    width_out = 64 if nou == 0 else nou
    go = torch.randn(n, width_out, generator=g)
    desc = ops.make_mlp_desc(3, 2 * nl, 2, n_out, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_TC_F16)
    assert ops.sdf_fused_supported(desc, plan, True)

    def run(fused):
        xg = x.cuda().requires_grad_(True)
        tg = table.cuda().requires_grad_(True)
        fg = flat.cuda().requires_grad_(True)
        if fused:
            y = ops.sdf_fused(xg, tg, fg, desc, plan, act, nou, group)
        else:
            y = ops.mlp_apply(xg, ops.hashgrid_encode(xg, tg, plan, act, group), fg, desc, nou)
        y.backward(go.cuda())
        torch.cuda.synchronize()
        return y.detach(), xg.grad, tg.grad, fg.grad

    y_f, dx_f, dt_f, dp_f = run(True)
    y_u, dx_u, dt_u, dp_u = run(False)
    assert y_f.shape == (n, width_out)
    _assert_rel(y_f, y_u, "out (fused vs unfused)", rtol=1e-4, floor=1e-2, l2=1e-5)
    _assert_rel(dt_f, dt_u, "d table (fused vs unfused)", rtol=1e-3, floor=1e-2, l2=1e-5)
    _assert_rel(dp_f, dp_u, "d params (fused vs unfused)", rtol=1e-3, floor=1e-2, l2=1e-5)
    _assert_rel(dx_f, dx_u, "d x (fused vs unfused)", rtol=1e-3, floor=1e-2, l2=1e-5)
    if act < nl:
        lo = plan.offset[act] * 2
        assert float(dt_f[lo:].abs().max()) == 0.0, "masked levels must receive exactly zero gradient"
    if n <= 3000:        # oracle: tcnn_ref encode + float64 network
        xr = x.clone().requires_grad_(True)
        tr = table.clone().requires_grad_(True)
        Wr = [w.clone().double().requires_grad_(True) for w in Ws]
        br = [b.clone().double().requires_grad_(True) for b in bs]
        h = torch.cat([xr * 2 - 1, tc.hashgrid_forward(xr, tr, plan_ref, act)], dim=1).double()
        h = torch.nn.functional.softplus(h @ Wr[0].t() + br[0], beta=100)
        h = torch.nn.functional.softplus(h @ Wr[1].t() + br[1], beta=100)
        y_r = h if nou == 0 else (h @ Wr[2].t() + br[2])[:, :nou]
        y_r.backward(go.double())
        _assert_rel(y_f.cpu(), y_r, "out (fused vs oracle)")
        _assert_rel(dt_f.cpu(), tr.grad, "d table (fused vs oracle)")
        _assert_rel(dx_f.cpu(), xr.grad, "d x (fused vs oracle)")
        want_p = torch.cat([(t.grad if t.grad is not None else torch.zeros_like(t)).reshape(-1) for pair in zip(Wr, br) for t in pair])
        _assert_rel(dp_f.cpu(), want_p, "d params (fused vs oracle)")


def test_sdf_taps_fused_abi_errors(cuda_lib):
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    plan = ops.make_grid_plan(**GRID_CFGS["small_mixed"])
    x = torch.rand(10, 3, device="cuda")
    table = torch.zeros(plan.n_params, device="cuda")
    params = torch.zeros(20000, device="cuda")
    out = torch.empty(10, 1, device="cuda")
    s = L.stream()
    bad = ops.make_mlp_desc(3, 16, 2, 65, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_FP32)
    rc = cuda_lib.ia_sdf_taps_fused_fwd(C.byref(bad), C.byref(plan), 8, x.data_ptr(), 10, table.data_ptr(), params.data_ptr(), 1, out.data_ptr(), 1, s)
    assert rc != 0 and b"tensor-core" in cuda_lib.ia_last_error_string()
    bad = ops.make_mlp_desc(3, 14, 2, 65, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_TC_F16)
    rc = cuda_lib.ia_sdf_taps_fused_fwd(C.byref(bad), C.byref(plan), 8, x.data_ptr(), 10, table.data_ptr(), params.data_ptr(), 1, out.data_ptr(), 1, s)
    assert rc != 0 and b"n_in1" in cuda_lib.ia_last_error_string()
    ok = ops.make_mlp_desc(3, 16, 2, 65, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_TC_F16)
    rc = cuda_lib.ia_sdf_taps_fused_bwd(C.byref(ok), C.byref(plan), 8, x.data_ptr(), 10, table.data_ptr(), params.data_ptr(), out.data_ptr(), 1, 1, 6,
                                        table.data_ptr(), None, None, None, None, s)
    assert rc != 0 and b"workspace" in cuda_lib.ia_last_error_string()
    assert cuda_lib.ia_sdf_taps_fused_fwd(C.byref(ok), C.byref(plan), 8, x.data_ptr(), 0, table.data_ptr(), params.data_ptr(), 1, out.data_ptr(), 1, s) == 0


def test_mlp_rejects_unsupported(cuda_lib):
    from instant_angelo_b200 import _lib as L
    from instant_angelo_b200 import ops
    desc = ops.make_mlp_desc(0, 16, 2, 3, L.IA_ACT_RELU)
    desc.width = 32
    x = torch.zeros(4, 16, device="cuda")
    p = torch.zeros(10000, device="cuda")
    with pytest.raises(RuntimeError, match="width"):
        ops.mlp_apply(None, x, p, desc)


# ---------------------------------------------------------------------------------------------
# AABB + marching: BIT-EXACT
# ---------------------------------------------------------------------------------------------
def _rays(n, seed, outside_frac=0.3):
    g = torch.Generator().manual_seed(seed)
    o = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * (0.6 + 0.6 * torch.rand(n, 1, generator=g))
    k = int(n * outside_frac)
    o[:k] *= 4.0
    tgt = (torch.rand(n, 3, generator=g) - 0.5) * 1.6
    tgt[-n // 16:] += 5.0
    tgt[: max(4, k // 8)] += 8.0          # origins outside the box looking away: misses (t = 1e10)
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    # axis-aligned rays (zero direction components exercise the inf/NaN paths of the DDA)
    d[k] = torch.tensor([1.0, 0.0, 0.0])
    d[k + 1] = torch.tensor([0.0, -1.0, 0.0])
    d[k + 2] = torch.tensor([0.0, 0.0, 1.0])
    return o.contiguous(), d.contiguous(), g


def test_aabb_bit_exact(cuda_lib):
    from instant_angelo_b200 import nerfacc_api as na
    o, d, _ = _rays(4096, 1)
    aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    for clamp in (True, False):
        tmin_ref, tmax_ref = nf.ray_aabb_intersect(o, d, aabb, clamp)
        tmin, tmax = na.ray_aabb_intersect(o.cuda(), d.cuda(), aabb, clamp)
        assert np.array_equal(tmin.cpu().numpy().view(np.uint32), tmin_ref.numpy().view(np.uint32))
        assert np.array_equal(tmax.cpu().numpy().view(np.uint32), tmax_ref.numpy().view(np.uint32))
    tmin_c, _ = nf.ray_aabb_intersect(o, d, aabb, True)
    assert (tmin_c == 1e10).any() and (tmin_c == 0).any() and ((tmin_c > 0) & (tmin_c < 1e9)).any()
    assert (tmin_ref < 0).any(), "without the clamp a camera inside the box has a negative t_min"


def _bit_equal(a, b, name):
    a = a.cpu().numpy()
    b = b.cpu().numpy() if isinstance(b, torch.Tensor) else b
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    if a.dtype == np.float32:
        a, b = a.view(np.uint32), b.view(np.uint32)
    assert np.array_equal(a, b), f"{name}: {np.count_nonzero(a != b)} of {a.size} differ"


@pytest.mark.parametrize("grid_kind", ["shell", "random", "none", "empty"])
def test_march_foreground_bit_exact(cuda_lib, grid_kind):
    """nerfacc.ray_marching as called at models/neus.py:209-220 (AABB grid, fixed step, stratified)."""
    from instant_angelo_b200 import nerfacc_api as na
    from tests.golden.scenes import sphere_shell_binary
    n = 2048
    o, d, g = _rays(n, 2)
    aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    step = 1.732 * 2 * 1.5 / 512
    u = torch.rand(n, generator=g)
    grid_ref = grid_gpu = None
    if grid_kind != "none":
        grid_ref = nf.OccupancyGrid(aabb, 128, nf.ContractionType.AABB)
        if grid_kind == "shell":
            grid_ref.binary = sphere_shell_binary(128, 1.5)
        elif grid_kind == "random":
            grid_ref.binary = torch.rand(128, 128, 128, generator=g) < 0.3
        grid_gpu = na.OccupancyGrid(aabb, 128, na.ContractionType.AABB).cuda()
        grid_gpu.set_binary(grid_ref.binary)
    ri_ref, ts_ref, te_ref, pk_ref = nf.ray_marching(o, d, scene_aabb=aabb, grid=grid_ref, render_step_size=step,
                                                     stratified=True, stratified_u=u, return_packed=True)
    ri, ts, te, pk = na.ray_marching(o.cuda(), d.cuda(), scene_aabb=aabb.cuda(), grid=grid_gpu, render_step_size=step,
                                     stratified=True, stratified_u=u.cuda(), return_packed=True)
    _bit_equal(pk, pk_ref, "packed_info")
    _bit_equal(ri, ri_ref, "ray_indices")
    _bit_equal(ts, ts_ref, "t_starts")
    _bit_equal(te, te_ref, "t_ends")
    if grid_kind == "empty":
        assert ri.numel() == 0
    else:
        assert ri.numel() > 1000
        # size-independent structure: sorted by ray, increasing along each ray
        assert bool((ri[1:] >= ri[:-1]).all())
        same = ri[1:] == ri[:-1]
        assert bool((ts[1:, 0][same] >= te[:-1, 0][same] - 1e-6).all())


def test_march_background_bit_exact(cuda_lib):
    """nerfacc.ray_marching as called at models/neus.py:159-169 (UN_BOUNDED_SPHERE grid, cone stepping,
    per-ray near plane) + render_visibility filtering on identical alphas."""
    from instant_angelo_b200 import nerfacc_api as na
    from instant_angelo_b200 import ops
    n = 1024
    o, d, g = _rays(n, 3, outside_frac=0.0)
    aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    _, t_max = nf.ray_aabb_intersect(o, d, aabb)
    near = torch.where(t_max > 1e9, torch.tensor(0.1), t_max)
    cone = 10 ** (math.log10(1e3) / 256) - 1.0
    u = torch.rand(n, generator=g)
    grid_ref = nf.OccupancyGrid(aabb, 256, nf.ContractionType.UN_BOUNDED_SPHERE)
    grid_ref.binary = torch.rand(256, 256, 256, generator=g) < 0.5
    grid_gpu = na.OccupancyGrid(aabb, 256, na.ContractionType.UN_BOUNDED_SPHERE).cuda()
    grid_gpu.set_binary(grid_ref.binary)
    kw = dict(scene_aabb=None, render_step_size=0.01, stratified=True, cone_angle=cone, far_plane=1e3, return_packed=True)
    ri_ref, ts_ref, te_ref, pk_ref = nf.ray_marching(o, d, grid=grid_ref, near_plane=near, stratified_u=u, **kw)
    ri, ts, te, pk = na.ray_marching(o.cuda(), d.cuda(), grid=grid_gpu, near_plane=near.cuda(), stratified_u=u.cuda(), **kw)
    _bit_equal(pk, pk_ref, "packed_info")
    _bit_equal(ri, ri_ref, "ray_indices")
    _bit_equal(ts, ts_ref, "t_starts")
    _bit_equal(te, te_ref, "t_ends")
    assert ri.numel() > 10000
    # visibility on identical alphas
    alphas = torch.rand(ri.numel(), generator=g) * 0.2
    vis_ref = nf.render_visibility(alphas, packed_info=pk_ref, early_stop_eps=1e-4, alpha_thre=0.0)
    vis = ops.visibility(alphas.cuda(), pk, 1e-4, 0.0)
    assert np.array_equal(vis.cpu().numpy(), vis_ref.numpy())
    assert 0 < int(vis_ref.sum()) < vis_ref.numel()
    vis_ref2 = nf.render_visibility(alphas, packed_info=pk_ref, early_stop_eps=1e-3, alpha_thre=0.05)
    vis2 = ops.visibility(alphas.cuda(), pk, 1e-3, 0.05)
    assert np.array_equal(vis2.cpu().numpy(), vis_ref2.numpy())


def test_march_scan_large(cuda_lib):
    """prefix sum over 300k rays (multi-chunk path of the single-CTA scan) against numpy."""
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    n = 300_001
    num = torch.randint(0, 600, (n,), dtype=torch.int32)
    num_g = num.cuda()
    packed = torch.empty(n, 2, dtype=torch.int32, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    L.check(cuda_lib.ia_march_scan(num_g.data_ptr(), n, packed.data_ptr(), total.data_ptr(), None, L.stream()))
    out = C.c_int64(0)
    L.check(cuda_lib.ia_march_total(total.data_ptr(), C.byref(out), L.stream()))
    cum = np.cumsum(num.numpy().astype(np.int64))
    assert out.value == cum[-1]
    assert np.array_equal(packed[:, 1].cpu().numpy(), num.numpy())
    assert np.array_equal(packed[:, 0].cpu().numpy().astype(np.int64), cum - num.numpy())


# ---------------------------------------------------------------------------------------------
# occupancy grid update: BIT-EXACT on identical occupancy values
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("res", [32, 128])
def test_occupancy_update_bit_exact(cuda_lib, res):
    from instant_angelo_b200 import nerfacc_api as na
    aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    g = torch.Generator().manual_seed(res)
    ref = nf.OccupancyGrid(aabb, res, nf.ContractionType.AABB)
    gpu = na.OccupancyGrid(aabb, res, na.ContractionType.AABB).cuda()
    gpu.train()
    nc = ref.num_cells
    for it in range(4):
        if it < 2:       # warm-up: every cell
            idx = torch.arange(nc)
            occ = torch.rand(nc, generator=g) ** 8 * (0.05 if it == 0 else 0.002)
            idx_gpu = None
        else:            # later: random subset with duplicates
            idx = torch.randint(nc, (nc // 2,), generator=g)
            occ = torch.rand(nc // 2, generator=g) ** 4 * 0.01
            idx_gpu = idx.cuda()
        ref.apply_update(idx, occ, occ_thre=0.001, ema_decay=0.95)
        gpu._update(step=0, occ_eval_fn=lambda pts: occ.cuda(), occ_thre=0.001, ema_decay=0.95,
                    indices=idx_gpu, jitter=torch.zeros(idx.numel(), 3, device="cuda"))
        _bit_equal(gpu.occs, ref.occs, f"occs it{it}")
        assert np.array_equal(gpu.binary.cpu().numpy(), ref.binary.numpy()), f"binary it{it}"
        # the packed bitfield is the same grid
        bits = np.unpackbits(gpu.bitfield.cpu().numpy().view(np.uint8), bitorder="little")[:nc]
        assert np.array_equal(bits.astype(bool), ref.binary.numpy().reshape(-1)), f"bitfield it{it}"
        assert 0 < int(ref.binary.sum()) < nc


def test_occupancy_cell_points(cuda_lib):
    from instant_angelo_b200 import nerfacc_api as na
    aabb = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
    g = torch.Generator().manual_seed(9)
    for ctype_ref, ctype in [(nf.ContractionType.AABB, na.ContractionType.AABB),
                             (nf.ContractionType.UN_BOUNDED_SPHERE, na.ContractionType.UN_BOUNDED_SPHERE)]:
        ref = nf.OccupancyGrid(aabb, 16, ctype_ref)
        gpu = na.OccupancyGrid(aabb, 16, ctype).cuda()
        idx = torch.randint(ref.num_cells, (500,), generator=g)
        jit = torch.rand(500, 3, generator=g)
        idx_r, pts_r = ref.cell_points(idx, jit)
        idx_g, pts_g = gpu.cell_points(idx.cuda(), jit.cuda())
        assert np.array_equal(idx_g.cpu().numpy(), idx_r.numpy())
        assert_close(pts_g, pts_r, rtol=1e-5, atol=1e-5, name="cell points")


# ---------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------
def _packed(n_rays, g, max_n=90):
    num = torch.randint(0, max_n, (n_rays,), generator=g)
    num[0] = 0
    num[1] = 1
    num[2] = 32
    num[3] = 33
    num[4] = 64
    cum = torch.cumsum(num, 0)
    return torch.stack([cum - num, num], 1).int(), int(cum[-1])


def test_composite_neus(cuda_lib):
    """get_alpha + render_weight_from_alpha + 4x accumulate_along_rays (models/neus.py:117-139, 234-239)."""
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(21)
    R = 300
    packed, S = _packed(R, g)
    ri = nf.unpack_info(packed)
    mk = lambda *s: torch.randn(*s, generator=g)
    sdf = (mk(S) * 0.05).requires_grad_(True)
    normal = torch.nn.functional.normalize(mk(S, 3), dim=-1).requires_grad_(True)
    dirs = torch.nn.functional.normalize(mk(S, 3), dim=-1)
    dists = torch.rand(S, generator=g) * 0.02 + 0.005
    rgb = torch.rand(S, 3, generator=g).requires_grad_(True)
    tmid = torch.rand(S, generator=g) * 3
    var = torch.tensor(0.3, requires_grad=True)
    go, gd, gc, gn, gw = mk(R), mk(R), mk(R, 3), mk(R, 3), mk(S)
    for anneal in (0.0, 0.37, 1.0):
        model = mr.RefNeuSModel.__new__(mr.RefNeuSModel)
        torch.nn.Module.__init__(model)
        model.variance = mr.RefVarianceNetwork({"init_val": 0.3})
        model.variance.variance = torch.nn.Parameter(var.detach().clone())
        model.cos_anneal_ratio = anneal
        for t in (sdf, normal, rgb):
            t.grad = None
        alpha = model.get_alpha(sdf, normal, dirs, dists)[:, None]
        w = nf.render_weight_from_alpha(alpha, ray_indices=ri, n_rays=R)
        op = nf.accumulate_along_rays(w, ri, None, R)
        dp = nf.accumulate_along_rays(w, ri, tmid[:, None], R)
        cr = nf.accumulate_along_rays(w, ri, rgb, R)
        cn = nf.accumulate_along_rays(w, ri, normal, R)
        loss = (op[:, 0] * go).sum() + (dp[:, 0] * gd).sum() + (cr * gc).sum() + (cn * gn).sum() + (w[:, 0] * gw).sum()
        loss.backward()

        c = lambda t: t.detach().cuda()
        sdf_g, nrm_g, rgb_g = c(sdf).requires_grad_(True), c(normal).requires_grad_(True), c(rgb).requires_grad_(True)
        var_g = c(var).requires_grad_(True)
        inv_s = torch.exp(var_g * 10.0).reshape(1).clip(1e-6, 1e6)
        W, OP, DP, CR, CN, A = ops.composite_neus(sdf_g, nrm_g, c(dirs), c(dists), inv_s, anneal, packed.cuda(),
                                                  t_mid=c(tmid), rgb=rgb_g, nrm=nrm_g)
        lg = (OP * c(go)).sum() + (DP * c(gd)).sum() + (CR * c(gc)).sum() + (CN * c(gn)).sum() + (W * c(gw)).sum()
        lg.backward()
        torch.cuda.synchronize()
        assert_close(A, alpha[:, 0], rtol=1e-3, atol=2e-6, name="alpha")
        assert_close(W, w[:, 0], rtol=1e-3, atol=2e-6, name="weights")
        assert_close(OP, op[:, 0], rtol=1e-4, atol=1e-5, name="opacity")
        assert_close(DP, dp[:, 0], rtol=1e-4, atol=1e-5, name="depth")
        assert_close(CR, cr, rtol=1e-4, atol=1e-5, name="comp_rgb")
        assert_close(CN, cn, rtol=1e-4, atol=1e-5, name="comp_normal")
        for got, want, nm in [(sdf_g.grad, sdf.grad, "d sdf"), (nrm_g.grad, normal.grad, "d normal"),
                              (rgb_g.grad, rgb.grad, "d rgb"), (var_g.grad, model.variance.variance.grad, "d variance")]:
            rt, at = grad_tol(want, 1e-3)
            assert_close(got, want, rtol=rt, atol=at, name=f"{nm} (anneal={anneal})")


def test_composite_density_and_alpha(cuda_lib):
    """render_weight_from_density / render_weight_from_alpha / accumulate (models/neus.py:181-184)."""
    from instant_angelo_b200 import nerfacc_api as na
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(22)
    R = 257
    packed, S = _packed(R, g, max_n=300)
    ri = nf.unpack_info(packed)
    sigma = (torch.rand(S, generator=g) * 3).requires_grad_(True)
    ts = torch.rand(S, generator=g)
    te = ts + torch.rand(S, generator=g) * 0.1
    rgb = torch.rand(S, 3, generator=g).requires_grad_(True)
    tmid = (ts + te) / 2
    go, gd, gc = torch.randn(R, generator=g), torch.randn(R, generator=g), torch.randn(R, 3, generator=g)
    w = nf.render_weight_from_density(ts[:, None], te[:, None], sigma[:, None], ray_indices=ri, n_rays=R)
    op = nf.accumulate_along_rays(w, ri, None, R)
    dp = nf.accumulate_along_rays(w, ri, tmid[:, None], R)
    cr = nf.accumulate_along_rays(w, ri, rgb, R)
    ((op[:, 0] * go).sum() + (dp[:, 0] * gd).sum() + (cr * gc).sum()).backward()
    c = lambda t: t.detach().cuda()
    sg, rg = c(sigma).requires_grad_(True), c(rgb).requires_grad_(True)
    W, OP, DP, CR, _, _ = ops.composite_density(sg, c(ts), c(te), packed.cuda(), t_mid=c(tmid), rgb=rg)
    ((OP * c(go)).sum() + (DP * c(gd)).sum() + (CR * c(gc)).sum()).backward()
    assert_close(W, w[:, 0], rtol=1e-4, atol=1e-6, name="weights")
    assert_close(OP, op[:, 0], rtol=1e-4, atol=1e-5, name="opacity")
    assert_close(CR, cr, rtol=1e-4, atol=1e-5, name="comp_rgb")
    rt, at = grad_tol(sigma.grad, 1e-3)
    assert_close(sg.grad, sigma.grad, rtol=rt, atol=at, name="d sigma")
    rt, at = grad_tol(rgb.grad, 1e-3)
    assert_close(rg.grad, rgb.grad, rtol=rt, atol=at, name="d rgb")
    # nerfacc-style free functions
    al = torch.rand(S, 1, generator=g).requires_grad_(True)
    w_ref = nf.render_weight_from_alpha(al, ray_indices=ri, n_rays=R)
    gw = torch.randn(S, 1, generator=g)
    w_ref.backward(gw)
    alg = c(al).requires_grad_(True)
    w2 = na.render_weight_from_alpha(alg, ray_indices=ri.cuda().int(), n_rays=R)
    w2.backward(gw.cuda())
    assert_close(w2, w_ref, rtol=1e-4, atol=1e-6, name="render_weight_from_alpha")
    rt, at = grad_tol(al.grad, 1e-3)
    assert_close(alg.grad, al.grad, rtol=rt, atol=at, name="d alpha")
    acc = na.accumulate_along_rays(w2.detach(), ri.cuda(), c(rgb), R)
    assert_close(acc, nf.accumulate_along_rays(w_ref.detach(), ri, rgb.detach(), R), rtol=1e-4, atol=1e-5, name="accumulate")
    # known answer: constant alpha a over n samples => weights a(1-a)^i, opacity 1-(1-a)^n
    a, n = 0.25, 40
    pk = torch.tensor([[0, n]], dtype=torch.int32).cuda()
    Wk, OPk, *_ = ops.composite_alpha(torch.full((n,), a, device="cuda"), pk)
    assert_close(Wk, a * (1 - a) ** torch.arange(n, dtype=torch.float64), rtol=1e-5, atol=1e-7, name="geometric series")
    assert abs(float(OPk[0]) - (1 - (1 - a) ** n)) < 1e-6


def test_adamw_matches_torch(cuda_lib):
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(1)
    n = 100_003
    p0 = torch.randn(n, generator=g)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=0.01, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.01)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(n, generator=g) * (0.0 if step == 2 else 1.0)   # step 2: all-zero grads still decay
        p_ref.grad = grad.clone()
        opt.step()
        ops.adamw_step(p, grad.cuda(), m, v, 0.01, 0.9, 0.99, 1e-15, 0.01, step)
    assert_close(p, p_ref, rtol=1e-5, atol=1e-6, name="adamw")


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_feat,n_enc", [(1, 65, 16), (1000, 65, 16), (4097, 13, 9)])
def test_sdf_head_matches_torch(cuda_lib, n, n_feat, n_enc):
    """ops.sdf_head = linear64 into the colour-head input row + cat[., pts*2-1, enc, normal] (reference
    models/geometry.py:206-207, models/texture.py:26-27), with the sdf / diffuse columns' gradients merged in backward."""
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(n + n_feat)
    mk = lambda *s: torch.randn(*s, generator=g).cuda().requires_grad_(True)
    h, W, b, pts, enc, nrm = mk(n, 64), mk(n_feat, 64), mk(n_feat), mk(n, 3), mk(n, n_enc), mk(n, 3)
    tin, sdf, raw = ops.sdf_head(h, W, b, pts, enc, nrm)
    ld = n_feat + 6 + n_enc
    wt, ws, wr = torch.randn(n, ld, generator=g).cuda(), torch.randn(n, generator=g).cuda(), torch.randn(n, 3, generator=g).cuda()
    ((tin * wt).sum() + (sdf * ws).sum() + (raw * wr).sum()).backward()
    got = [t.grad.clone() for t in (h, W, b, pts, enc, nrm)]
    for t in (h, W, b, pts, enc, nrm):
        t.grad = None
    out = h.double() @ W.double().T + b.double()
    tin_ref = torch.cat([out, pts.double() * 2 - 1, enc.double(), nrm.double()], dim=1)
    ((tin_ref * wt).sum() + (out[:, 0] * ws).sum() + (out[:, 1:4] * wr).sum()).backward()
    assert_close(tin, tin_ref.float(), rtol=1e-5, atol=1e-5, name="tin")
    assert_close(sdf, out[:, 0].float(), rtol=1e-5, atol=1e-5, name="sdf")
    assert_close(raw, out[:, 1:4].float(), rtol=1e-5, atol=1e-5, name="raw")
    for name, a, t in zip(("dh", "dW", "db", "dpts", "denc", "dnrm"), got, (h, W, b, pts, enc, nrm)):
        ref = t.grad.float()
        assert_close(a, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()) + 1e-6, name=name)


@pytest.mark.gpu
@pytest.mark.parametrize("n,n_enc", [(1, 16), (1003, 16), (70_001, 9)])
def test_colour_in_matches_torch(cuda_lib, n, n_enc):
    """ia_colour_in_fwd/bwd (geometry output layer folded into the colour network's first layer): the assembled row, the four
    geometry outputs that are used outside the colour network, and every gradient against plain torch."""
    from instant_angelo_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n)
    h = torch.randn(n, 64, device="cuda", generator=g).requires_grad_(True)
    W4 = (torch.randn(4, 64, device="cuda", generator=g) * 0.2).requires_grad_(True)
    b4 = torch.randn(4, device="cuda", generator=g).requires_grad_(True)
    pts = torch.rand(n, 3, device="cuda", generator=g).requires_grad_(True)
    enc = torch.randn(n, n_enc, device="cuda", generator=g).requires_grad_(True)
    nrm = torch.randn(n, 3, device="cuda", generator=g).requires_grad_(True)
    ld = (64 + 3 + n_enc + 3 + 3) // 4 * 4
    tin, sdf, rgb = ops.colour_in(h, W4, b4, pts, enc, nrm, ld)
    ct, cs, cr = (torch.randn(n, ld, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g),
                  torch.randn(n, 3, device="cuda", generator=g))
    ((tin * ct).sum() + (sdf * cs).sum() + (rgb * cr).sum()).backward()
    got = [t.grad.clone() for t in (h, W4, b4, pts, enc, nrm)]
    for t in (h, W4, b4, pts, enc, nrm):
        t.grad = None
    out4 = h.double() @ W4.double().t() + b4.double()
    tin_ref = torch.cat([h.double(), pts.double() * 2 - 1, enc.double(), nrm.double(),
                         torch.zeros(n, ld - 64 - 6 - n_enc, device="cuda", dtype=torch.float64)], 1)
    ((tin_ref * ct.double()).sum() + (out4[:, 0] * cs.double()).sum() + (out4[:, 1:4] * cr.double()).sum()).backward()
    assert_close(tin, tin_ref, rtol=1e-6, atol=1e-6, name="tin")
    assert_close(sdf, out4[:, 0], rtol=1e-5, atol=1e-5, name="sdf")
    assert_close(rgb, out4[:, 1:4], rtol=1e-5, atol=1e-5, name="rgb_raw")
    for name, a, t in zip(("dh", "dW4", "db4", "dpts", "denc", "dnormal"), got, (h, W4, b4, pts, enc, nrm)):
        rt, at = grad_tol(t.grad, 2e-5)
        assert_close(a, t.grad, rtol=rt, atol=at, name=name)


def test_empty_inputs_are_accepted_everywhere(cuda_lib):
    """Zero rows / zero rays (a batch whose rays all miss): every operator returns an empty result of the right shape and
    a zero / empty gradient, as the nerfacc and tcnn bindings do (the C ABI sees NULL data pointers with n = 0)."""
    from instant_angelo_b200 import _lib as L, ops
    dev = "cuda"
    z = lambda *s: torch.zeros(*s, device=dev, requires_grad=True)
    plan = ops.make_grid_plan(16, 2, 19, 32, 1.3195079107728942)
    table = (torch.rand(plan.n_params, device=dev) * 1e-2).requires_grad_(True)
    for group in (1, 6):
        x = z(0, 3)
        y = ops.hashgrid_encode(x, table, plan, 16, group=group)
        assert y.shape == (0, 32)
        y.sum().backward()
        assert float(table.grad.abs().max()) == 0.0
    assert ops.sh_encode(z(0, 3), 4).shape == (0, 16)
    for prec in (L.IA_MLP_FP32, L.IA_MLP_TC_F16):
        desc = ops.make_mlp_desc(3, 32, 2, 65, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, prec)
        params = (torch.randn(L.load().ia_mlp_param_count(desc), device=dev) * 0.1).requires_grad_(True)
        for nou in (1, 65):
            out = ops.mlp_apply(z(0, 3), z(0, 32), params, desc, nou)
            assert out.shape == (0, nou)
            out.sum().backward()
            assert float(params.grad.abs().max()) == 0.0
    W, b = (torch.randn(65, 64, device=dev)).requires_grad_(True), z(65)
    tin, sdf, raw = ops.sdf_head(z(0, 64), W, b, z(0, 3), z(0, 16), z(0, 3))
    assert tin.shape == (0, 87) and sdf.shape == (0,) and raw.shape == (0, 3)
    (tin.sum() + sdf.sum() + raw.sum()).backward()
    assert float(W.grad.abs().max()) == 0.0
    taps = ops.fd_taps(z(0, 3), 1e-3, 1.5)
    assert taps.shape == (0, 6, 3)
    assert ops.fd_grad(z(0, 6), 1e-3).shape == (0, 3)
    nrm, sh = ops.curv_shift(z(0, 3), z(0, 3), z(0, 3), 1e-3)
    assert ops.curv_angle(nrm, sh).shape[0] == 0
    # zero rays
    o, d = torch.zeros(0, 3, device=dev), torch.zeros(0, 3, device=dev)
    t0, t1 = ops.aabb_intersect(o, d, [-1.5] * 3 + [1.5] * 3)
    assert t0.shape == (0,)
    gd = ops.make_grid_desc([-1.5] * 3 + [1.5] * 3, [128] * 3, L.IA_AABB)
    bitfield = torch.full((128 ** 3 // 32,), -1, dtype=torch.int32, device=dev)
    packed, ri, ts, te = ops.march(o, d, t0, t1, gd, bitfield, 0.01, 0.0)
    assert packed.shape == (0, 2) and ri.numel() == 0
    # rays without samples: compositing of empty segments gives zero opacity and passes zero gradients
    packed = torch.zeros(4, 2, dtype=torch.int32, device=dev)
    w, op, dep, rgb, n_, _ = ops.composite_neus(z(0), z(0, 3), z(0, 3), z(0), torch.ones(1, device=dev, requires_grad=True), 1.0, packed,
                                                t_mid=z(0), rgb=z(0, 3), nrm=z(0, 3))
    assert w.numel() == 0 and float(op.abs().max()) == 0.0 and rgb.shape == (4, 3)
    (op.sum() + rgb.sum()).backward()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg_name", ["sparse_2p19", "small_mixed"])
@pytest.mark.parametrize("active", [None, 4])
def test_hashgrid_input_grad_second_order(cuda_lib, cfg_name, active):
    """ops.hashgrid_input_grad = J(x; table)^T dy (the analytic SDF normal, reference models/geometry.py:214-218) and its
    own adjoints (ia_hashgrid_jvp, ia_hashgrid_bwd_input_bwd_table) against torch's double backward through the oracle."""
    from instant_angelo_b200 import ops
    cfg = GRID_CFGS[cfg_name]
    plan_ref, plan = tc.grid_plan(**cfg), ops.make_grid_plan(**cfg)
    act = active if active is None else min(active, cfg["n_levels"])
    n = 600
    g = torch.Generator().manual_seed(21)
    x = _points(n, 13)[4:].clone()                       # drop the points that sit exactly on cell faces
    n = x.shape[0]
    table = (torch.randn(plan_ref.n_params, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(n, plan_ref.n_output_dims, generator=g).requires_grad_(True)
    v = torch.randn(n, 3, generator=g)

    xr = x.clone().requires_grad_(True)
    y_ref = tc.hashgrid_forward(xr, table, plan_ref, act)
    (dx_ref,) = torch.autograd.grad(y_ref, xr, dy, create_graph=True)
    (dx_ref * v).sum().backward()

    tg = table.detach().cuda().requires_grad_(True)
    dg = dy.detach().cuda().requires_grad_(True)
    dx = ops.hashgrid_input_grad(x.cuda(), tg, dg, plan, act)
    (dx * v.cuda()).sum().backward()
    torch.cuda.synchronize()
    rt, at = grad_tol(dx_ref.detach(), 1e-4)
    assert_close(dx, dx_ref.detach(), rtol=rt, atol=at, name="dx")
    rt, at = grad_tol(dy.grad, 1e-4)
    assert_close(dg.grad, dy.grad, rtol=rt, atol=at, name="d(dy) = J v")
    if act is not None:
        assert torch.count_nonzero(dg.grad[:, act * 2:]) == 0, "masked levels receive exactly zero"
    rt, at = grad_tol(table.grad, 1e-4)
    assert_close(tg.grad, table.grad, rtol=rt, atol=at, name="d(table)")


@pytest.mark.gpu
def test_weightnorm_flat_matches_torch(cuda_lib):
    """ops.weightnorm_flat = cat[(g * v / ||v||_row).flatten(), b, ...] (torch weight_norm folded, reference
    models/network_utils.py:115-134) and its adjoint, mixed weight-normed / plain layers."""
    from instant_angelo_b200 import ops
    g_ = torch.Generator().manual_seed(3)
    shapes = [(64, 35, True), (64, 64, True), (65, 64, True)], [(64, 87, False), (64, 64, False), (3, 64, False)], [(64, 24, True), (8, 64, False)]
    for layers in shapes:
        leaves, ref_leaves = [], []
        for n_out, n_in, wn in layers:
            v = torch.randn(n_out, n_in, generator=g_)
            b = torch.randn(n_out, generator=g_)
            gg = torch.rand(n_out, 1, generator=g_) + 0.5 if wn else None
            ref_leaves.append(tuple(None if t is None else t.clone().double().requires_grad_(True) for t in (gg, v, b)))
            leaves.append(tuple(None if t is None else t.clone().cuda().requires_grad_(True) for t in (gg, v, b)))
        flat = ops.weightnorm_flat(leaves)
        parts = []
        for gg, v, b in ref_leaves:
            parts += [(v if gg is None else v * (gg / v.norm(dim=1, keepdim=True))).reshape(-1), b]
        flat_ref = torch.cat(parts)
        w = torch.randn(flat_ref.numel(), generator=g_)
        (flat * w.cuda()).sum().backward()
        (flat_ref * w.double()).sum().backward()
        assert_close(flat, flat_ref.float(), rtol=1e-6, atol=1e-6, name="flat")
        for (gg, v, b), (gr, vr, br) in zip(leaves, ref_leaves):
            for a, r, nm in ((gg, gr, "dg"), (v, vr, "dv"), (b, br, "db")):
                if a is not None:
                    assert_close(a.grad, r.grad.float(), rtol=1e-5, atol=1e-6, name=nm)


@pytest.mark.parametrize("with_mask,with_curv,n_rays,n_samples", [(False, True, 777, 50021), (True, False, 64, 300), (True, True, 1, 1),
                                                                   (False, True, 3000, 0)])
def test_neus_losses_match_tensor_expressions(cuda_lib, with_mask, with_curv, n_rays, n_samples):
    """ops.neus_losses (ia_neus_losses_fwd / _bwd: every per-ray and per-sample loss term of reference systems/neus.py:132-160
    and their weighted sum, one launch each way) against the same expressions as float64 tensor operators with autograd:
    values to 1e-5, gradients to 1e-4 of each tensor's scale; clamp edges of the opacity, invalid rays, exact zeros of
    sdf / laplace (sign(0) = 0) and a vanishing sdf gradient included."""
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(n_rays + n_samples)
    comp = torch.rand(n_rays, 3, generator=g) * 1.2
    gt = torch.rand(n_rays, 3, generator=g)
    valid = torch.rand(n_rays, generator=g) > 0.2
    if n_rays > 4:
        valid[0], valid[1] = True, False
    opacity = torch.rand(n_rays, 1, generator=g)
    if n_rays > 8:
        opacity[:6, 0] = torch.tensor([0.0, 1.0, 1.001e-3, 1.0 - 1.001e-3, 5e-4, 0.99999])    # both sides of the clamp edges
    mask = (torch.rand(n_rays, generator=g) > 0.5) if with_mask else None
    sgrad = torch.randn(n_samples, 3, generator=g) * 1.5
    sdf = torch.randn(n_samples, generator=g) * 0.01
    lap = torch.rand(n_samples, 1, generator=g) - 0.3
    if n_samples > 8:
        sgrad[0] = 0.0
        sdf[1] = 0.0
        lap[2] = 0.0
    lambdas = {"rgb_mse": 10.0, "rgb_l1": 0.5, "eikonal": 0.1, "mask": 0.7 if with_mask else 0.0, "opaque": 0.3, "sparsity": 0.2,
               "curvature": 5e-4 if with_curv else 0.0}
    scale = 1.0

    def expr(comp, opacity, sgrad, sdf, lap, dt):
        diff = torch.where(valid[:, None], comp - gt.to(dt), torch.zeros((), dtype=dt))
        n_valid = valid.sum().to(dt) * 3
        t = {"rgb_mse": (diff * diff).sum() / n_valid, "rgb_l1": diff.abs().sum() / n_valid,
             "eikonal": ((torch.linalg.norm(sgrad, ord=2, dim=-1) - 1.0) ** 2).mean()}
        o = torch.clamp(opacity.squeeze(-1), 1.0e-3, 1.0 - 1.0e-3)
        bce = lambda a, b: -(b * torch.log(a) + (1 - b) * torch.log(1 - a)).mean()
        if with_mask:
            t["mask"] = bce(o, mask.to(dt))
        t["opaque"] = bce(o, o)
        t["sparsity"] = torch.exp(-scale * sdf.abs()).mean()
        if with_curv:
            t["curvature"] = lap.abs().mean()
        return sum(t[k] * lambdas[k] for k in t), t

    leaves64 = [v.double().requires_grad_(True) for v in (comp, opacity, sgrad, sdf, lap)]
    loss64, terms64 = expr(*leaves64, torch.float64)
    leaves = [v.cuda().requires_grad_(True) for v in (comp, opacity, sgrad, sdf, lap)]
    loss, terms = ops.neus_losses(leaves[0], gt.cuda(), valid.cuda(), leaves[1], mask.cuda().float() if with_mask else None, leaves[2],
                                  leaves[3], leaves[4] if with_curv else None, lambdas, scale)
    assert set(terms) == set(terms64)
    if n_samples == 0:
        assert torch.isnan(loss) and torch.isnan(loss64)           # means over zero samples, as in the reference
        return
    for k in terms64:
        assert not terms[k].requires_grad
        assert_close(terms[k], terms64[k], rtol=1e-5, atol=1e-7, name=k)
    assert_close(loss, loss64, rtol=1e-5, atol=1e-7, name="loss")
    up = 0.37
    (loss * up).backward()
    (loss64 * up).backward()
    for name, a, b in zip(("comp_rgb", "opacity", "sdf_grad", "sdf", "laplace"), leaves, leaves64):
        if name == "laplace" and not with_curv:
            assert a.grad is None
            continue
        rt, at = grad_tol(b.grad, 1e-4, floor=1e-9)
        assert_close(a.grad, b.grad, rtol=rt, atol=at, name="d " + name)


def test_ray_samples_bit_exact_and_normalize3(cuda_lib):
    """ops.ray_samples == the tensor expressions of reference models/neus.py:153-157 / 218-223 bit for bit (the background
    marcher prunes on a density evaluated at these positions); ops.normalize3 == F.normalize (values to 1 ulp, adjoint to
    1e-6), including rows at and below eps."""
    from instant_angelo_b200 import ops
    g = torch.Generator().manual_seed(9)
    R, S = 300, 20011
    rays_o = (torch.randn(R, 3, generator=g) * 2).cuda()
    rays_d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).cuda()
    ri = torch.sort(torch.randint(0, R, (S,), generator=g)).values.int().cuda()
    t0 = (torch.rand(S, 1, generator=g) * 5).cuda()
    t1 = t0 + (torch.rand(S, 1, generator=g) * 0.02).cuda()
    pos, dirs, mid, dist = ops.ray_samples(rays_o, rays_d, ri, t0, t1)
    l = ri.long()
    assert torch.equal(mid, (t0 + t1) / 2.0) and torch.equal(dist, t1 - t0) and torch.equal(dirs, rays_d[l])
    assert torch.equal(pos, rays_o[l] + rays_d[l] * ((t0 + t1) / 2.0))
    assert torch.equal(pos, rays_o[l] + rays_d[l] * (t0 + t1) / 2.0)          # the marcher's sigma_fn spelling
    only = ops.ray_samples(rays_o, rays_d, ri, t0, t1, False, False, False)
    assert torch.equal(only[0], pos) and only[1] is None and only[2] is None and only[3] is None
    empty = ops.ray_samples(rays_o, rays_d, ri[:0], t0[:0], t1[:0])
    assert empty[0].shape == (0, 3) and empty[2].shape == (0, 1)

    x = torch.randn(5003, 3, generator=g)
    x[0] = 0.0
    x[1] = torch.tensor([1e-13, 0.0, 0.0])
    x[2] = torch.tensor([3e-7, -4e-7, 0.0])
    up = torch.randn(5003, 3, generator=g)
    xa, xb = x.cuda().requires_grad_(True), x.double().requires_grad_(True)
    ya, yb = ops.normalize3(xa), torch.nn.functional.normalize(xb, p=2, dim=-1)
    assert_close(ya, yb, rtol=3e-7, atol=1e-12, name="normalize3")
    ya.backward(up.cuda())
    yb.backward(up.double())
    assert_close(xa.grad[3:], xb.grad[3:], rtol=1e-5, atol=1e-6, name="normalize3 adjoint")
    assert_close(xa.grad[:2], xb.grad[:2], rtol=1e-6, atol=0.0, name="normalize3 adjoint below eps")     # g / eps
    rt, at = grad_tol(xb.grad[2:3], 1e-5)
    assert_close(xa.grad[2:3], xb.grad[2:3], rtol=rt, atol=at, name="normalize3 adjoint, tiny row")
    y3 = ops.normalize3(x.cuda().view(5003, 1, 3))
    assert y3.shape == (5003, 1, 3)


def test_contract_matches_tensor_expressions(cuda_lib):
    """ops.contract (ia_contract) == contract_to_unisphere's tensor expressions (reference models/geometry.py:19-31) for
    both contractions, inside and far outside the box; inputs that require grad keep the differentiable expressions."""
    from instant_angelo_b200 import geometry as geo
    from instant_angelo_b200.nerfacc_api import ContractionType
    g = torch.Generator().manual_seed(4)
    x = torch.cat([torch.randn(4001, 3, generator=g), torch.randn(500, 3, generator=g) * 40.0,
                   torch.tensor([[0.0, 0.0, 0.0], [1.5, -1.5, 1.5], [3.0, 0.0, 0.0]])]).cuda()
    for ct in (ContractionType.AABB, ContractionType.UN_BOUNDED_SPHERE):
        fast = geo.contract_to_unisphere(x, 1.5, ct)
        xr = x.clone().requires_grad_(True)
        slow = geo.contract_to_unisphere(xr, 1.5, ct)                    # tensor expressions (differentiable)
        assert slow.requires_grad and not fast.requires_grad and fast.shape == slow.shape
        assert_close(fast, slow, rtol=2e-6, atol=2e-7, name=f"contract {ct.name}")
        if ct == ContractionType.UN_BOUNDED_SPHERE:
            assert float(fast.min()) >= 0.0 and float(fast.max()) <= 1.0
    assert geo.contract_to_unisphere(x.view(-1, 1, 3), 1.5, ContractionType.AABB).shape == (x.shape[0], 1, 3)
    with pytest.raises(RuntimeError):
        from instant_angelo_b200 import ops
        ops.contract(x, 1.5, 1)                                           # UN_BOUNDED_TANH is not on the path


@pytest.mark.parametrize("alpha_thre", [0.0, 0.02])
def test_ray_marching_fused_prune_is_bit_exact(cuda_lib, monkeypatch, alpha_thre):
    """nerfacc.ray_marching with a sigma_fn (the background marcher, reference models/neus.py:144-169): the fused pruning
    (ia_prune_count -> ia_march_scan -> ia_prune_write) returns exactly what the tensor-operator chain returns -- alphas from
    the densities, render_visibility, boolean compaction, pack_info --, including rays that keep a single sample."""
    from instant_angelo_b200.nerfacc_api import ContractionType, OccupancyGrid, ray_marching
    g = torch.Generator().manual_seed(31)
    R = 700
    rays_o = (torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * 0.8).cuda()
    rays_d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).cuda()
    grid = OccupancyGrid(roi_aabb=torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5]), resolution=64,
                         contraction_type=ContractionType.UN_BOUNDED_SPHERE).cuda()
    grid.set_binary(torch.rand(64, 64, 64, generator=g) > 0.4)

    def sigma_fn(t_starts, t_ends, ray_indices):
        ri = ray_indices.long()
        x = rays_o[ri] + rays_d[ri] * (t_starts + t_ends) / 2.0
        s = 40.0 * torch.sin(7.0 * x).prod(dim=-1, keepdim=True).abs() * (x.norm(dim=-1, keepdim=True) < 3.0)
        return torch.where(ri[:, None] % 11 == 0, torch.full_like(s, 1e4), s)      # some rays go opaque at once

    def run():
        return ray_marching(rays_o, rays_d, grid=grid, sigma_fn=sigma_fn, near_plane=0.1, far_plane=1e3, render_step_size=0.01,
                            stratified=False, cone_angle=0.01, alpha_thre=alpha_thre, return_packed=True)

    fused = run()
    monkeypatch.setenv("IA_NO_FUSED_PRUNE", "1")
    plain = run()
    assert fused[0].numel() > 1000 and fused[0].numel() == plain[0].numel()
    for a, b, name in zip(fused, plain, ("ray_indices", "t_starts", "t_ends", "packed_info")):
        assert a.dtype == b.dtype and a.shape == b.shape, name
        assert torch.equal(a, b), name
    assert int((fused[3][:, 1] <= 1).sum()) > 0            # rays that went opaque at their first sample keep only that one


def test_point_losses_match_tensor_expressions(cuda_lib):
    """ops.point_losses (ia_point_losses_fwd / _bwd) against reference systems/neus.py:173-186 as float64 tensor operators:
    sdf_l1 = (F.l1_loss(sdf, 0) * weights).mean() -- a scalar times the weights, Appendix C-11 --, normal_cos, their weighted sum
    and its gradients w.r.t. the SDF values and the SDF gradients (zero SDF value and vanishing gradient rows included)."""
    from instant_angelo_b200 import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(12)
    n = 8192
    sdf = torch.randn(n, generator=g) * 0.02
    grad = torch.randn(n, 3, generator=g)
    ngt = torch.randn(n, 3, generator=g) * 3.0
    w = torch.rand(n, generator=g)
    sdf[0] = 0.0
    grad[1] = 0.0
    lam1, lam2 = 0.37, 0.11
    s64, g64 = sdf.double().requires_grad_(True), grad.double().requires_grad_(True)
    t1 = (F.l1_loss(s64, torch.zeros_like(s64)) * w.double()).mean(dim=0)
    t2 = (1.0 - torch.sum(F.normalize(g64, p=2, dim=-1) * F.normalize(ngt.double(), p=2, dim=-1), dim=-1)).mean()
    tot64 = t1 * lam1 + t2 * lam2
    sc, gc = sdf.cuda().requires_grad_(True), grad.cuda().requires_grad_(True)
    tot, terms = ops.point_losses(sc, gc, ngt.cuda(), w.cuda(), lam1, lam2)
    assert_close(terms["sdf_l1"], t1, rtol=1e-5, atol=1e-9, name="sdf_l1")
    assert_close(terms["normal_cos"], t2, rtol=1e-5, atol=1e-8, name="normal_cos")
    assert_close(tot, tot64, rtol=1e-5, atol=1e-8, name="total")
    (tot * 1.7).backward()
    (tot64 * 1.7).backward()
    rt, at = grad_tol(s64.grad, 1e-5, floor=1e-12)
    assert_close(sc.grad, s64.grad, rtol=rt, atol=at, name="d sdf")
    rt, at = grad_tol(g64.grad[2:], 1e-4, floor=1e-12)
    assert_close(gc.grad[2:], g64.grad[2:], rtol=rt, atol=at, name="d grad")
    assert float(sc.grad[0]) == 0.0 and torch.isfinite(gc.grad).all()
