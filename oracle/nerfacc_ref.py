"""ORACLE (test infrastructure only) -- CPU restatement of the nerfacc==0.3.3 API surface that
Instant-angelo's NeuS path calls (models/neus.py:11-12, 64-74, 108-111, 153, 159-169, 181-184,
209-220, 234-239).  nerfacc is an un-vendored dependency (requirements.txt:3), absent from
/root/reference and not installable here; the algorithms follow SURVEY.md Appendix A.3-A.7.
PARITY UNPINNED by upstream tests (there are none); known-answer tests are authored in tests/.

The per-ray marching loops run in C (oracle/march_ref.c, built into oracle/_build/) because they
must be IEEE-binary32 exact and pure-Python loops are too slow beyond toy sizes; a pure-numpy
float32 version (march_python) restates the same loop independently and is used to cross-check the
C build on small cases.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from enum import IntEnum
from typing import Callable, Optional, Tuple

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "libia_oracle.so")
_lib = None


class ContractionType(IntEnum):
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


def build_c_oracle(force: bool = False) -> str:
    """gcc recipe for the C restatement (also driven by oracle/Makefile)."""
    src = os.path.join(_HERE, "march_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               src, "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


def _c():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _np32(t) -> np.ndarray:
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# A.3 ray_aabb_intersect
# ---------------------------------------------------------------------------------------------

def ray_aabb_intersect(rays_o: torch.Tensor, rays_d: torch.Tensor, aabb: torch.Tensor, clamp_zero: bool = True
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    o, d, bb = _np32(rays_o), _np32(rays_d), _np32(aabb)
    n = o.shape[0]
    tmin = np.empty(n, np.float32)
    tmax = np.empty(n, np.float32)
    _c().ia_ref_aabb(_fp(o), _fp(d), ctypes.c_int64(n), _fp(bb), ctypes.c_int(int(clamp_zero)), _fp(tmin), _fp(tmax))
    return torch.from_numpy(tmin), torch.from_numpy(tmax)


# ---------------------------------------------------------------------------------------------
# A.5 contraction (Python side, used by OccupancyGrid._update)
# ---------------------------------------------------------------------------------------------

def contract_inv(x: torch.Tensor, roi: torch.Tensor, type: ContractionType) -> torch.Tensor:
    """unit cube -> world."""
    lo, hi = roi[:3], roi[3:]
    if type == ContractionType.AABB:
        return x * (hi - lo) + lo
    u = (x - 0.5) * 4.0
    n = u.norm(dim=-1, keepdim=True)
    u = torch.where(n > 1.0, (u / n) / (2.0 - n), u)
    return (u * 0.5 + 0.5) * (hi - lo) + lo


def contract(x: torch.Tensor, roi: torch.Tensor, type: ContractionType) -> torch.Tensor:
    """world -> unit cube (float32 torch ops; the bit-exact form lives in march_ref.c)."""
    lo, hi = roi[:3], roi[3:]
    u = (x - lo) / (hi - lo)
    if type == ContractionType.AABB:
        return u
    v = u * 2.0 - 1.0
    n = v.norm(dim=-1, keepdim=True)
    v = torch.where(n > 1.0, (2.0 - 1.0 / n) * (v / n), v)
    return v * 0.25 + 0.5


# ---------------------------------------------------------------------------------------------
# A.6 OccupancyGrid
# ---------------------------------------------------------------------------------------------

class OccupancyGrid:
    """nerfacc.OccupancyGrid(roi_aabb, resolution, contraction_type) restated.  All randomness
    (cell choice, jitter) can be injected for parity runs."""

    def __init__(self, roi_aabb, resolution: int = 128, contraction_type: ContractionType = ContractionType.AABB):
        self.roi_aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten()
        self.res = [int(resolution)] * 3 if isinstance(resolution, int) else [int(r) for r in resolution]
        self.contraction_type = ContractionType(contraction_type)
        self.num_cells = self.res[0] * self.res[1] * self.res[2]
        self.occs = torch.zeros(self.num_cells, dtype=torch.float32)
        self.binary = torch.zeros(self.res, dtype=torch.bool)

    def grid_coords(self, indices: torch.Tensor) -> torch.Tensor:
        """Integer cell coordinates of flat indices (meshgrid 'ij', x-major: idx = x*ry*rz + y*rz + z)."""
        rx, ry, rz = self.res
        return torch.stack([indices // (ry * rz), (indices // rz) % ry, indices % rz], dim=-1)

    def sample_indices(self, step: int, warmup_steps: int = 256, gen: Optional[torch.Generator] = None) -> torch.Tensor:
        if step < warmup_steps:
            return torch.arange(self.num_cells)
        n = self.num_cells // 4
        uniform = torch.randint(self.num_cells, (n,), generator=gen)
        occupied = torch.nonzero(self.binary.flatten())[:, 0]
        if n < occupied.numel():
            sel = torch.randint(occupied.numel(), (n,), generator=gen)
            occupied = occupied[sel]
        return torch.cat([uniform, occupied], dim=0)

    def cell_points(self, indices: torch.Tensor, jitter: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (indices kept, world points).  jitter: U[0,1) of shape [len(indices),3]."""
        res = torch.tensor(self.res, dtype=torch.float32)
        x = (self.grid_coords(indices).float() + jitter) / res
        if self.contraction_type == ContractionType.UN_BOUNDED_SPHERE:
            keep = (x - 0.5).norm(dim=1) < 0.5
            x, indices = x[keep], indices[keep]
        return indices, contract_inv(x, self.roi_aabb, self.contraction_type)

    def apply_update(self, indices: torch.Tensor, occ: torch.Tensor, occ_thre: float = 1e-2, ema_decay: float = 0.95) -> None:
        """occs[i] = max(occs[i]*decay, max over duplicates of occ); binary = occs > min(mean, thre).
        (nerfacc writes duplicates with index_put, whose winner is unspecified; the max is the
        deterministic resolution used by oracle and CUDA path alike.)  The mean is taken in float64
        and rounded to float32."""
        occ = occ.reshape(-1).float()
        new = torch.zeros_like(self.occs)
        new.scatter_reduce_(0, indices, occ, reduce="amax", include_self=True)
        touched = torch.zeros(self.num_cells, dtype=torch.bool)
        touched[indices] = True
        decayed = self.occs * ema_decay
        self.occs = torch.where(touched, torch.maximum(decayed, new), self.occs)
        mean = self.occs.double().mean().float()
        thr = torch.clamp(mean, max=occ_thre)
        self.binary = (self.occs > thr).reshape(self.res)

    def every_n_step(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                     warmup_steps: int = 256, n: int = 16, indices=None, jitter=None,
                     gen: Optional[torch.Generator] = None) -> None:
        if step % n != 0:
            return
        if indices is None:
            indices = self.sample_indices(step, warmup_steps, gen)
        if jitter is None:
            jitter = torch.rand(indices.numel(), 3, generator=gen)
        indices, pts = self.cell_points(indices, jitter)
        with torch.no_grad():
            occ = occ_eval_fn(pts)
        self.apply_update(indices, occ, occ_thre, ema_decay)


# ---------------------------------------------------------------------------------------------
# A.4 ray_marching
# ---------------------------------------------------------------------------------------------

def _march_c(o, d, tmin, tmax, roi, binary, res, ctype, step, cone):
    n = o.shape[0]
    res_arr = np.asarray(res, dtype=np.int32)
    num = np.zeros(n, np.int32)
    lib = _c()
    args = (_fp(o), _fp(d), _fp(tmin), _fp(tmax), ctypes.c_int64(n), _fp(roi), _fp(binary), _fp(res_arr),
            ctypes.c_int(int(ctype)), ctypes.c_float(step), ctypes.c_float(cone))
    lib.ia_ref_march(*args, None, _fp(num), None, None, None)
    cum = np.cumsum(num, dtype=np.int64)
    total = int(cum[-1]) if n else 0
    packed = np.stack([cum - num, num], axis=1).astype(np.int32)
    ri = np.zeros(total, np.int32)
    ts = np.zeros(total, np.float32)
    te = np.zeros(total, np.float32)
    lib.ia_ref_march(*args, _fp(packed), _fp(num), _fp(ri), _fp(ts), _fp(te))
    return packed, ri, ts, te


def ray_marching(rays_o, rays_d, t_min=None, t_max=None, scene_aabb=None, grid: Optional[OccupancyGrid] = None,
                 sigma_fn=None, alpha_fn=None, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 near_plane=None, far_plane=None, render_step_size: float = 1e-3, stratified: bool = False,
                 cone_angle: float = 0.0, stratified_u: Optional[torch.Tensor] = None, clamp_zero: bool = True,
                 return_packed: bool = False):
    """nerfacc.ray_marching restated; `stratified_u` injects the U[0,1) jitter draw."""
    o, d = _np32(rays_o), _np32(rays_d)
    n = o.shape[0]
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            t_min, t_max = ray_aabb_intersect(rays_o, rays_d, scene_aabb, clamp_zero)
        else:
            t_min = torch.zeros(n)
            t_max = torch.full((n,), 1e10)
    t_min = torch.as_tensor(t_min, dtype=torch.float32)
    t_max = torch.as_tensor(t_max, dtype=torch.float32)
    if near_plane is not None:
        t_min = torch.maximum(t_min, torch.as_tensor(near_plane, dtype=torch.float32))
    if far_plane is not None:
        t_max = torch.minimum(t_max, torch.as_tensor(far_plane, dtype=torch.float32))
    if stratified:
        u = stratified_u if stratified_u is not None else torch.rand(n)
        t_min = t_min + u * render_step_size
    if grid is not None:
        roi, binary, res, ctype = _np32(grid.roi_aabb), grid.binary.numpy().astype(np.uint8), grid.res, grid.contraction_type
    else:
        roi = np.array([-1e10] * 3 + [1e10] * 3, np.float32)
        binary, res, ctype = np.ones(1, np.uint8), [1, 1, 1], ContractionType.AABB
    binary = np.ascontiguousarray(binary.reshape(-1))
    packed, ri, ts, te = _march_c(o, d, _np32(t_min), _np32(t_max), roi, binary, res, ctype,
                                  float(render_step_size), float(cone_angle))
    ray_indices = torch.from_numpy(ri)
    t_starts = torch.from_numpy(ts)[:, None]
    t_ends = torch.from_numpy(te)[:, None]
    packed_info = torch.from_numpy(packed)
    if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (sigma_fn is not None or alpha_fn is not None):
        with torch.no_grad():
            if sigma_fn is not None:
                sigmas = sigma_fn(t_starts, t_ends, ray_indices)
                alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
            else:
                alphas = alpha_fn(t_starts, t_ends, ray_indices)
        masks = render_visibility(alphas, packed_info=packed_info, early_stop_eps=early_stop_eps, alpha_thre=alpha_thre)
        ray_indices, t_starts, t_ends = ray_indices[masks], t_starts[masks], t_ends[masks]
        packed_info = pack_info(ray_indices, n)
    if return_packed:
        return ray_indices, t_starts, t_ends, packed_info
    return ray_indices, t_starts, t_ends


def march_python(o, d, t_min, t_max, roi, binary, res, ctype, step, cone):
    """Independent numpy-float32 restatement of the per-ray loop (slow; tiny cases only)."""
    f = np.float32
    o, d, roi = _np32(o), _np32(d), _np32(roi)
    binary = np.asarray(binary).reshape(-1)

    def fma(a, b, c):
        return f(np.float64(a) * np.float64(b) + np.float64(c))

    def calc_dt(t):
        return min(max(f(t * f(cone)), f(step)), f(1e10))

    def unit(xyz):
        u = [f(f(xyz[k] - roi[k]) / f(roi[3 + k] - roi[k])) for k in range(3)]
        if ctype == ContractionType.UN_BOUNDED_SPHERE:
            v = [f(f(u[k] * f(2)) - f(1)) for k in range(3)]
            nsq = fma(v[2], v[2], fma(v[1], v[1], f(v[0] * v[0])))
            nn = f(np.sqrt(nsq))
            if nn > 1:
                s = f(f(2) - f(f(1) / nn))
                v = [f(s * f(v[k] / nn)) for k in range(3)]
            u = [f(f(v[k] * f(0.25)) + f(0.5)) for k in range(3)]
        return u

    def occupied(xyz):
        if ctype == ContractionType.AABB:
            for k in range(3):
                if xyz[k] < roi[k] or xyz[k] > roi[3 + k]:
                    return False
        u = unit(xyz)
        idx = [min(max(int(f(u[k] * f(res[k]))), 0), res[k] - 1) for k in range(3)]
        return bool(binary[idx[0] * res[1] * res[2] + idx[1] * res[2] + idx[2]])

    def sgn(x):
        return f(1) if x > 0 else (f(-1) if x < 0 else f(0))

    out = []
    with np.errstate(all="ignore"):
        for i in range(o.shape[0]):
            inv = [f(f(1) / d[i, k]) for k in range(3)]
            near, far = f(t_min[i]), f(t_max[i])
            t0 = near
            t1 = f(t0 + calc_dt(t0))
            tm = f(f(t0 + t1) * f(0.5))
            samples = []
            while tm < far:
                xyz = [fma(tm, d[i, k], o[i, k]) for k in range(3)]
                if occupied(xyz):
                    samples.append((t0, t1))
                    t0 = t1
                    t1 = f(t0 + calc_dt(t0))
                    tm = f(f(t0 + t1) * f(0.5))
                elif ctype == ContractionType.AABB:
                    best = []
                    for k in range(3):
                        r = f(res[k])
                        p = f(f(f(xyz[k] - roi[k]) / f(roi[3 + k] - roi[k])) * r)
                        tgt = f(np.floor(f(f(p + f(0.5)) + f(f(0.5) * sgn(d[i, k])))))
                        best.append(f(f(f(f(tgt - p) * inv[k]) / r) * f(roi[3 + k] - roi[k])))
                    tt = np.fmin(np.fmin(best[0], best[1]), best[2])
                    tt = np.fmax(tt, f(0))
                    target = np.fmin(f(tm + tt), far)
                    while True:
                        tm = f(tm + f(step))
                        if not (tm < target):
                            break
                    dt = calc_dt(tm)
                    t0 = f(tm - f(dt * f(0.5)))
                    t1 = f(tm + f(dt * f(0.5)))
                else:
                    t0 = t1
                    t1 = f(t0 + calc_dt(t0))
                    tm = f(f(t0 + t1) * f(0.5))
            out.append(samples)
    return out


# ---------------------------------------------------------------------------------------------
# A.7 compositing
# ---------------------------------------------------------------------------------------------

def pack_info(ray_indices: torch.Tensor, n_rays: int) -> torch.Tensor:
    num = torch.zeros(n_rays, dtype=torch.int64)
    num.index_add_(0, ray_indices.long(), torch.ones_like(ray_indices, dtype=torch.int64))
    cum = torch.cumsum(num, 0)
    return torch.stack([cum - num, num], dim=1).int()


def _exclusive_cumprod_by_ray(one_minus_alpha: torch.Tensor, ray_indices: torch.Tensor, n_rays: int) -> torch.Tensor:
    """T_i = prod_{j<i, same ray} (1-alpha_j), differentiable.  Samples are sorted by ray."""
    s = one_minus_alpha.reshape(-1)
    ri = ray_indices.long()
    packed = pack_info(ri, n_rays).long()
    max_n = int(packed[:, 1].max()) if packed.numel() else 0
    if s.numel() == 0:
        return s.clone()
    # scatter to a padded [R, max_n] matrix, cumprod along rows, gather back
    pos = torch.arange(s.numel()) - packed[ri, 0]
    pad = torch.ones(n_rays, max_n + 1, dtype=s.dtype)
    pad = pad.index_put((ri, pos + 1), s)
    T = torch.cumprod(pad, dim=1)
    return T[ri, pos]


def render_weight_from_alpha(alphas: torch.Tensor, *, packed_info=None, ray_indices=None, n_rays=None) -> torch.Tensor:
    assert alphas.dim() == 2 and alphas.shape[1] == 1
    if ray_indices is None:
        ray_indices = unpack_info(packed_info)
        n_rays = packed_info.shape[0]
    T = _exclusive_cumprod_by_ray(1.0 - alphas, ray_indices, n_rays)
    return T[:, None] * alphas


def render_weight_from_density(t_starts, t_ends, sigmas, *, packed_info=None, ray_indices=None, n_rays=None) -> torch.Tensor:
    alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
    return render_weight_from_alpha(alphas, packed_info=packed_info, ray_indices=ray_indices, n_rays=n_rays)


def unpack_info(packed_info: torch.Tensor) -> torch.Tensor:
    return torch.repeat_interleave(torch.arange(packed_info.shape[0]), packed_info[:, 1].long())


def render_visibility(alphas: torch.Tensor, *, packed_info: torch.Tensor, early_stop_eps: float = 1e-4,
                      alpha_thre: float = 0.0) -> torch.Tensor:
    a = _np32(alphas.reshape(-1))
    pk = np.ascontiguousarray(packed_info.numpy().astype(np.int32))
    vis = np.zeros(a.shape[0], np.uint8)
    _c().ia_ref_visibility(_fp(a), _fp(pk), ctypes.c_int64(pk.shape[0]), ctypes.c_float(early_stop_eps),
                           ctypes.c_float(alpha_thre), _fp(vis))
    return torch.from_numpy(vis.astype(bool))


def accumulate_along_rays(weights: torch.Tensor, ray_indices: torch.Tensor, values: Optional[torch.Tensor] = None,
                          n_rays: Optional[int] = None) -> torch.Tensor:
    assert weights.dim() == 2 and weights.shape[1] == 1
    src = weights if values is None else weights * values
    out = torch.zeros(n_rays, src.shape[1], dtype=src.dtype)
    return out.index_add(0, ray_indices.long(), src)
