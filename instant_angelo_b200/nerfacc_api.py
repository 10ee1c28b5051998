"""The slice of the nerfacc==0.3.3 Python API that Instant-angelo imports (reference models/neus.py:11-12,
models/geometry.py:15), re-implemented on the sm_100a kernels of libia_b200.so.

Same names, argument meaning and error behaviour: non-CUDA inputs raise
NotImplementedError("Only support cuda inputs.") as nerfacc's `_C` stubs do.
Extra keyword arguments (`stratified_u`, `indices`, `jitter`) exist only so that parity tests can inject
the random draws that nerfacc makes internally.
"""
from __future__ import annotations

import os

from enum import IntEnum
from typing import Callable, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class ContractionType(IntEnum):
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


def _contract_inv(x: torch.Tensor, roi: torch.Tensor, type: ContractionType) -> torch.Tensor:
    lo, hi = roi[:3], roi[3:]
    if type == ContractionType.AABB:
        return x * (hi - lo) + lo
    u = (x - 0.5) * 4.0
    n = u.norm(dim=-1, keepdim=True)
    u = torch.where(n > 1.0, (u / n) / (2.0 - n), u)
    return (u * 0.5 + 0.5) * (hi - lo) + lo


class OccupancyGrid(nn.Module):
    """nerfacc.OccupancyGrid(roi_aabb, resolution, contraction_type) (reference models/neus.py:64-74).

    Buffers keep nerfacc's names so reference checkpoints load: `_roi_aabb`, `_binary` (bool [rx,ry,rz]),
    `resolution`, `occs`.  `bitfield` (1 bit / cell) is the form the marching kernels read; it is derived
    state (non-persistent) and is re-packed from `_binary` after load_state_dict.
    """

    NUM_DIM = 3

    def __init__(self, roi_aabb, resolution: Union[int, list] = 128, contraction_type: ContractionType = ContractionType.AABB):
        super().__init__()
        if isinstance(resolution, int):
            resolution = [resolution] * 3
        res = [int(r) for r in resolution]
        self._res = res
        self._contraction_type = ContractionType(contraction_type)
        self.num_cells = res[0] * res[1] * res[2]
        roi = torch.as_tensor(roi_aabb, dtype=torch.float32).flatten().clone()
        self._roi_host = [float(v) for v in roi.tolist()]
        self.register_buffer("_roi_aabb", roi)
        self.register_buffer("resolution", torch.tensor(res, dtype=torch.int32))
        self.register_buffer("occs", torch.zeros(self.num_cells))
        self.register_buffer("_binary", torch.zeros(res, dtype=torch.bool))
        self.register_buffer("bitfield", torch.zeros((self.num_cells + 31) // 32, dtype=torch.int32), persistent=False)
        self._workspace = None
        self.generator = None   # optional torch.Generator for cell sampling/jitter (data-parallel replicas share its seed)
        self._grid_desc = ops.make_grid_desc(self._roi_host, res, int(self._contraction_type))

    # -- nerfacc properties
    @property
    def roi_aabb(self) -> torch.Tensor:
        return self._roi_aabb

    @property
    def binary(self) -> torch.Tensor:
        return self._binary

    @property
    def contraction_type(self) -> ContractionType:
        return self._contraction_type

    @property
    def grid_desc(self) -> L.GridDesc:
        return self._grid_desc

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        if self._binary.is_cuda:
            self.repack()

    def repack(self) -> None:
        """Rebuild the bitfield from `_binary` (after loading a checkpoint or editing the grid by hand)."""
        ops.occ_pack(self._binary.view(torch.uint8).reshape(-1), self.bitfield)

    def set_binary(self, binary: torch.Tensor) -> None:
        self._binary.copy_(binary.reshape(self._res).to(self._binary.device))
        self.repack()

    # -- sampling of the cells to refresh (nerfacc OccupancyGrid._update)
    @torch.no_grad()
    def _sample_indices(self, step: int, warmup_steps: int) -> Optional[torch.Tensor]:
        if step < warmup_steps:
            return None  # all cells, in order
        n = self.num_cells // 4
        dev = self.occs.device
        uniform = torch.randint(self.num_cells, (n,), device=dev, generator=self.generator)
        occupied = torch.nonzero(self._binary.flatten())[:, 0]
        if n < occupied.numel():
            occupied = occupied[torch.randint(occupied.numel(), (n,), device=dev, generator=self.generator)]
        return torch.cat([uniform, occupied], dim=0)

    @torch.no_grad()
    def cell_points(self, indices: Optional[torch.Tensor], jitter: Optional[torch.Tensor]):
        dev = self.occs.device
        idx = indices if indices is not None else torch.arange(self.num_cells, device=dev)
        rx, ry, rz = self._res
        gx = torch.div(idx, ry * rz, rounding_mode="floor")
        gy = torch.div(idx, rz, rounding_mode="floor") % ry
        gz = idx % rz
        coords = torch.stack([gx, gy, gz], dim=-1).float()
        if jitter is None:
            jitter = torch.rand(idx.shape[0], 3, device=dev, generator=self.generator)
        x = (coords + jitter) / self.resolution.float()
        if self._contraction_type == ContractionType.UN_BOUNDED_SPHERE:
            keep = (x - 0.5).norm(dim=1) < 0.5
            x, idx = x[keep], idx[keep]
            indices = idx
        return indices, _contract_inv(x, self._roi_aabb, self._contraction_type)

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95,
                warmup_steps: int = 256, indices: Optional[torch.Tensor] = None, jitter: Optional[torch.Tensor] = None) -> None:
        if not self.occs.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        if indices is None:
            indices = self._sample_indices(step, warmup_steps)
        indices, pts = self.cell_points(indices, jitter)
        occ = occ_eval_fn(pts).reshape(-1)
        if self._workspace is None or self._workspace.device != self.occs.device:
            nbytes = L.load().ia_occ_workspace_bytes(self.num_cells)
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.occs.device)
        ops.occ_update(indices, occ, self.occs, ema_decay, occ_thre, self._binary.view(torch.uint8).reshape(-1),
                       self.bitfield, self._workspace)

    @torch.no_grad()
    def every_n_step(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                     warmup_steps: int = 256, n: int = 16, indices=None, jitter=None) -> None:
        if not self.training:
            raise RuntimeError("You should only call this function only during training. "
                               "Please call _update() directly if you want to update the field during inference.")
        if step % n == 0 and self.training:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps, indices, jitter)


@torch.no_grad()
def ray_aabb_intersect(rays_o: torch.Tensor, rays_d: torch.Tensor, aabb: torch.Tensor, clamp_zero: bool = True
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """nerfacc.intersection.ray_aabb_intersect (reference models/neus.py:153): miss => (1e10, 1e10)."""
    if not (rays_o.is_cuda and rays_d.is_cuda):
        raise NotImplementedError("Only support cuda inputs.")
    bb = aabb.tolist() if isinstance(aabb, torch.Tensor) else list(aabb)
    return ops.aabb_intersect(rays_o, rays_d, bb, clamp_zero)


def pack_info(ray_indices: torch.Tensor, n_rays: int) -> torch.Tensor:
    num = torch.zeros(n_rays, dtype=torch.int64, device=ray_indices.device)
    num.index_add_(0, ray_indices.long(), torch.ones_like(ray_indices, dtype=torch.int64))
    cum = torch.cumsum(num, 0)
    return torch.stack([cum - num, num], dim=1).int()


def unpack_info(packed_info: torch.Tensor) -> torch.Tensor:
    n = packed_info.shape[0]
    return torch.repeat_interleave(torch.arange(n, device=packed_info.device, dtype=torch.int32), packed_info[:, 1].long())


@torch.no_grad()
def render_visibility(alphas: torch.Tensor, *, ray_indices=None, packed_info=None, early_stop_eps: float = 1e-4,
                      alpha_thre: float = 0.0, n_rays: Optional[int] = None) -> torch.Tensor:
    if packed_info is None:
        packed_info = pack_info(ray_indices, n_rays)
    return ops.visibility(alphas, packed_info.contiguous(), early_stop_eps, alpha_thre)


@torch.no_grad()
def ray_marching(rays_o: torch.Tensor, rays_d: torch.Tensor, t_min: Optional[torch.Tensor] = None,
                 t_max: Optional[torch.Tensor] = None, scene_aabb: Optional[torch.Tensor] = None,
                 grid: Optional[OccupancyGrid] = None, sigma_fn: Optional[Callable] = None,
                 alpha_fn: Optional[Callable] = None, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 near_plane=None, far_plane=None, render_step_size: float = 1e-3, stratified: bool = False,
                 cone_angle: float = 0.0, stratified_u: Optional[torch.Tensor] = None, return_packed: bool = False,
                 scene_aabb_host=None):
    """nerfacc.ray_marching (reference models/neus.py:159-169, 209-220).

    Returns (ray_indices int32 [S], t_starts [S,1], t_ends [S,1]) (+ packed_info [R,2] when return_packed)."""
    if not rays_o.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    if alpha_fn is not None and sigma_fn is not None:
        raise ValueError("Only one of `alpha_fn` and `sigma_fn` should be provided.")
    t_min, t_max, gdesc, bitfield = march_inputs(rays_o, rays_d, t_min, t_max, scene_aabb, grid, near_plane, far_plane, render_step_size,
                                                 stratified, stratified_u, scene_aabb_host)
    packed_info, ray_indices, t_starts, t_ends = ops.march(rays_o, rays_d, t_min, t_max, gdesc, bitfield,
                                                            float(render_step_size), float(cone_angle))
    return march_finish(rays_o.shape[0], packed_info, ray_indices, t_starts, t_ends, sigma_fn, alpha_fn, early_stop_eps, alpha_thre,
                        return_packed)


def march_inputs(rays_o, rays_d, t_min, t_max, scene_aabb, grid, near_plane, far_plane, render_step_size, stratified, stratified_u,
                 scene_aabb_host=None):
    """The part of nerfacc.ray_marching in front of the marching kernel: ray / AABB test, near / far clamps, stratified
    offset, grid descriptor.  -> (t_min, t_max, grid desc, bitfield)."""
    n_rays = rays_o.shape[0]
    if t_min is None or t_max is None:
        if scene_aabb is not None:
            bb = scene_aabb_host if scene_aabb_host is not None else scene_aabb.tolist()
            t_min, t_max = ops.aabb_intersect(rays_o, rays_d, bb, True)
        else:
            t_min = torch.zeros(n_rays, device=rays_o.device)
            t_max = torch.full((n_rays,), 1e10, device=rays_o.device)
    if near_plane is not None:
        t_min = torch.clamp(t_min, min=near_plane)
    if far_plane is not None:
        t_max = torch.clamp(t_max, max=far_plane)
    if stratified:
        u = stratified_u if stratified_u is not None else torch.rand_like(t_min)
        t_min = t_min + u * render_step_size
    if grid is not None:
        gdesc, bitfield = grid.grid_desc, grid.bitfield
    else:
        gdesc, bitfield = ops.make_grid_desc([-1e10] * 3 + [1e10] * 3, [1, 1, 1], int(ContractionType.AABB)), None
    return t_min, t_max, gdesc, bitfield


def march_finish(n_rays, packed_info, ray_indices, t_starts, t_ends, sigma_fn, alpha_fn, early_stop_eps, alpha_thre, return_packed):
    """The part of nerfacc.ray_marching behind the marching kernel: visibility pruning through sigma_fn / alpha_fn."""
    t_starts, t_ends = t_starts[:, None], t_ends[:, None]
    if (alpha_thre > 0.0 or early_stop_eps > 0.0) and (sigma_fn is not None or alpha_fn is not None):
        if sigma_fn is not None and os.environ.get("IA_NO_FUSED_PRUNE") is None:
            # alphas, visibility, compaction and the new packed_info as count -> scan -> write (ops.prune_samples)
            sigmas = sigma_fn(t_starts, t_ends, ray_indices)
            ray_indices, t_starts, t_ends, packed_kept = ops.prune_samples(sigmas, t_starts, t_ends, packed_info, early_stop_eps, alpha_thre)
            if return_packed:
                return ray_indices, t_starts, t_ends, packed_kept
            return ray_indices, t_starts, t_ends
        if sigma_fn is not None:
            sigmas = sigma_fn(t_starts, t_ends, ray_indices)
            alphas = 1.0 - torch.exp(-sigmas * (t_ends - t_starts))
        else:
            alphas = alpha_fn(t_starts, t_ends, ray_indices)
        masks = ops.visibility(alphas, packed_info, early_stop_eps, alpha_thre)
        keep = torch.nonzero(masks).squeeze(1)          # one compaction (one host read-back) shared by the three arrays
        ray_indices, t_starts, t_ends = ray_indices[keep], t_starts[keep], t_ends[keep]
        if return_packed:
            packed_info = pack_info(ray_indices, n_rays)
    if return_packed:
        return ray_indices, t_starts, t_ends, packed_info
    return ray_indices, t_starts, t_ends


def _packed(packed_info, ray_indices, n_rays):
    if packed_info is None:
        if ray_indices is None or n_rays is None:
            raise ValueError("Either `packed_info` or (`ray_indices`, `n_rays`) must be given.")
        packed_info = pack_info(ray_indices, n_rays)
    return packed_info.contiguous()


def render_weight_from_alpha(alphas: torch.Tensor, *, packed_info=None, ray_indices=None, n_rays=None) -> torch.Tensor:
    """nerfacc.render_weight_from_alpha (reference models/neus.py:234): w_i = alpha_i prod_{j<i}(1-alpha_j)."""
    if not alphas.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    assert alphas.dim() == 2 and alphas.shape[-1] == 1, "alphas must be [n_samples, 1]"
    w = ops.composite_alpha(alphas.reshape(-1), _packed(packed_info, ray_indices, n_rays))[0]
    return w[:, None]


def render_weight_from_density(t_starts, t_ends, sigmas, *, packed_info=None, ray_indices=None, n_rays=None) -> torch.Tensor:
    """nerfacc.render_weight_from_density (reference models/neus.py:181)."""
    if not sigmas.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    w = ops.composite_density(sigmas.reshape(-1), t_starts.reshape(-1), t_ends.reshape(-1),
                              _packed(packed_info, ray_indices, n_rays))[0]
    return w[:, None]


def accumulate_along_rays(weights: torch.Tensor, ray_indices: torch.Tensor, values: Optional[torch.Tensor] = None,
                          n_rays: Optional[int] = None) -> torch.Tensor:
    """nerfacc.accumulate_along_rays (reference models/neus.py:182-184, 235-239).  The fused NeuS path does not
    come through here (ops.composite_* folds the reductions into the scan); this free function keeps the API for
    other callers and is a plain deterministic segmented sum expressed with index_add."""
    assert weights.dim() == 2 and weights.shape[-1] == 1
    if not weights.is_cuda:
        raise NotImplementedError("Only support cuda inputs.")
    src = weights if values is None else weights * values
    if n_rays is None:
        n_rays = int(ray_indices.max()) + 1 if ray_indices.numel() else 0
    out = torch.zeros(n_rays, src.shape[-1], device=src.device, dtype=src.dtype)
    return out.index_add_(0, ray_indices.long(), src)
