"""NeuS renderer with the reference's surface (models/neus.py:15-318) on the sm_100a kernels.

Orchestration follows reference `NeuSModel.forward_` / `forward_bg_` / `update_step` step for step; what
changes is the execution: marching reads a packed occupancy bitfield and returns packed_info, the
NeuS alpha + transmittance scan + four per-ray reductions are one kernel (ops.composite_neus), the
background density compositing another (ops.composite_density).  Random draws made inside the reference
(stratified jitter, curvature directions, occupancy jitter) can be injected for parity runs.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from . import registry as models
from .nerfacc_api import ContractionType, OccupancyGrid, ray_aabb_intersect, ray_marching
from .network_utils import update_module_step
from . import geometry as _geometry  # noqa: F401  (registers volume-sdf / volume-density)
from . import texture as _texture    # noqa: F401  (registers the colour heads)


class VarianceNetwork(nn.Module):
    """reference models/neus.py:15-43."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.init_val = self.config["init_val"]
        self.register_parameter("variance", nn.Parameter(torch.tensor(float(self.config["init_val"]))))
        self.modulate = self.config.get("modulate", False)
        self.do_mod = False
        if self.modulate:
            self.mod_start_steps = self.config["mod_start_steps"]
            self.reach_max_steps = self.config["reach_max_steps"]
            self.max_inv_s = self.config["max_inv_s"]

    @property
    def inv_s(self):
        val = torch.exp(self.variance * 10.0)
        if self.modulate and self.do_mod:
            val = val.clamp_max(self.mod_val)
        return val

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * self.inv_s

    def update_step(self, epoch, global_step):
        if self.modulate:
            self.do_mod = global_step > self.mod_start_steps
            if not self.do_mod:
                self.prev_inv_s = self.inv_s.item()
            else:
                self.mod_val = min((global_step / self.reach_max_steps) * (self.max_inv_s - self.prev_inv_s) + self.prev_inv_s,
                                   self.max_inv_s)


@models.register("neus")
class NeuSModel(nn.Module):
    """reference models/neus.py:46-318."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.setup()

    def setup(self):
        cfg = self.config
        self.geometry = models.make(cfg["geometry"]["name"], cfg["geometry"])
        self.texture = models.make(cfg["texture"]["name"], cfg["texture"])
        self.geometry.contraction_type = ContractionType.AABB
        self.learned_background = bool(cfg.get("learned_background", False))
        if self.learned_background:
            self.geometry_bg = models.make(cfg["geometry_bg"]["name"], cfg["geometry_bg"])
            self.texture_bg = models.make(cfg["texture_bg"]["name"], cfg["texture_bg"])
            self.geometry_bg.contraction_type = ContractionType.UN_BOUNDED_SPHERE
            self.near_plane_bg, self.far_plane_bg = 0.1, 1e3
            self.cone_angle_bg = 10 ** (math.log10(self.far_plane_bg) / cfg["num_samples_per_ray_bg"]) - 1.0
            self.render_step_size_bg = 0.01
        self.variance = VarianceNetwork(cfg["variance"])
        r = cfg["radius"]
        self._aabb_host = [-r, -r, -r, r, r, r]
        self.register_buffer("scene_aabb", torch.as_tensor(self._aabb_host, dtype=torch.float32))
        self.grid_prune = bool(cfg.get("grid_prune", True))
        if self.grid_prune:
            self.occupancy_grid = OccupancyGrid(roi_aabb=self.scene_aabb, resolution=128, contraction_type=ContractionType.AABB)
            if self.learned_background:
                self.occupancy_grid_bg = OccupancyGrid(roi_aabb=self.scene_aabb, resolution=256,
                                                       contraction_type=ContractionType.UN_BOUNDED_SPHERE)
        self.randomized = cfg.get("randomized", True)
        self.background_color = None
        self.render_step_size = 1.732 * 2 * r / cfg["num_samples_per_ray"]
        self.cos_anneal_ratio = 1.0

    # ---- reference models/neus.py:79-111 ----------------------------------------------------------------
    def occ_eval_fn(self, x):
        sdf = self.geometry(x, with_grad=False, with_feature=False)
        inv_s = self.variance.inv_s.reshape(1, 1).clip(1e-6, 1e6).expand(sdf.shape[0], 1)
        estimated_next_sdf = sdf[..., None] - self.render_step_size * 0.5
        estimated_prev_sdf = sdf[..., None] + self.render_step_size * 0.5
        prev_cdf = torch.sigmoid(estimated_prev_sdf * inv_s)
        next_cdf = torch.sigmoid(estimated_next_sdf * inv_s)
        return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).view(-1, 1).clip(0.0, 1.0)

    def occ_eval_fn_bg(self, x):
        return self.geometry_bg.density(x)[..., None] * self.render_step_size_bg

    def update_step(self, epoch, global_step, occ_inputs: Optional[dict] = None, update_occupancy: bool = True):
        update_module_step(self.geometry, epoch, global_step)
        update_module_step(self.texture, epoch, global_step)
        if self.learned_background:
            update_module_step(self.geometry_bg, epoch, global_step)
            update_module_step(self.texture_bg, epoch, global_step)
        update_module_step(self.variance, epoch, global_step)
        cos_anneal_end = self.config.get("cos_anneal_end", 0)
        self.cos_anneal_ratio = 1.0 if cos_anneal_end == 0 else min(1.0, global_step / cos_anneal_end)
        if self.training and self.grid_prune and update_occupancy:
            oi = occ_inputs or {}
            self.occupancy_grid.every_n_step(step=global_step, occ_eval_fn=self.occ_eval_fn,
                                             occ_thre=self.config.get("grid_prune_occ_thre", 0.01),
                                             indices=oi.get("indices"), jitter=oi.get("jitter"))
            if self.learned_background:
                self.occupancy_grid_bg.every_n_step(step=global_step, occ_eval_fn=self.occ_eval_fn_bg,
                                                    occ_thre=self.config.get("grid_prune_occ_thre_bg", 0.01),
                                                    indices=oi.get("indices_bg"), jitter=oi.get("jitter_bg"))

    # ---- reference models/neus.py:113-115, 308-318 -----------------------------------------------------------------
    def isosurface(self):
        return self.geometry.isosurface()

    @torch.no_grad()
    def export(self, export_config):
        mesh = self.isosurface()
        if export_config.get("export_vertex_color", False) and mesh["v_pos"].shape[0] > 0:
            device = self.scene_aabb.device
            chunk = int(export_config.get("chunk_size", 2097152))
            was_training = self.geometry.training
            self.geometry.eval()
            rgb, nrm = [], []
            for i in range(0, mesh["v_pos"].shape[0], chunk):
                pts = mesh["v_pos"][i:i + chunk].to(device).float().contiguous()
                _, sdf_grad, features = self.geometry(pts, with_grad=True, with_feature=True)
                nrm.append(F.normalize(sdf_grad, p=2, dim=-1).cpu())
                rgb.append(torch.sigmoid(features[..., 1:4]).cpu())
            self.geometry.train(was_training)
            mesh["v_rgb"], mesh["v_norm"] = torch.cat(rgb, dim=0), torch.cat(nrm, dim=0)
        return mesh

    # ---- reference models/neus.py:117-139 (stand-alone; forward_ uses the fused kernel) ---------------
    def get_alpha(self, sdf, normal, dirs, dists):
        inv_s = self.variance.inv_s.reshape(1).clip(1e-6, 1e6)
        packed = torch.stack([torch.arange(sdf.shape[0], device=sdf.device, dtype=torch.int32),
                              torch.ones(sdf.shape[0], device=sdf.device, dtype=torch.int32)], dim=1).contiguous()
        # one sample per pseudo-ray: weights == alpha, and gradients flow through the same kernel
        w = ops.composite_neus(sdf, normal, dirs, dists.reshape(-1), inv_s, self.cos_anneal_ratio, packed)[0]
        return w

    # ---- reference models/neus.py:141-203 --------------------------------------------------------------
    def march_bg_(self, rays_o, rays_d, stratified_u: Optional[torch.Tensor] = None):
        """The marching half of reference forward_bg_ (models/neus.py:141-169), including the early-stop pruning
        through sigma_fn (a no-grad background density evaluation).  It depends on the rays, the background grid and
        the background networks only, which lets forward_ issue it next to the foreground march: all host read-backs
        of sample counts then sit at the start of the step instead of draining the GPU queue in the middle of it."""

        def sigma_fn(t_starts, t_ends, ray_indices):
            # rays_o[ri] + rays_d[ri] * (t_starts + t_ends) / 2.0 : the same roundings (a product halved == the product of the half)
            positions = ops.ray_samples(rays_o, rays_d, ray_indices, t_starts, t_ends, False, False, False)[0]
            return self.geometry_bg.density(positions)[..., None]

        _, t_max = ray_aabb_intersect(rays_o, rays_d, self._aabb_host)
        near_plane = torch.where(t_max > 1e9, self.near_plane_bg, t_max)
        with torch.no_grad():
            return ray_marching(
                rays_o, rays_d, scene_aabb=None, grid=self.occupancy_grid_bg if self.grid_prune else None,
                sigma_fn=sigma_fn, near_plane=near_plane, far_plane=self.far_plane_bg,
                render_step_size=self.render_step_size_bg, stratified=self.randomized, cone_angle=self.cone_angle_bg,
                alpha_thre=0.0, stratified_u=stratified_u, return_packed=True)

    @torch.no_grad()
    def march_both_(self, rays_o, rays_d, stratified_u=None, stratified_u_bg=None):
        """The foreground march of forward_ (models/neus.py:209-220) and the background march of forward_bg_ (:153-169,
        march_bg_) with the marching kernels of the two launched as a pair: same inputs, same per-ray function, same
        outputs as ray_marching(...) twice."""
        from .nerfacc_api import march_finish, march_inputs
        fg = march_inputs(rays_o, rays_d, None, None, self.scene_aabb, self.occupancy_grid if self.grid_prune else None, None, None,
                          self.render_step_size, self.randomized, stratified_u, self._aabb_host)
        _, t_max = ray_aabb_intersect(rays_o, rays_d, self._aabb_host)
        near_plane = torch.where(t_max > 1e9, self.near_plane_bg, t_max)
        bg = march_inputs(rays_o, rays_d, None, None, None, self.occupancy_grid_bg if self.grid_prune else None, near_plane,
                          self.far_plane_bg, self.render_step_size_bg, self.randomized, stratified_u_bg)
        out_fg, out_bg = ops.march_pair(rays_o, rays_d, (*fg, self.render_step_size, 0.0), (*bg, self.render_step_size_bg, self.cone_angle_bg))

        def sigma_fn(t_starts, t_ends, ray_indices):
            positions = ops.ray_samples(rays_o, rays_d, ray_indices, t_starts, t_ends, False, False, False)[0]
            return self.geometry_bg.density(positions)[..., None]

        n = rays_o.shape[0]
        marched_fg = march_finish(n, *out_fg, None, None, 1e-4, 0.0, True)
        marched_bg = march_finish(n, *out_bg, sigma_fn, None, 1e-4, 0.0, True)
        return marched_fg, marched_bg

    def forward_bg_(self, rays, stratified_u: Optional[torch.Tensor] = None, marched=None, mix_later: bool = False):
        """mix_later (forward_ only): leave `comp_rgb` without the background colour and `rays_valid` unset -- forward_ applies
        both together with the foreground / background mix in one kernel (ops.ray_mix)."""
        n_rays = rays.shape[0]
        rays_o, rays_d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
        if marched is None:
            marched = self.march_bg_(rays_o, rays_d, stratified_u)
        ray_indices, t_starts, t_ends, packed_info = marched
        ri = ray_indices.long()
        positions, t_dirs, midpoints, intervals = ops.ray_samples(rays_o, rays_d, ray_indices, t_starts, t_ends)
        density, feature = self.geometry_bg(positions)
        rgb = self.texture_bg(feature, t_dirs)
        weights, opacity, depth, comp_rgb, _, _ = ops.composite_density(
            density, t_starts.reshape(-1), t_ends.reshape(-1), packed_info, t_mid=midpoints.reshape(-1), rgb=rgb)
        opacity, depth = opacity[:, None], depth[:, None]
        if not mix_later:
            comp_rgb = comp_rgb + self.background_color * (1.0 - opacity)
        out = {"comp_rgb": comp_rgb, "opacity": opacity, "depth": depth, "rays_valid": None if mix_later else opacity > 0,
               "num_samples": torch.full((1,), len(t_starts), dtype=torch.int32, device=rays.device)}
        if self.training:
            out.update({"weights": weights.view(-1), "points": midpoints.view(-1), "intervals": intervals.view(-1),
                        "ray_indices": ri.view(-1)})
        return out

    # ---- reference models/neus.py:205-283 --------------------------------------------------------------
    def forward_(self, rays, stratified_u: Optional[torch.Tensor] = None, rand_directions: Optional[torch.Tensor] = None,
                 stratified_u_bg: Optional[torch.Tensor] = None):
        n_rays = rays.shape[0]
        rays_o, rays_d = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
        if self.learned_background and rays.is_cuda and os.environ.get("IA_NO_MARCH_PAIR") is None:
            # both marches in one count launch, one read-back of both totals and one write launch (ops.march_pair): each of
            # them alone is one thread per ray -- 128 small CTAs on 148 SMs, latency bound
            (ray_indices, t_starts, t_ends, packed_info), marched_bg = self.march_both_(rays_o, rays_d, stratified_u, stratified_u_bg)
        else:
            with torch.no_grad():
                ray_indices, t_starts, t_ends, packed_info = ray_marching(
                    rays_o, rays_d, scene_aabb=self.scene_aabb, scene_aabb_host=self._aabb_host,
                    grid=self.occupancy_grid if self.grid_prune else None, alpha_fn=None, near_plane=None, far_plane=None,
                    render_step_size=self.render_step_size, stratified=self.randomized, cone_angle=0.0, alpha_thre=0.0,
                    stratified_u=stratified_u, return_packed=True)
            # background marching is independent of the foreground networks: issue it now so that its sample-count
            # read-back does not drain the GPU queue in the middle of the step
            marched_bg = self.march_bg_(rays_o, rays_d, stratified_u_bg) if self.learned_background else None
        ri = ray_indices.long()
        positions, t_dirs, midpoints, dists = ops.ray_samples(rays_o, rays_d, ray_indices, t_starts, t_ends)
        fused_head = getattr(self.geometry, "supports_fused_head", lambda: False)() and \
            getattr(self.texture, "supports_fused_head", lambda g: False)(self.geometry)
        if fused_head:
            # same arithmetic as the generic branch below; the 65-wide `feature` and the colour head's 87-wide input row
            # are assembled in one buffer (ops.sdf_head) instead of out -> cat -> cat
            h, pts01, sdf_grad, sdf_laplace, w_last, b_last = self.geometry.forward_hidden(positions, rand_directions)
            normal = ops.normalize3(sdf_grad)
            rgb, sdf = self.texture.forward_fused_head(h, w_last, b_last, pts01, t_dirs, normal)
            if not self.training:
                sdf, sdf_grad, sdf_laplace, rgb, normal = (v.detach() for v in (sdf, sdf_grad, sdf_laplace, rgb, normal))
        else:
            sdf, sdf_grad, feature, sdf_laplace = self.geometry(positions, with_grad=True, with_feature=True, with_laplace=True,
                                                                rand_directions=rand_directions)
            normal = ops.normalize3(sdf_grad)
            rgb = self.texture(feature, t_dirs, normal)
        inv_s = self.variance.inv_s.reshape(1).clip(1e-6, 1e6)
        weights, opacity, depth, comp_rgb, comp_normal, alpha = ops.composite_neus(
            sdf, normal, t_dirs, dists.reshape(-1), inv_s, self.cos_anneal_ratio, packed_info,
            t_mid=midpoints.reshape(-1), rgb=rgb, nrm=normal)
        opacity, depth = opacity[:, None], depth[:, None]
        fused_mix = self.learned_background and rays.is_cuda and os.environ.get("IA_NO_RAY_MIX") is None
        rays_fg = opacity > 0.1
        comp_normal = ops.normalize3(comp_normal)
        comp_normal = comp_normal * rays_fg.float()      # Appendix C-9
        out = {"comp_rgb": comp_rgb, "comp_normal": comp_normal, "opacity": opacity, "depth": depth,
               "rays_valid": None if fused_mix else opacity > 0,
               "num_samples": torch.full((1,), len(t_starts), dtype=torch.int32, device=rays.device)}
        if self.training:
            out.update({"sdf_samples": sdf, "sdf_grad_samples": sdf_grad, "weights": weights.view(-1),
                        "points": midpoints.view(-1), "intervals": dists.view(-1), "ray_indices": ri.view(-1),
                        "sdf_laplace_samples": sdf_laplace})
        if fused_mix:
            # comp_rgb_bg + background colour, the foreground / background mix and the three validity masks in one kernel
            out_bg = self.forward_bg_(rays, stratified_u=stratified_u_bg, marched=marched_bg, mix_later=True)
            bgc = self.background_color.to(device=rays.device, dtype=torch.float32).reshape(3)
            rgb_bg, rgb_full, valid, valid_bg, valid_full = ops.ray_mix(comp_rgb, opacity, out_bg["comp_rgb"], out_bg["opacity"], bgc)
            out["rays_valid"], out_bg["comp_rgb"], out_bg["rays_valid"] = valid, rgb_bg, valid_bg
        elif self.learned_background:
            out_bg = self.forward_bg_(rays, stratified_u=stratified_u_bg, marched=marched_bg)
        else:
            out_bg = {"comp_rgb": self.background_color[None, :].expand(*comp_rgb.shape),
                      "num_samples": torch.zeros_like(out["num_samples"]),
                      "rays_valid": torch.zeros_like(out["rays_valid"])}
        # host-side copies of the marched sample counts (already read back by the marcher): systems.NeuSSystem's
        # dynamic ray sampling (reference systems/neus.py:125-128) uses them instead of num_samples_full.item()
        self.last_num_samples = len(t_starts)
        self.last_num_samples_full = len(t_starts) + (len(marched_bg[1]) if marched_bg is not None else 0)
        if fused_mix:
            out_full = {"comp_rgb": rgb_full, "num_samples": out["num_samples"] + out_bg["num_samples"], "rays_valid": valid_full}
        else:
            out_full = {"comp_rgb": out["comp_rgb"] + out_bg["comp_rgb"] * (1.0 - out["opacity"]),
                        "num_samples": out["num_samples"] + out_bg["num_samples"],
                        "rays_valid": out["rays_valid"] | out_bg["rays_valid"]}
        return {**out, **{k + "_bg": v for k, v in out_bg.items()}, **{k + "_full": v for k, v in out_full.items()}}

    def forward(self, rays, **rng):
        if self.training:
            out = self.forward_(rays, **rng)
        else:
            chunk = int(self.config.get("ray_chunk", 2048))
            parts = [self.forward_(rays[i:i + chunk], **rng) for i in range(0, rays.shape[0], chunk)]
            out = {k: (torch.cat([p[k] for p in parts], dim=0)) for k in parts[0]}
        return {**out, "inv_s": self.variance.inv_s}

    def train(self, mode=True):
        self.randomized = mode and self.config.get("randomized", True)
        return super().train(mode=mode)

    def eval(self):
        self.randomized = False
        return super().eval()

    def regularizations(self, out):
        losses = {}
        losses.update(self.geometry.regularizations(out))
        losses.update(self.texture.regularizations(out))
        return losses
