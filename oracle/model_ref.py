"""ORACLE (test infrastructure only) -- CPU/PyTorch restatement of the Instant-angelo training hot
path that lives in the reference's own Python files:

  models/network_utils.py:40-140   ProgressiveBandHashGrid, CompositeEncoding, VanillaMLP
  models/geometry.py:19-31,152-314 contract_to_unisphere, VolumeDensity, VolumeSDF (FD gradient, curvature)
  models/neus.py:15-43,79-293      VarianceNetwork, occupancy refresh, get_alpha, forward_/forward_bg_
  models/texture.py:10-64,113-149  VolumeRadiance, VolumeDualColor, VolumeDualColorV3
  models/utils.py:54-114           trunc_exp, get_activation, scale_anything
  systems/neus.py:130-194          loss terms;  systems/base.py:28-45  C() schedules

The third-party arithmetic (tcnn hash grid / SH, nerfacc marching / compositing) comes from
oracle/tcnn_ref.py and oracle/nerfacc_ref.py.  This file is pinned against the reference's own
Python by tests/golden/make_golden.py, which imports /root/reference/models with the missing
third-party packages stubbed by those two oracle modules and stores the outputs as fixtures.

Every random draw of the reference (randn_like at geometry.py:238, stratified jitter, occupancy
jitter) is an explicit argument here.  Configs are plain nested dicts.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nerfacc_ref as nf
from . import tcnn_ref as tc


# ---------------------------------------------------------------------------------------------
# helpers (models/utils.py)
# ---------------------------------------------------------------------------------------------

def _f32(x):
    """`.float()` of the reference (tcnn returns fp16, the reference casts back: models/geometry.py:207, texture.py:28) --
    except for float64 tensors, which stay float64: running the whole oracle on `.double()` parameters and inputs gives the
    high-precision arbiter the parity tests use to measure the fp32 oracle's own rounding noise."""
    return x if x.dtype == torch.float64 else x.float()


def scale_anything(dat, inp_scale, tgt_scale):
    dat = (dat - inp_scale[0]) / (inp_scale[1] - inp_scale[0])
    return dat * (tgt_scale[1] - tgt_scale[0]) + tgt_scale[0]


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


def get_activation(name):
    if name is None or str(name).lower() == "none":
        return lambda x: x
    name = str(name).lower()
    if name == "sigmoid":
        return torch.sigmoid
    if name == "trunc_exp":
        return _TruncExp.apply
    if name == "tanh":
        return torch.tanh
    if name.startswith("scale"):
        s = float(name[5:])
        return lambda x: x.clamp(0.0, s) / s
    if name.startswith("clamp"):
        s = float(name[5:])
        return lambda x: x.clamp(0.0, s)
    if name.startswith("mul"):
        s = float(name[3:])
        return lambda x: x * s
    if name[0] in "+-":
        s = float(name)
        return lambda x: x + s
    return getattr(F, name)


def schedule_value(value, global_step: int, current_epoch: int = 0) -> float:
    """systems/base.py:28-45  C()."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    start_step, start_value, end_value, end_step = value
    cur = global_step if isinstance(end_step, int) else current_epoch
    return start_value + (end_value - start_value) * max(min(1.0, (cur - start_step) / (end_step - start_step)), 0.0)


# ---------------------------------------------------------------------------------------------
# encodings (models/network_utils.py:40-93)
# ---------------------------------------------------------------------------------------------

class RefTcnnEncoding(nn.Module):
    """tcnn.Encoding stand-in: HashGrid or SphericalHarmonics."""

    def __init__(self, n_input_dims: int, cfg: dict, seed: int = 1337):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.otype = cfg["otype"]
        if self.otype == "HashGrid":
            self.plan = tc.grid_plan(cfg["n_levels"], cfg["n_features_per_level"], cfg["log2_hashmap_size"],
                                     cfg["base_resolution"], cfg["per_level_scale"])
            g = torch.Generator().manual_seed(seed)
            self.params = nn.Parameter((torch.rand(self.plan.n_params, generator=g) * 2 - 1) * 1e-4)
            self.n_output_dims = self.plan.n_output_dims
        elif self.otype == "SphericalHarmonics":
            self.degree = int(cfg["degree"])
            self.params = nn.Parameter(torch.zeros(0))
            self.n_output_dims = self.degree ** 2
        else:
            raise NotImplementedError(self.otype)

    def forward(self, x, active_levels=None):
        if self.otype == "HashGrid":
            return tc.hashgrid_forward(x, self.params, self.plan, active_levels)
        return tc.sh_forward(x, self.degree)


class RefProgressiveBandHashGrid(nn.Module):
    def __init__(self, in_channels: int, cfg: dict):
        super().__init__()
        self.n_input_dims = in_channels
        ecfg = dict(cfg)
        ecfg["otype"] = "HashGrid"
        self.encoding = RefTcnnEncoding(in_channels, ecfg)
        self.n_output_dims = self.encoding.n_output_dims
        self.n_level, self.n_features_per_level = cfg["n_levels"], cfg["n_features_per_level"]
        self.start_level, self.start_step, self.update_steps = cfg["start_level"], cfg["start_step"], cfg["update_steps"]
        self.current_level = self.start_level
        self.mask = torch.zeros(self.n_level * self.n_features_per_level)
        self.mask[: self.current_level * self.n_features_per_level] = 1.0

    def forward(self, x):
        return self.encoding(x) * self.mask

    def update_step(self, epoch, global_step):
        self.current_level = min(self.start_level + max(global_step - self.start_step, 0) // self.update_steps, self.n_level)
        self.mask[: self.current_level * self.n_features_per_level] = 1.0


class RefCompositeEncoding(nn.Module):
    def __init__(self, encoding, include_xyz=False, xyz_scale=1.0, xyz_offset=0.0):
        super().__init__()
        self.encoding = encoding
        self.include_xyz, self.xyz_scale, self.xyz_offset = include_xyz, xyz_scale, xyz_offset
        self.n_output_dims = int(include_xyz) * encoding.n_input_dims + encoding.n_output_dims

    def forward(self, x):
        e = self.encoding(x)
        return e if not self.include_xyz else torch.cat([x * self.xyz_scale + self.xyz_offset, e], dim=-1)

    def update_step(self, epoch, global_step):
        if hasattr(self.encoding, "update_step"):
            self.encoding.update_step(epoch, global_step)


def ref_get_encoding(n_input_dims: int, cfg: dict) -> RefCompositeEncoding:
    if cfg["otype"] == "ProgressiveBandHashGrid":
        enc = RefProgressiveBandHashGrid(n_input_dims, cfg)
    else:
        enc = RefTcnnEncoding(n_input_dims, cfg)
    return RefCompositeEncoding(enc, include_xyz=cfg.get("include_xyz", False), xyz_scale=2.0, xyz_offset=-1.0)


# ---------------------------------------------------------------------------------------------
# VanillaMLP (models/network_utils.py:96-140)
# ---------------------------------------------------------------------------------------------

class RefWNLinear(nn.Module):
    """nn.Linear, optionally under torch weight_norm (dim=0): W = g * v / ||v||_row.  Parameter names
    follow the reference checkpoint keys (weight_g / weight_v / bias, or weight / bias)."""

    def __init__(self, dim_in, dim_out, weight_norm: bool):
        super().__init__()
        self.weight_norm = weight_norm
        w = torch.empty(dim_out, dim_in)
        self.bias = nn.Parameter(torch.zeros(dim_out))
        if weight_norm:
            self.weight_v = nn.Parameter(w)
            self.weight_g = nn.Parameter(torch.ones(dim_out, 1))
        else:
            self.weight = nn.Parameter(w)

    def raw_weight(self):
        return self.weight_v if self.weight_norm else self.weight

    def finish_init(self):
        if self.weight_norm:
            with torch.no_grad():
                self.weight_g.copy_(self.weight_v.norm(dim=1, keepdim=True))

    def effective_weight(self):
        if not self.weight_norm:
            return self.weight
        return self.weight_v * (self.weight_g / self.weight_v.norm(dim=1, keepdim=True))

    def forward(self, x):
        return F.linear(x, self.effective_weight(), self.bias)


class RefVanillaMLP(nn.Module):
    def __init__(self, dim_in, dim_out, cfg: dict):
        super().__init__()
        self.n_neurons, self.n_hidden_layers = cfg["n_neurons"], cfg["n_hidden_layers"]
        self.sphere_init, self.weight_norm = cfg.get("sphere_init", False), cfg.get("weight_norm", False)
        self.sphere_init_radius = cfg.get("sphere_init_radius", 0.5)
        dims = [dim_in] + [self.n_neurons] * self.n_hidden_layers + [dim_out]
        mods = []
        for i in range(len(dims) - 1):
            lin = self._make_linear(dims[i], dims[i + 1], is_first=(i == 0), is_last=(i == len(dims) - 2))
            mods.append(lin)
            if i < len(dims) - 2:
                mods.append(nn.Softplus(beta=100) if self.sphere_init else nn.ReLU())
        self.layers = nn.Sequential(*mods)
        self.output_activation = get_activation(cfg.get("output_activation", None))

    def _make_linear(self, dim_in, dim_out, is_first, is_last):
        layer = RefWNLinear(dim_in, dim_out, self.weight_norm)
        w = layer.raw_weight()
        with torch.no_grad():
            if self.sphere_init:
                if is_last:
                    layer.bias.fill_(-self.sphere_init_radius)
                    nn.init.normal_(w, mean=math.sqrt(math.pi) / math.sqrt(dim_in), std=0.0001)
                elif is_first:
                    w.zero_()
                    nn.init.normal_(w[:, :3], 0.0, math.sqrt(2) / math.sqrt(dim_out))
                else:
                    nn.init.normal_(w, 0.0, math.sqrt(2) / math.sqrt(dim_out))
            else:
                nn.init.kaiming_uniform_(w, nonlinearity="relu")
        layer.finish_init()
        return layer

    def forward(self, x):
        return self.output_activation(self.layers(_f32(x)))


class RefEncodingWithNetwork(nn.Module):
    def __init__(self, encoding, network):
        super().__init__()
        self.encoding, self.network = encoding, network

    def forward(self, x):
        return self.network(self.encoding(x))

    def update_step(self, epoch, global_step):
        self.encoding.update_step(epoch, global_step)


# ---------------------------------------------------------------------------------------------
# geometry (models/geometry.py)
# ---------------------------------------------------------------------------------------------

def contract_to_unisphere(x, radius, contraction_type):
    x = scale_anything(x, (-radius, radius), (0, 1))
    if contraction_type == nf.ContractionType.UN_BOUNDED_SPHERE:
        x = x * 2 - 1
        mag = x.norm(dim=-1, keepdim=True)
        x = torch.where(mag > 1, (2 - 1 / mag) * (x / mag), x)
        x = x / 4 + 0.5
    return x


_FD_SIGNS = torch.tensor([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0], [0, 0, 1.0], [0, 0, -1.0]])


class RefVolumeSDF(nn.Module):
    """models/geometry.py:180-314."""

    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        self.radius = cfg["radius"]
        self.contraction_type = nf.ContractionType.AABB
        self.n_output_dims = cfg["feature_dim"]
        self.encoding = ref_get_encoding(3, cfg["xyz_encoding_config"])
        self.network = RefVanillaMLP(self.encoding.n_output_dims, self.n_output_dims, cfg["mlp_network_config"])
        self.grad_type = cfg["grad_type"]
        self.finite_difference_eps = cfg.get("finite_difference_eps", 1e-3)
        self._finite_difference_eps = None

    def _sdf_net(self, pts01):
        return self.network(self.encoding(pts01.reshape(-1, 3)))

    def _fd_gradient(self, world_pts, eps):
        """6-tap central differences, geometry.py:219-234 (taps clamped in world space)."""
        taps = (world_pts[..., None, :] + _FD_SIGNS * eps).clamp(-self.radius, self.radius)
        taps01 = scale_anything(taps, (-self.radius, self.radius), (0, 1))
        s = _f32(self._sdf_net(taps01)[..., 0].view(*world_pts.shape[:-1], 6))
        return 0.5 * (s[..., 0::2] - s[..., 1::2]) / eps

    def forward(self, points, with_grad=True, with_feature=True, with_laplace=False, rand_directions=None):
        points_world = points
        if with_grad and self.grad_type == "analytic":
            points_world = points_world.requires_grad_(True) if points_world.is_leaf else points_world
        pts01 = contract_to_unisphere(points_world, self.radius, self.contraction_type)
        out = _f32(self._sdf_net(pts01).view(*pts01.shape[:-1], self.n_output_dims))
        sdf = out[..., 0]
        feature = torch.cat([out, pts01 * 2 - 1], dim=-1)
        grad = None
        if with_grad:
            if self.grad_type == "analytic":
                grad = torch.autograd.grad(sdf, points_world, grad_outputs=torch.ones_like(sdf), create_graph=True,
                                           retain_graph=True, only_inputs=True)[0]
            else:
                grad = self._fd_gradient(points_world, self._finite_difference_eps)
        laplace = None
        if with_laplace:
            eps = self._finite_difference_eps
            if rand_directions is None:
                rand_directions = torch.randn_like(pts01)
            rnd = F.normalize(rand_directions, dim=-1)
            normals = F.normalize(grad, dim=-1)
            tangent = torch.cross(normals, rnd, dim=-1)
            # Appendix C-1: the shift is applied to the NORMALISED coordinates and the result is then
            # treated as a world-space point (geometry.py:246, 264-265).
            shifted = pts01 + tangent * eps
            g_shift = self._fd_gradient(shifted, eps)
            n_shift = F.normalize(g_shift, dim=-1)
            dot = (normals * n_shift).sum(dim=-1, keepdim=True)
            laplace = torch.acos(torch.clamp(dot, -1.0 + 1e-6, 1.0 - 1e-6)) / math.pi
        rv = [sdf]
        if with_grad:
            rv.append(grad)
        if with_feature:
            rv.append(feature)
        if with_laplace:
            rv.append(laplace)
        return rv[0] if len(rv) == 1 else rv

    def forward_level(self, points):
        pts01 = contract_to_unisphere(points, self.radius, self.contraction_type)
        return self._sdf_net(pts01).view(*pts01.shape[:-1], self.n_output_dims)[..., 0]

    def update_step(self, epoch, global_step):
        self.encoding.update_step(epoch, global_step)
        if isinstance(self.finite_difference_eps, float):
            self._finite_difference_eps = self.finite_difference_eps
        elif self.finite_difference_eps == "progressive":
            hg = self.cfg["xyz_encoding_config"]
            level = min(hg["start_level"] + max(global_step - hg["start_step"], 0) // hg["update_steps"], hg["n_levels"])
            grid_res = hg["base_resolution"] * hg["per_level_scale"] ** (level - 1)
            self._finite_difference_eps = 2 * self.radius / grid_res
        else:
            raise ValueError(self.finite_difference_eps)


class RefVolumeDensity(nn.Module):
    """models/geometry.py:152-177 (background NeRF++ density field)."""

    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        self.radius = cfg["radius"]
        self.contraction_type = nf.ContractionType.UN_BOUNDED_SPHERE
        self.n_output_dims = cfg["feature_dim"]
        enc = ref_get_encoding(3, cfg["xyz_encoding_config"])
        net = RefVanillaMLP(enc.n_output_dims, self.n_output_dims, cfg["mlp_network_config"])
        self.encoding_with_network = RefEncodingWithNetwork(enc, net)

    def forward(self, points):
        pts = contract_to_unisphere(points, self.radius, self.contraction_type)
        out = _f32(self.encoding_with_network(pts.view(-1, 3)).view(*pts.shape[:-1], self.n_output_dims))
        density, feature = out[..., 0], out
        if "density_activation" in self.cfg:
            density = get_activation(self.cfg["density_activation"])(density + float(self.cfg["density_bias"]))
        if "feature_activation" in self.cfg:
            feature = get_activation(self.cfg["feature_activation"])(feature)
        return density, feature

    def update_step(self, epoch, global_step):
        self.encoding_with_network.update_step(epoch, global_step)


# ---------------------------------------------------------------------------------------------
# colour heads (models/texture.py)
# ---------------------------------------------------------------------------------------------

class RefVolumeRadiance(nn.Module):
    dual = False

    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        self.encoding = ref_get_encoding(3, cfg["dir_encoding_config"])
        self.n_input_dims = cfg["input_feature_dim"] + self.encoding.n_output_dims
        self.network = RefVanillaMLP(self.n_input_dims, 3, cfg["mlp_network_config"])

    def forward(self, features, dirs, *args):
        emb = self.encoding(((dirs + 1.0) / 2.0).view(-1, 3))
        inp = torch.cat([features.view(-1, features.shape[-1]), emb] + [a.view(-1, a.shape[-1]) for a in args], dim=-1)
        color = _f32(self.network(inp).view(*features.shape[:-1], 3))
        if "color_activation" in self.cfg:
            act = get_activation(self.cfg["color_activation"])
            color = act(color) + act(features[..., 1:4]) if self.dual else act(color)
        return color

    def update_step(self, epoch, global_step):
        pass


class RefVolumeDualColor(RefVolumeRadiance):
    """texture.py:38-64: sigmoid(mlp) + sigmoid(feature[1:4])."""
    dual = True


class RefVolumeDualColorV3(nn.Module):
    """texture.py:113-149 (UniSDF: camera net + reflected-direction net blended by a weight net)."""

    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        self.encoding = ref_get_encoding(3, cfg["dir_encoding_config"])
        self.n_input_dims = cfg["input_feature_dim"] + self.encoding.n_output_dims
        self.cam_network = RefVanillaMLP(self.n_input_dims, 3, cfg["mlp_network_config"])
        self.ref_network = RefVanillaMLP(self.n_input_dims, 3, cfg["mlp_network_config"])
        self.weight_network = RefVanillaMLP(cfg["input_feature_dim"], 1, cfg["weitht_network_config"])

    def forward(self, features, viewdirs, normals):
        emb = self.encoding(((viewdirs + 1.0) / 2.0).view(-1, 3))
        vdn = (-viewdirs * normals).sum(-1, keepdim=True)
        refdirs = 2 * vdn * normals + viewdirs
        remb = self.encoding(((refdirs + 1.0) / 2.0).view(-1, 3))
        inp = torch.cat([features.view(-1, features.shape[-1]), normals.view(-1, 3)], dim=-1)
        w = self.weight_network(inp)
        cam = _f32(self.cam_network(torch.cat([inp, emb], dim=-1)).view(*features.shape[:-1], 3))
        ref = _f32(self.ref_network(torch.cat([inp, remb], dim=-1)).view(*features.shape[:-1], 3))
        act = get_activation(self.cfg["color_activation"])
        return w * act(ref) + (1 - w) * act(cam)

    def update_step(self, epoch, global_step):
        pass


_TEXTURES = {"volume-radiance": RefVolumeRadiance, "volume-dual-color": RefVolumeDualColor,
             "volume-dual-colorV3": RefVolumeDualColorV3}


# ---------------------------------------------------------------------------------------------
# NeuS renderer (models/neus.py)
# ---------------------------------------------------------------------------------------------

class RefVarianceNetwork(nn.Module):
    def __init__(self, cfg: dict):
        super().__init__()
        self.variance = nn.Parameter(torch.tensor(float(cfg["init_val"])))
        self.modulate = cfg.get("modulate", False)
        self.cfg = cfg
        self.do_mod = False

    @property
    def inv_s(self):
        val = torch.exp(self.variance * 10.0)
        if self.modulate and self.do_mod:
            val = val.clamp_max(self.mod_val)
        return val

    def update_step(self, epoch, global_step):
        if self.modulate:
            self.do_mod = global_step > self.cfg["mod_start_steps"]
            if not self.do_mod:
                self.prev_inv_s = self.inv_s.item()
            else:
                self.mod_val = min((global_step / self.cfg["reach_max_steps"]) * (self.cfg["max_inv_s"] - self.prev_inv_s)
                                   + self.prev_inv_s, self.cfg["max_inv_s"])


class RefNeuSModel(nn.Module):
    def __init__(self, cfg: dict):
        super().__init__()
        self.cfg = cfg
        self.geometry = RefVolumeSDF(cfg["geometry"])
        self.texture = _TEXTURES[cfg["texture"]["name"]](cfg["texture"])
        self.learned_background = cfg.get("learned_background", False)
        r = cfg["radius"]
        if self.learned_background:
            self.geometry_bg = RefVolumeDensity(cfg["geometry_bg"])
            self.texture_bg = _TEXTURES[cfg["texture_bg"]["name"]](cfg["texture_bg"])
            self.near_plane_bg, self.far_plane_bg = 0.1, 1e3
            self.cone_angle_bg = 10 ** (math.log10(self.far_plane_bg) / cfg["num_samples_per_ray_bg"]) - 1.0
            self.render_step_size_bg = 0.01
        self.variance = RefVarianceNetwork(cfg["variance"])
        self.scene_aabb = torch.tensor([-r, -r, -r, r, r, r], dtype=torch.float32)
        self.grid_prune = cfg.get("grid_prune", True)
        if self.grid_prune:
            self.occupancy_grid = nf.OccupancyGrid(self.scene_aabb, 128, nf.ContractionType.AABB)
            if self.learned_background:
                self.occupancy_grid_bg = nf.OccupancyGrid(self.scene_aabb, 256, nf.ContractionType.UN_BOUNDED_SPHERE)
        self.randomized = cfg.get("randomized", True)
        self.background_color = None
        self.render_step_size = 1.732 * 2 * r / cfg["num_samples_per_ray"]
        self.cos_anneal_ratio = 1.0

    # -- neus.py:79-111 ------------------------------------------------------------------------
    def occ_eval_fn(self, x):
        sdf = self.geometry(x, with_grad=False, with_feature=False)
        inv_s = self.variance.inv_s.reshape(1, 1).clip(1e-6, 1e6).expand(sdf.shape[0], 1)
        nxt = sdf[..., None] - self.render_step_size * 0.5
        prv = sdf[..., None] + self.render_step_size * 0.5
        prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).view(-1, 1).clip(0.0, 1.0)

    def occ_eval_fn_bg(self, x):
        density, _ = self.geometry_bg(x)
        return density[..., None] * self.render_step_size_bg

    def update_step(self, epoch, global_step, occ_inputs: Optional[dict] = None, update_occupancy: bool = True):
        self.geometry.update_step(epoch, global_step)
        if self.learned_background:
            self.geometry_bg.update_step(epoch, global_step)
        self.variance.update_step(epoch, global_step)
        end = self.cfg.get("cos_anneal_end", 0)
        self.cos_anneal_ratio = 1.0 if end == 0 else min(1.0, global_step / end)
        if self.training and self.grid_prune and update_occupancy:
            oi = occ_inputs or {}
            self.occupancy_grid.every_n_step(global_step, self.occ_eval_fn, occ_thre=self.cfg.get("grid_prune_occ_thre", 0.01),
                                             indices=oi.get("indices"), jitter=oi.get("jitter"))
            if self.learned_background:
                self.occupancy_grid_bg.every_n_step(global_step, self.occ_eval_fn_bg,
                                                    occ_thre=self.cfg.get("grid_prune_occ_thre_bg", 0.01),
                                                    indices=oi.get("indices_bg"), jitter=oi.get("jitter_bg"))

    # -- neus.py:117-139 -----------------------------------------------------------------------
    def get_alpha(self, sdf, normal, dirs, dists):
        inv_s = self.variance.inv_s.reshape(1, 1).clip(1e-6, 1e6).expand(sdf.shape[0], 1)
        true_cos = (dirs * normal).sum(-1, keepdim=True)
        a = self.cos_anneal_ratio
        iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - a) + F.relu(-true_cos) * a)
        nxt = sdf[..., None] + iter_cos * dists.reshape(-1, 1) * 0.5
        prv = sdf[..., None] - iter_cos * dists.reshape(-1, 1) * 0.5
        prev_cdf, next_cdf = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
        return ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).view(-1).clip(0.0, 1.0)

    # -- neus.py:141-203 -----------------------------------------------------------------------
    def forward_bg_(self, rays, stratified_u=None):
        n_rays = rays.shape[0]
        rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]

        def sigma_fn(t_starts, t_ends, ray_indices):
            ri = ray_indices.long()
            pos = rays_o[ri] + rays_d[ri] * (t_starts + t_ends) / 2.0
            density, _ = self.geometry_bg(pos)
            return density[..., None]

        _, t_max = nf.ray_aabb_intersect(rays_o, rays_d, self.scene_aabb)
        near_plane = torch.where(t_max > 1e9, torch.tensor(self.near_plane_bg), t_max)
        with torch.no_grad():
            ray_indices, t_starts, t_ends = nf.ray_marching(
                rays_o, rays_d, scene_aabb=None, grid=self.occupancy_grid_bg if self.grid_prune else None,
                sigma_fn=sigma_fn, near_plane=near_plane, far_plane=self.far_plane_bg,
                render_step_size=self.render_step_size_bg, stratified=self.randomized,
                cone_angle=self.cone_angle_bg, alpha_thre=0.0, stratified_u=stratified_u)
        ri = ray_indices.long()
        t_dirs = rays_d[ri]
        midpoints = (t_starts + t_ends) / 2.0
        positions = rays_o[ri] + t_dirs * midpoints
        density, feature = self.geometry_bg(positions)
        rgb = self.texture_bg(feature, t_dirs)
        weights = nf.render_weight_from_density(t_starts, t_ends, density[..., None], ray_indices=ri, n_rays=n_rays)
        opacity = nf.accumulate_along_rays(weights, ri, values=None, n_rays=n_rays)
        depth = nf.accumulate_along_rays(weights, ri, values=midpoints, n_rays=n_rays)
        comp_rgb = nf.accumulate_along_rays(weights, ri, values=rgb, n_rays=n_rays)
        comp_rgb = comp_rgb + self.background_color * (1.0 - opacity)
        return {"comp_rgb": comp_rgb, "opacity": opacity, "depth": depth, "rays_valid": opacity > 0,
                "num_samples": torch.tensor([len(t_starts)], dtype=torch.int32),
                "weights": weights.view(-1), "points": midpoints.view(-1), "intervals": (t_ends - t_starts).view(-1),
                "ray_indices": ri.view(-1), "t_starts": t_starts, "t_ends": t_ends}

    # -- neus.py:205-283 -----------------------------------------------------------------------
    def forward_(self, rays, stratified_u=None, rand_directions=None, stratified_u_bg=None):
        n_rays = rays.shape[0]
        rays_o, rays_d = rays[:, 0:3], rays[:, 3:6]
        with torch.no_grad():
            ray_indices, t_starts, t_ends = nf.ray_marching(
                rays_o, rays_d, scene_aabb=self.scene_aabb, grid=self.occupancy_grid if self.grid_prune else None,
                alpha_fn=None, near_plane=None, far_plane=None, render_step_size=self.render_step_size,
                stratified=self.randomized, cone_angle=0.0, alpha_thre=0.0, stratified_u=stratified_u)
        ri = ray_indices.long()
        t_dirs = rays_d[ri]
        midpoints = (t_starts + t_ends) / 2.0
        positions = rays_o[ri] + t_dirs * midpoints
        dists = t_ends - t_starts
        sdf, sdf_grad, feature, sdf_laplace = self.geometry(positions, with_grad=True, with_feature=True,
                                                            with_laplace=True, rand_directions=rand_directions)
        normal = F.normalize(sdf_grad, p=2, dim=-1)
        alpha = self.get_alpha(sdf, normal, t_dirs, dists)[..., None]
        rgb = self.texture(feature, t_dirs, normal)
        weights = nf.render_weight_from_alpha(alpha, ray_indices=ri, n_rays=n_rays)
        opacity = nf.accumulate_along_rays(weights, ri, values=None, n_rays=n_rays)
        depth = nf.accumulate_along_rays(weights, ri, values=midpoints, n_rays=n_rays)
        comp_rgb = nf.accumulate_along_rays(weights, ri, values=rgb, n_rays=n_rays)
        rays_fg = opacity > 0.1
        comp_normal = nf.accumulate_along_rays(weights, ri, values=normal, n_rays=n_rays)
        comp_normal = F.normalize(comp_normal, p=2, dim=-1)
        comp_normal = comp_normal * rays_fg.float()
        out = {"comp_rgb": comp_rgb, "comp_normal": comp_normal, "opacity": opacity, "depth": depth,
               "rays_valid": opacity > 0, "num_samples": torch.tensor([len(t_starts)], dtype=torch.int32),
               "sdf_samples": sdf, "sdf_grad_samples": sdf_grad, "weights": weights.view(-1),
               "points": midpoints.view(-1), "intervals": dists.view(-1), "ray_indices": ri.view(-1),
               "sdf_laplace_samples": sdf_laplace, "t_starts": t_starts, "t_ends": t_ends, "alpha": alpha.view(-1)}
        if self.learned_background:
            out_bg = self.forward_bg_(rays, stratified_u=stratified_u_bg)
        else:
            out_bg = {"comp_rgb": self.background_color[None, :].expand(*comp_rgb.shape),
                      "num_samples": torch.zeros_like(out["num_samples"]), "rays_valid": torch.zeros_like(out["rays_valid"])}
        out_full = {"comp_rgb": out["comp_rgb"] + out_bg["comp_rgb"] * (1.0 - out["opacity"]),
                    "num_samples": out["num_samples"] + out_bg["num_samples"],
                    "rays_valid": out["rays_valid"] | out_bg["rays_valid"]}
        return {**out, **{k + "_bg": v for k, v in out_bg.items()}, **{k + "_full": v for k, v in out_full.items()},
                "inv_s": self.variance.inv_s}


# ---------------------------------------------------------------------------------------------
# losses (systems/neus.py:130-194)
# ---------------------------------------------------------------------------------------------

def binary_cross_entropy(inp, target):
    return -(target * torch.log(inp) + (1 - target) * torch.log(1 - inp)).mean()


def training_loss(model: RefNeuSModel, out: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], loss_cfg: dict,
                  global_step: int, has_mask: bool = False) -> Dict[str, torch.Tensor]:
    C = lambda v: schedule_value(v, global_step)
    terms = {}
    valid = out["rays_valid_full"][..., 0]
    terms["rgb_mse"] = F.mse_loss(out["comp_rgb_full"][valid], batch["rgb"][valid])
    loss = terms["rgb_mse"] * C(loss_cfg["lambda_rgb_mse"])
    terms["rgb_l1"] = F.l1_loss(out["comp_rgb_full"][valid], batch["rgb"][valid])
    loss = loss + terms["rgb_l1"] * C(loss_cfg["lambda_rgb_l1"])
    terms["eikonal"] = ((torch.linalg.norm(out["sdf_grad_samples"], ord=2, dim=-1) - 1.0) ** 2).mean()
    loss = loss + terms["eikonal"] * C(loss_cfg["lambda_eikonal"])
    opacity = torch.clamp(out["opacity"].squeeze(-1), 1e-3, 1 - 1e-3)
    if has_mask and "fg_mask" in batch:
        terms["mask"] = binary_cross_entropy(opacity, batch["fg_mask"].float())
        loss = loss + terms["mask"] * C(loss_cfg["lambda_mask"])
    terms["opaque"] = binary_cross_entropy(opacity, opacity)
    loss = loss + terms["opaque"] * C(loss_cfg["lambda_opaque"])
    terms["sparsity"] = torch.exp(-loss_cfg["sparsity_scale"] * out["sdf_samples"].abs()).mean()
    loss = loss + terms["sparsity"] * C(loss_cfg["lambda_sparsity"])
    if C(loss_cfg["lambda_curvature"]) > 0:
        terms["curvature"] = out["sdf_laplace_samples"].abs().mean()
        loss = loss + terms["curvature"] * C(loss_cfg["lambda_curvature"])
    if C(loss_cfg["lambda_sdf_l1"]) > 0 and "pts" in batch:
        sdf_p, grad_p = model.geometry(batch["pts"], with_grad=True, with_feature=False)
        # Appendix C-11: scalar L1 mean times the per-point weights, then mean
        terms["sdf_l1"] = (F.l1_loss(sdf_p, torch.zeros_like(sdf_p)) * batch["pts_weights"]).mean(dim=0)
        n_gt = F.normalize(batch["pts_normal"], p=2, dim=-1)
        n_pr = F.normalize(grad_p, p=2, dim=-1)
        terms["normal_cos"] = (1.0 - torch.sum(n_pr * n_gt, dim=-1)).mean()
        loss = loss + terms["sdf_l1"] * C(loss_cfg["lambda_sdf_l1"])
        loss = loss + terms["normal_cos"] * C(loss_cfg.get("lambda_normal", loss_cfg["lambda_sdf_l1"]))
    terms["loss"] = loss
    return terms
