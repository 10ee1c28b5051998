"""Geometry fields with the reference's surface (models/geometry.py): VolumeSDF and VolumeDensity.

VolumeSDF.forward reproduces reference models/geometry.py:195-287 for grad_type='finite_difference'
bug-for-bug: taps clamped in world space, curvature shift applied to the
normalised coordinates and then re-normalised as if it were a world point, un-normalised tangent, no
detach on the shifted tap positions.  The 1 + 6 + 6 network evaluations per sample run as three
hash-grid launches + three fused-MLP launches; the 12 tap evaluations compute the SDF column only.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import ops
from . import registry as models
from .nerfacc_api import ContractionType
from .network_utils import fused_encode_mlp, get_encoding, get_encoding_with_network, get_mlp, update_module_step
from .utils import get_activation, scale_anything

# Rows of a centre evaluation are the packed samples of the marched rays: consecutive rows are neighbours along a ray (one
# render step apart), so at the coarse levels runs of them share a cell exactly as the six finite-difference taps of a sample
# do -- the grouped scatter (ia_hashgrid_bwd_grouped: same-cell contributions merged in registers before the atomic adds)
# applies to them too.  IA_CENTER_GROUP=1 restores the plain scatter (A/B).
_CENTER_GROUP = int(os.environ.get("IA_CENTER_GROUP", "6"))


def contract_to_unisphere(x, radius, contraction_type):
    """reference models/geometry.py:19-31."""
    if x.is_cuda and not x.requires_grad and x.shape[-1] == 3 and x.dtype == torch.float32 and \
            contraction_type in (ContractionType.AABB, ContractionType.UN_BOUNDED_SPHERE):
        return ops.contract(x, float(radius), int(contraction_type))           # one launch instead of 3 / 11
    if contraction_type == ContractionType.AABB:
        return scale_anything(x, (-radius, radius), (0, 1))
    if contraction_type == ContractionType.UN_BOUNDED_SPHERE:
        x = scale_anything(x, (-radius, radius), (0, 1))
        x = x * 2 - 1
        mag = x.norm(dim=-1, keepdim=True)
        x = torch.where(mag > 1, (2 - 1 / mag) * (x / mag), x)
        return x / 4 + 0.5
    raise NotImplementedError


class BaseImplicitGeometry(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.radius = config["radius"]
        self.contraction_type = None  # assigned by the model, as in the reference
        self.setup()

    def regularizations(self, out):
        return {}

    def forward_level(self, points):
        raise NotImplementedError

    @torch.no_grad()
    def isosurface(self):
        """reference models/geometry.py:80-113: marching cubes over the +-radius box on the lattice of `isosurface.resolution`
        steps per 2 units, at the reference's fixed threshold 0.001; the SDF blocks are evaluated by forward_level on the
        parameters' device and meshed there (isosurface.py)."""
        isocfg = self.config.get("isosurface", None)
        if isocfg is None:
            raise NotImplementedError
        assert isocfg["method"] in ["mc", "CuMCubes"]
        from .isosurface import MarchingCubeHelper
        r = self.radius
        device = next(self.parameters()).device
        helper = MarchingCubeHelper(lambda x: -self.forward_level(x), [(-r, r), (-r, r), (-r, r)], int(isocfg["resolution"]),
                                    block_res=int(isocfg.get("block_res", 256)), method=isocfg["method"], device=device)
        return helper(threshold=0.001)


@models.register("volume-density")
class VolumeDensity(BaseImplicitGeometry):
    """reference models/geometry.py:152-177 (background NeRF++ density)."""

    def setup(self):
        self.n_input_dims = self.config.get("n_input_dims", 3)
        self.n_output_dims = self.config["feature_dim"]
        self.encoding_with_network = get_encoding_with_network(self.n_input_dims, self.n_output_dims,
                                                               self.config["xyz_encoding_config"],
                                                               self.config["mlp_network_config"])

    def forward(self, points):
        points = contract_to_unisphere(points, self.radius, self.contraction_type)
        out = self.encoding_with_network(points.view(-1, self.n_input_dims), group=_CENTER_GROUP).view(*points.shape[:-1], self.n_output_dims).float()
        density, feature = out[..., 0], out
        if "density_activation" in self.config:
            density = get_activation(self.config["density_activation"])(density + float(self.config["density_bias"]))
        if "feature_activation" in self.config:
            feature = get_activation(self.config["feature_activation"])(feature)
        return density, feature

    def density(self, points):
        """forward(points)[0] without the feature columns (the output layer evaluates column 0 only): what the
        background marching's visibility pruning (models/neus.py:144-149 sigma_fn) and occ_eval_fn_bg (:103-106) consume."""
        points = contract_to_unisphere(points, self.radius, self.contraction_type)
        density = self.encoding_with_network(points.reshape(-1, self.n_input_dims), n_out_used=1).reshape(*points.shape[:-1]).float()
        if "density_activation" in self.config:
            density = get_activation(self.config["density_activation"])(density + float(self.config["density_bias"]))
        return density

    def forward_level(self, points):
        points = contract_to_unisphere(points, self.radius, self.contraction_type)
        density = self.encoding_with_network(points.reshape(-1, self.n_input_dims), n_out_used=1).reshape(*points.shape[:-1])
        if "density_activation" in self.config:
            density = get_activation(self.config["density_activation"])(density + float(self.config["density_bias"]))
        return -density

    def update_step(self, epoch, global_step):
        update_module_step(self.encoding_with_network, epoch, global_step)


@models.register("volume-sdf")
class VolumeSDF(BaseImplicitGeometry):
    """reference models/geometry.py:180-314."""

    def setup(self):
        self.n_output_dims = self.config["feature_dim"]
        self.encoding = get_encoding(3, self.config["xyz_encoding_config"])
        self.network = get_mlp(self.encoding.n_output_dims, self.n_output_dims, self.config["mlp_network_config"])
        self.grad_type = self.config["grad_type"]
        self.finite_difference_eps = self.config.get("finite_difference_eps", 1e-3)
        self._finite_difference_eps = None
        self.register_buffer("_fd_signs", torch.tensor([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [0, -1.0, 0],
                                                        [0, 0, 1.0], [0, 0, -1.0]]), persistent=False)

    def _net(self, pts01, n_out_used, flat, group=1):
        return fused_encode_mlp(self.encoding, self.network, pts01.reshape(-1, 3), n_out_used, flat, group)

    def _analytic_grid(self):
        from .network_utils import Encoding, ProgressiveBandHashGrid
        inner = self.encoding.encoding
        if isinstance(inner, ProgressiveBandHashGrid):
            return inner.encoding, inner.active_levels        # the tcnn-style Encoding holding .params
        if isinstance(inner, Encoding) and inner.otype == "HashGrid":
            return inner, inner.plan.n_levels
        return None, None

    def _analytic_hidden(self, pts01, flat=None):
        """Kernel path of grad_type 'analytic': (last hidden layer [S,64], grad [*,3] w.r.t. the un-normalised points, flat) or
        None when the shape is not the one ia_mlp_fwd_grad implements (then _analytic_eval differentiates torch operators)."""
        enc_mod, net = self.encoding, self.network
        grid, active = self._analytic_grid()
        if grid is None or not enc_mod.include_xyz or "sdf_activation" in self.config or \
                not (net.output_activation_name is None or str(net.output_activation_name).lower() == "none"):
            return None
        desc = ops.make_mlp_desc(3, grid.n_output_dims, net.n_hidden_layers, net.n_output_dims, net.hidden_act,
                                 enc_mod.xyz_scale, enc_mod.xyz_offset, net.precision)
        if not ops.mlp_fwd_grad_supported(desc):
            return None
        x = pts01.reshape(-1, 3)
        flat = net.flat_params() if flat is None else flat
        enc = ops.hashgrid_encode(x, grid.params, grid.plan, active, table_h=grid.shadow())
        h, g0, g1 = ops.mlp_fwd_grad(x, enc, flat, desc)
        g01 = ops.hashgrid_input_grad(x, grid.params, g1, grid.plan, active, grid.shadow()) + g0
        # points01 = (points + r) / (2 r)   (AABB contraction, models/geometry.py:24)
        grad = (g01 / (2.0 * self.radius)).view(*pts01.shape[:-1], 3)
        return h, grad, flat

    def _analytic_eval(self, pts01, flat=None):
        """grad_type 'analytic' (reference models/geometry.py:198-218): the centre evaluation together with
        d sdf / d points by the chain rule, kept differentiable (create_graph) so that the eikonal / rendering losses
        back-propagate through the normals.  Per sample: hash-grid forward (kernel) -> network as torch operators
        (VanillaMLP.forward_twice_differentiable) -> d sdf / d(network input) by autograd -> J_enc^T (.) through
        ops.hashgrid_input_grad, whose own backward is the pair of second-order hash-grid kernels.
        Returns (out [S, n_output_dims], grad [*, 3] w.r.t. the un-normalised points)."""
        from .network_utils import Encoding, ProgressiveBandHashGrid
        enc_mod = self.encoding
        inner = enc_mod.encoding
        x = pts01.reshape(-1, 3)
        if isinstance(inner, ProgressiveBandHashGrid):
            grid, active = inner.encoding, inner.active_levels        # the tcnn-style Encoding holding .params
        elif isinstance(inner, Encoding) and inner.otype == "HashGrid":
            grid, active = inner, inner.plan.n_levels
        else:
            raise NotImplementedError("grad_type='analytic' is implemented for (ProgressiveBand)HashGrid encodings")
        hg = self._analytic_hidden(pts01, flat)
        if hg is not None:
            # tensor-core path: network and d sdf / d(network input) in one kernel whose backward is the second-order adjoint
            # (ops.mlp_fwd_grad); the wide output layer by linear64 as for the first-order centre evaluation
            h, grad, flat = hg
            net = self.network
            n_out, width = net.n_output_dims, net.n_neurons
            n_hidden = flat.numel() - (n_out * width + n_out)
            out = ops.linear64(h, flat[n_hidden:n_hidden + n_out * width].view(n_out, width), flat[n_hidden + n_out * width:])
            return out, grad
        enc = ops.hashgrid_encode(x, grid.params, grid.plan, active, table_h=grid.shadow())
        parts = ([x * enc_mod.xyz_scale + enc_mod.xyz_offset] if enc_mod.include_xyz else []) + [enc]
        e = torch.cat(parts, dim=-1)
        if not e.requires_grad:
            e.requires_grad_(True)
        out = self.network.forward_twice_differentiable(e).float()
        sdf = out[:, 0]
        if "sdf_activation" in self.config:
            sdf = get_activation(self.config["sdf_activation"])(sdf + float(self.config["sdf_bias"]))
        keep = self.training
        (g_e,) = torch.autograd.grad(sdf, e, grad_outputs=torch.ones_like(sdf), create_graph=keep, retain_graph=True)
        n_xyz = 3 if enc_mod.include_xyz else 0
        g01 = ops.hashgrid_input_grad(x, grid.params, g_e[:, n_xyz:].contiguous(), grid.plan, active, grid.shadow())
        if enc_mod.include_xyz:
            g01 = g01 + g_e[:, :3] * enc_mod.xyz_scale
        # points01 = (points + r) / (2 r)   (AABB contraction, models/geometry.py:24)
        grad = (g01 / (2.0 * self.radius)).view(*pts01.shape[:-1], 3)
        return out, grad

    def _fd_gradient(self, world_pts, eps, flat):
        """geometry.py:219-234: six taps clamped to the AABB in world space, central differences."""
        lead = world_pts.shape[:-1]
        taps01 = ops.fd_taps(world_pts.reshape(-1, 3), eps, self.radius)          # fused add / clamp / normalise
        s = self._net(taps01, 1, flat, group=6).view(-1, 6)     # rows = 6 taps per sample: grouped scatter in backward
        return ops.fd_grad(s, eps).view(*lead, 3)                                 # fused central differences

    def forward(self, points, with_grad=True, with_feature=True, with_laplace=False, with_auxiliary_feature=False,
                rand_directions: Optional[torch.Tensor] = None):
        if with_auxiliary_feature:
            raise NotImplementedError("with_auxiliary_feature is not used by any shipped config")
        analytic = with_grad and self.grad_type == "analytic"
        with torch.set_grad_enabled((self.training and torch.is_grad_enabled()) or analytic):
            flat = self.network.flat_params()
            points_ = points
            pts01 = contract_to_unisphere(points, self.radius, self.contraction_type)
            need_full = with_feature
            grad = None
            if analytic:
                out, grad = self._analytic_eval(pts01, flat)
            else:
                out = self._net(pts01, self.n_output_dims if need_full else 1, flat, group=_CENTER_GROUP)
            out = out.view(*pts01.shape[:-1], out.shape[-1])
            sdf = out[..., 0]
            feature = None
            if with_feature:
                feature = torch.cat([out, pts01 * 2 - 1], dim=-1)
            if "sdf_activation" in self.config:
                sdf = get_activation(self.config["sdf_activation"])(sdf + float(self.config["sdf_bias"]))
            if with_feature and "feature_activation" in self.config:
                feature = get_activation(self.config["feature_activation"])(feature)
            if with_grad and not analytic:
                grad = self._fd_gradient(points_, self._finite_difference_eps, flat)
            laplace = None
            if with_laplace:
                eps = self._finite_difference_eps
                if rand_directions is None:
                    rand_directions = torch.randn_like(pts01)
                lead = pts01.shape[:-1]
                # normals = normalize(grad); shifted = pts01 + cross(normals, normalize(rnd)) * eps   (Appendix C-1/2/3)
                normals, shifted = ops.curv_shift(grad.reshape(-1, 3), rand_directions.reshape(-1, 3), pts01.reshape(-1, 3), eps)
                g_shift = self._fd_gradient(shifted, eps, flat)
                laplace = ops.curv_angle(normals, g_shift).view(*lead, 1)
        rv = [sdf]
        if with_grad:
            rv.append(grad)
        if with_feature:
            rv.append(feature)
        if with_laplace:
            rv.append(laplace)
        rv = [v if self.training else v.detach() for v in rv]
        return rv[0] if len(rv) == 1 else rv

    # ---- fused-head path (used by NeuSModel.forward_ when the colour head supports it) ------------------------------
    def supports_fused_head(self) -> bool:
        """True when forward(with_grad, with_feature, with_laplace) can hand the centre evaluation's last hidden layer to
        the colour head (ops.sdf_head) instead of materialising `feature`: tensor-core MLP with a wide output layer,
        finite-difference gradients, no output activations."""
        from . import _lib as L
        act = self.network.output_activation_name
        return (os.environ.get("IA_NO_FUSED_HEAD") is None and self.network.precision == L.IA_MLP_TC_F16
                and self.n_output_dims > 8 and (self.grad_type == "finite_difference" or self._analytic_kernel_ok())
                and "sdf_activation" not in self.config and "feature_activation" not in self.config
                and (act is None or str(act).lower() == "none"))

    def _analytic_kernel_ok(self) -> bool:
        if self.grad_type != "analytic":
            return False
        grid, _ = self._analytic_grid()
        net = self.network
        if grid is None or not self.encoding.include_xyz:
            return False
        desc = ops.make_mlp_desc(3, grid.n_output_dims, net.n_hidden_layers, net.n_output_dims, net.hidden_act, 1.0, 0.0, net.precision)
        return ops.mlp_fwd_grad_supported(desc)

    def forward_hidden(self, points, rand_directions: Optional[torch.Tensor] = None):
        """Same evaluations as forward(points, with_grad=True, with_feature=True, with_laplace=True) (reference
        models/geometry.py:195-275) except that the centre evaluation stops at the last hidden layer.
        Returns (h [S,64], pts01 [S,3], grad [S,3], laplace [S,1], W_last [Fd,64], b_last [Fd])."""
        with torch.set_grad_enabled((self.training and torch.is_grad_enabled()) or self.grad_type == "analytic"):
            flat = self.network.flat_params()
            pts01 = contract_to_unisphere(points, self.radius, self.contraction_type)
            eps = self._finite_difference_eps
            if self.grad_type == "analytic":
                h, grad, flat = self._analytic_hidden(pts01, flat)
            else:
                h = self._net(pts01, 0, flat, group=_CENTER_GROUP)
                grad = self._fd_gradient(points, eps, flat)
            if rand_directions is None:
                rand_directions = torch.randn_like(pts01)
            normals, shifted = ops.curv_shift(grad.reshape(-1, 3), rand_directions.reshape(-1, 3), pts01.reshape(-1, 3), eps)
            g_shift = self._fd_gradient(shifted, eps, flat)
            laplace = ops.curv_angle(normals, g_shift).view(*pts01.shape[:-1], 1)
            n_out, width = self.n_output_dims, self.network.n_neurons
            n_hidden = flat.numel() - (n_out * width + n_out)
            w_last = flat[n_hidden:n_hidden + n_out * width].view(n_out, width)
            b_last = flat[n_hidden + n_out * width:]
        return h, pts01, grad, laplace, w_last, b_last

    def forward_level(self, points):
        pts01 = contract_to_unisphere(points, self.radius, self.contraction_type)
        sdf = self._net(pts01, 1, None).view(*pts01.shape[:-1])
        if "sdf_activation" in self.config:
            sdf = get_activation(self.config["sdf_activation"])(sdf + float(self.config["sdf_bias"]))
        return sdf

    def update_step(self, epoch, global_step):
        update_module_step(self.encoding, epoch, global_step)
        update_module_step(self.network, epoch, global_step)
        if isinstance(self.finite_difference_eps, float):
            self._finite_difference_eps = self.finite_difference_eps
        elif self.finite_difference_eps == "progressive":
            hg = self.config["xyz_encoding_config"]
            assert hg["otype"] == "ProgressiveBandHashGrid", \
                "finite_difference_eps='progressive' only works with ProgressiveBandHashGrid"
            level = min(hg["start_level"] + max(global_step - hg["start_step"], 0) // hg["update_steps"], hg["n_levels"])
            grid_res = hg["base_resolution"] * hg["per_level_scale"] ** (level - 1)
            self._finite_difference_eps = 2 * self.config["radius"] / grid_res
        else:
            raise ValueError(f"Unknown finite_difference_eps={self.finite_difference_eps}")
