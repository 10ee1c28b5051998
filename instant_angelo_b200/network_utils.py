"""Encoding + network factories with the reference's module surface (models/network_utils.py), backed by
the sm_100a kernels.

Same names and attributes (`n_input_dims`, `n_output_dims`, `update_step(epoch, global_step)`), same
state-dict keys (`encoding.encoding.params`, `layers.{0,2,4}.{weight_g,weight_v,bias}`), so a reference
checkpoint loads unchanged.  What differs is the execution: the progressive mask is folded into the
hash-grid kernel as a level count, the xyz pass-through of CompositeEncoding and the weight-norm of
VanillaMLP are folded into the fused MLP operator (`EncodingWithNetwork.forward` never materialises the
concatenated input).
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .config import to_primitive


def update_module_step(m, epoch, global_step):
    """reference systems/utils.py:349-351."""
    if hasattr(m, "update_step"):
        m.update_step(epoch, global_step)


class Encoding(nn.Module):
    """tcnn.Encoding(n_input_dims, encoding_config) for otype HashGrid | SphericalHarmonics
    (reference models/network_utils.py:47, 91).  `.params` is the flat fp32 tcnn parameter vector."""

    def __init__(self, n_input_dims: int, encoding_config: dict, seed: int = 1337):
        super().__init__()
        cfg = to_primitive(encoding_config)
        self.n_input_dims = n_input_dims
        self.otype = cfg["otype"]
        if self.otype == "HashGrid":
            assert n_input_dims == 3, "HashGrid is implemented for 3-D inputs"
            self.plan = ops.make_grid_plan(cfg["n_levels"], cfg["n_features_per_level"], cfg["log2_hashmap_size"],
                                           cfg["base_resolution"], cfg["per_level_scale"])
            g = torch.Generator().manual_seed(seed)
            self.params = nn.Parameter((torch.rand(self.plan.n_params, generator=g) * 2 - 1) * 1e-4)
            self.n_output_dims = self.plan.n_levels * self.plan.n_features
            # `table_precision: fp16` (not a tcnn key; IA_TABLE_FP16=1 sets it for every grid): the gathers read an fp16
            # shadow of .params -- tcnn's own arrangement of fp32 master parameters + fp16 compute copy.  Off by default:
            # results then differ from the fp32-table oracle by the rounding of the table entries (2^-11 relative).
            prec = str(cfg.get("table_precision", "fp16" if os.environ.get("IA_TABLE_FP16", "0") not in ("0", "") else "fp32"))
            if prec not in ("fp32", "fp16"):
                raise ValueError(f"table_precision must be fp32 or fp16 (got {prec})")
            self.table_precision = prec
            self._shadow, self._shadow_stale, self._shadow_version, self._shadow_epoch = None, True, -1, -1
        elif self.otype == "SphericalHarmonics":
            assert n_input_dims == 3
            self.degree = int(cfg["degree"])
            self.params = nn.Parameter(torch.zeros(0))
            self.n_output_dims = self.degree ** 2
        else:
            raise NotImplementedError(f"encoding otype {self.otype}")

    def shadow(self) -> Optional[torch.Tensor]:
        """The fp16 copy of .params the gathers read (None under table_precision fp32).  Re-derived when the parameters may
        have changed: after any fused optimizer step of this process (ops.param_epoch(): ia_adamw_step writes through raw
        pointers, which torch's version counter does not see), after update_step() (the reference's per-step hook) and after
        any in-place torch operation on .params (load_state_dict, a torch optimizer)."""
        if self.otype != "HashGrid" or self.table_precision != "fp16":
            return None
        p = self.params
        if self._shadow is None or self._shadow.device != p.device or self._shadow_stale or self._shadow_version != p._version \
                or self._shadow_epoch != ops.param_epoch():
            if self._shadow is not None and self._shadow.device != p.device:
                self._shadow = None
            self._shadow = ops.table_to_half(p, self._shadow)
            self._shadow_stale, self._shadow_version, self._shadow_epoch = False, p._version, ops.param_epoch()
        return self._shadow

    def update_step(self, epoch, global_step):
        if self.otype == "HashGrid":
            self._shadow_stale = True

    def forward(self, x: torch.Tensor, active_levels: Optional[int] = None, group: int = 1) -> torch.Tensor:
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        if self.otype == "HashGrid":
            return ops.hashgrid_encode(x, self.params, self.plan, active_levels, group, self.shadow())
        return ops.sh_encode(x, self.degree)


class ProgressiveBandHashGrid(nn.Module):
    """reference models/network_utils.py:40-66.  `mask` is kept (same attribute, same growth rule) but the
    forward pass hands the kernel `current_level` instead of multiplying: masked levels come out as exact
    zeros and receive exactly-zero gradients, which is what the multiply does."""

    def __init__(self, in_channels: int, config: dict):
        super().__init__()
        config = to_primitive(config)
        self.n_input_dims = in_channels
        encoding_config = dict(config)
        encoding_config["otype"] = "HashGrid"
        self.encoding = Encoding(in_channels, encoding_config)
        self.n_output_dims = self.encoding.n_output_dims
        self.n_level = config["n_levels"]
        self.n_features_per_level = config["n_features_per_level"]
        self.start_level, self.start_step, self.update_steps = config["start_level"], config["start_step"], config["update_steps"]
        self.current_level = self.start_level
        self._mask_level = self.start_level  # the mask only ever grows (Appendix C-7)
        self.mask = torch.zeros(self.n_level * self.n_features_per_level, dtype=torch.float32)
        self.mask[: self.current_level * self.n_features_per_level] = 1.0

    @property
    def active_levels(self) -> int:
        return self._mask_level

    def forward(self, x: torch.Tensor, group: int = 1) -> torch.Tensor:
        return self.encoding(x, self._mask_level, group)

    def update_step(self, epoch, global_step):
        self.current_level = min(self.start_level + max(global_step - self.start_step, 0) // self.update_steps, self.n_level)
        self._mask_level = max(self._mask_level, self.current_level)
        self.mask[: self.current_level * self.n_features_per_level] = 1.0
        self.encoding.update_step(epoch, global_step)


class CompositeEncoding(nn.Module):
    """reference models/network_utils.py:69-80."""

    def __init__(self, encoding, include_xyz=False, xyz_scale=1.0, xyz_offset=0.0):
        super().__init__()
        self.encoding = encoding
        self.include_xyz, self.xyz_scale, self.xyz_offset = include_xyz, xyz_scale, xyz_offset
        self.n_input_dims = encoding.n_input_dims
        self.n_output_dims = int(self.include_xyz) * self.encoding.n_input_dims + self.encoding.n_output_dims

    def forward(self, x, *args):
        enc = self.encoding(x, *args)
        return enc if not self.include_xyz else torch.cat([x * self.xyz_scale + self.xyz_offset, enc], dim=-1)

    def update_step(self, epoch, global_step):
        update_module_step(self.encoding, epoch, global_step)


def get_encoding(n_input_dims: int, config) -> CompositeEncoding:
    """reference models/network_utils.py:83-93 (input is supposed to be in [0, 1])."""
    otype = config["otype"]
    if otype == "ProgressiveBandHashGrid":
        encoding = ProgressiveBandHashGrid(n_input_dims, config)
    elif otype in ("HashGrid", "SphericalHarmonics"):
        encoding = Encoding(n_input_dims, config)
    else:
        raise NotImplementedError(f"encoding otype {otype} is outside the B200 hot path")
    return CompositeEncoding(encoding, include_xyz=config.get("include_xyz", False), xyz_scale=2.0, xyz_offset=-1.0)


class _Linear(nn.Module):
    """Parameter holder with nn.Linear / torch weight_norm key names (weight_g, weight_v, bias | weight, bias)."""

    def __init__(self, dim_in: int, dim_out: int, weight_norm: bool):
        super().__init__()
        self.dim_in, self.dim_out, self.weight_norm = dim_in, dim_out, weight_norm
        # registration order = torch's: nn.Linear registers (weight, bias); nn.utils.weight_norm then deletes `weight` and
        # appends (weight_g, weight_v), leaving (bias, weight_g, weight_v).  torch.optim state dicts index parameters by
        # position, so a reference checkpoint's optimizer_states only land on the right tensors with this order.
        if weight_norm:
            self.bias = nn.Parameter(torch.zeros(dim_out))
            self.weight_g = nn.Parameter(torch.ones(dim_out, 1))
            self.weight_v = nn.Parameter(torch.empty(dim_out, dim_in))
        else:
            self.weight = nn.Parameter(torch.empty(dim_out, dim_in))
            self.bias = nn.Parameter(torch.zeros(dim_out))

    def raw_weight(self) -> torch.Tensor:
        return self.weight_v if self.weight_norm else self.weight

    def effective_weight(self) -> torch.Tensor:
        if not self.weight_norm:
            return self.weight
        return self.weight_v * (self.weight_g / self.weight_v.norm(dim=1, keepdim=True))


class _ActMarker(nn.Module):
    """Occupies the activation slots of the reference's nn.Sequential so layer indices (0, 2, 4) match."""

    def __init__(self, name: str):
        super().__init__()
        self.name = name

    def extra_repr(self):
        return self.name


class VanillaMLP(nn.Module):
    """reference models/network_utils.py:96-140, executed by the fused width-64 MLP kernel.

    precision: IA_MLP_FP32 (default; the reference runs this module in fp32 with autocast disabled) or
    IA_MLP_TC_F16 (tensor cores, selected by get_mlp for otype FullyFusedMLP / CutlassMLP)."""

    def __init__(self, dim_in: int, dim_out: int, config: dict, precision: int = L.IA_MLP_FP32):
        super().__init__()
        config = to_primitive(config)
        self.n_input_dims, self.n_output_dims = dim_in, dim_out
        self.n_neurons, self.n_hidden_layers = config["n_neurons"], config["n_hidden_layers"]
        if self.n_neurons != 64 or self.n_hidden_layers not in (1, 2):
            raise NotImplementedError("the fused MLP supports n_neurons=64 and 1-2 hidden layers")
        self.sphere_init, self.weight_norm = config.get("sphere_init", False), config.get("weight_norm", False)
        self.sphere_init_radius = config.get("sphere_init_radius", 0.5)
        self.precision = precision
        dims = [dim_in] + [self.n_neurons] * self.n_hidden_layers + [dim_out]
        mods = []
        for i in range(len(dims) - 1):
            mods.append(self.make_linear(dims[i], dims[i + 1], is_first=(i == 0), is_last=(i == len(dims) - 2)))
            if i < len(dims) - 2:
                mods.append(_ActMarker("Softplus(beta=100)" if self.sphere_init else "ReLU"))
        self.layers = nn.Sequential(*mods)
        self.output_activation_name = config.get("output_activation", None)
        self.hidden_act = L.IA_ACT_SOFTPLUS100 if self.sphere_init else L.IA_ACT_RELU

    def make_linear(self, dim_in, dim_out, is_first, is_last) -> _Linear:
        layer = _Linear(dim_in, dim_out, self.weight_norm)
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
            if self.weight_norm:
                layer.weight_g.copy_(w.norm(dim=1, keepdim=True))
        return layer

    def linears(self):
        return [m for m in self.layers if isinstance(m, _Linear)]

    def flat_params(self) -> torch.Tensor:
        """Effective weights (weight-norm folded) in the ABI's flat layout; differentiable."""
        lins = self.linears()
        if lins[0].bias.is_cuda:
            # one launch (and one in backward) instead of norm / div / mul / cat per layer
            return ops.weightnorm_flat([(lin.weight_g if lin.weight_norm else None, lin.raw_weight(), lin.bias) for lin in lins])
        parts = []
        for lin in lins:
            parts.append(lin.effective_weight().reshape(-1))
            parts.append(lin.bias)
        return torch.cat(parts)

    def _post(self, y: torch.Tensor) -> torch.Tensor:
        name = self.output_activation_name
        if name is None or str(name).lower() == "none":
            return y
        from .utils import get_activation
        return get_activation(name)(y)

    def run(self, in0: Optional[torch.Tensor], in1: Optional[torch.Tensor], in0_scale=1.0, in0_offset=0.0,
            n_out_used: Optional[int] = None, flat: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Network on cat[in0*scale+offset, in1] without materialising the concatenation."""
        n0 = 0 if in0 is None else in0.shape[-1]
        n1 = 0 if in1 is None else in1.shape[-1]
        assert n0 + n1 == self.n_input_dims, f"MLP expects {self.n_input_dims} inputs, got {n0}+{n1}"
        desc = ops.make_mlp_desc(n0, n1, self.n_hidden_layers, self.n_output_dims, self.hidden_act, in0_scale, in0_offset,
                                 self.precision)
        if flat is None:
            flat = self.flat_params()
        nou = self.n_output_dims if n_out_used is None else n_out_used
        return self._post(ops.mlp_apply(in0, in1, flat, desc, nou))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        return self.run(None, x.float())

    def forward_twice_differentiable(self, x: torch.Tensor) -> torch.Tensor:
        """The same network as plain torch operators (exactly reference models/network_utils.py:108-113), for the one
        place that differentiates THROUGH a derivative of the network: grad_type 'analytic' (models/geometry.py:214-218).
        The fused kernels implement first-order adjoints only; the second-order MLP adjoint is the next step of
        SURVEY.md section 8f rank 1 (the hash-grid second-order adjoints are kernels already)."""
        if not x.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        h = x.float()
        lins = self.linears()
        for i, lin in enumerate(lins):
            h = torch.nn.functional.linear(h, lin.effective_weight(), lin.bias)
            if i < len(lins) - 1:
                h = torch.nn.functional.softplus(h, beta=100) if self.sphere_init else torch.relu(h)
        return self._post(h)


def get_mlp(n_input_dims: int, n_output_dims: int, config) -> VanillaMLP:
    """reference models/network_utils.py:177-185.  otype VanillaMLP -> fp32 arithmetic; the tcnn otypes
    (FullyFusedMLP, CutlassMLP) -> tensor-core arithmetic with fp32 accumulate."""
    otype = config.get("otype", "VanillaMLP")
    if otype == "VanillaMLP":
        return VanillaMLP(n_input_dims, n_output_dims, config, L.IA_MLP_FP32)
    if otype in ("FullyFusedMLP", "CutlassMLP"):
        return VanillaMLP(n_input_dims, n_output_dims, config, L.IA_MLP_TC_F16)
    raise NotImplementedError(f"network otype {otype}")


class EncodingWithNetwork(nn.Module):
    """reference models/network_utils.py:188-198, fused: hash grid -> MLP with the xyz pass-through handled
    inside the MLP operator."""

    def __init__(self, encoding: CompositeEncoding, network: VanillaMLP):
        super().__init__()
        self.encoding, self.network = encoding, network

    def forward(self, x: torch.Tensor, n_out_used: Optional[int] = None, flat: Optional[torch.Tensor] = None, group: int = 1) -> torch.Tensor:
        return fused_encode_mlp(self.encoding, self.network, x, n_out_used, flat, group)

    def update_step(self, epoch, global_step):
        update_module_step(self.encoding, epoch, global_step)
        update_module_step(self.network, epoch, global_step)


def fused_encode_mlp(encoding: CompositeEncoding, network: VanillaMLP, x: torch.Tensor, n_out_used: Optional[int] = None,
                     flat: Optional[torch.Tensor] = None, group: int = 1) -> torch.Tensor:
    """network(encoding(x)) for a CompositeEncoding without the torch.cat of reference
    models/network_utils.py:77: the include_xyz columns are produced inside the MLP kernel."""
    x = x.reshape(-1, encoding.n_input_dims)
    inner = encoding.encoding
    # fused path: the gather runs inside the tensor-core MLP kernel, the [N, L*F] encoding never exists (ops.sdf_fused)
    grid, active = None, None
    if isinstance(inner, ProgressiveBandHashGrid):
        grid, active = inner.encoding, inner.active_levels
    elif isinstance(inner, Encoding) and inner.otype == "HashGrid":
        grid, active = inner, inner.plan.n_levels
    if grid is not None and ops.sdf_fused_enabled() and encoding.include_xyz and x.is_cuda and \
            grid.table_precision == "fp32" and network.output_activation_name in (None, "none", "None"):
        desc = ops.make_mlp_desc(3, grid.n_output_dims, network.n_hidden_layers, network.n_output_dims, network.hidden_act,
                                 encoding.xyz_scale, encoding.xyz_offset, network.precision)
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or grid.params.requires_grad)
        if ops.sdf_fused_supported(desc, grid.plan, needs_grad):
            if flat is None:
                flat = network.flat_params()
            return ops.sdf_fused(x, grid.params, flat, desc, grid.plan, active, n_out_used, group)
    enc = inner(x, group=group) if group > 1 and isinstance(inner, (ProgressiveBandHashGrid, Encoding)) and \
        getattr(inner, "otype", "HashGrid") == "HashGrid" else inner(x)
    if encoding.include_xyz:
        return network.run(x, enc, encoding.xyz_scale, encoding.xyz_offset, n_out_used, flat)
    return network.run(None, enc, 1.0, 0.0, n_out_used, flat)


def get_encoding_with_network(n_input_dims, n_output_dims, encoding_config, network_config) -> EncodingWithNetwork:
    """reference models/network_utils.py:201-216."""
    encoding = get_encoding(n_input_dims, encoding_config)
    network = get_mlp(encoding.n_output_dims, n_output_dims, network_config)
    return EncodingWithNetwork(encoding, network)
