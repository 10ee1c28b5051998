"""Colour heads with the reference's surface (models/texture.py:10-64, 113-149) on the fused kernels.

The concatenations of the reference (`torch.cat([features, dirs_embd, normals])`) are kept -- they are one
[S, <=87] tensor per step against 13 geometry evaluations -- but the SH encoding and the MLPs run in
libia_b200.so."""
from __future__ import annotations

import torch
import torch.nn as nn

import os

from . import _lib as _L
from . import ops
from . import registry as models
from .network_utils import get_encoding, get_mlp, update_module_step
from .utils import get_activation


# IA_FOLD_HEAD=0: keep the geometry output layer as its own (fp32 FFMA) projection in front of the colour network (A/B runs)
_FOLD_HEAD = os.environ.get("IA_FOLD_HEAD", "1") not in ("0", "")


@models.register("volume-radiance")
class VolumeRadiance(nn.Module):
    """reference models/texture.py:10-36."""
    dual = False

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.n_dir_dims = self.config.get("n_dir_dims", 3)
        self.n_output_dims = 3
        self.encoding = get_encoding(self.n_dir_dims, self.config["dir_encoding_config"])
        self.n_input_dims = self.config["input_feature_dim"] + self.encoding.n_output_dims
        self.network = get_mlp(self.n_input_dims, self.n_output_dims, self.config["mlp_network_config"])

    def forward(self, features, dirs, *args):
        dirs = (dirs + 1.0) / 2.0
        dirs_embd = self.encoding(dirs.view(-1, self.n_dir_dims))
        network_inp = torch.cat([features.view(-1, features.shape[-1]), dirs_embd] +
                                [arg.view(-1, arg.shape[-1]) for arg in args], dim=-1)
        color = self.network(network_inp).view(*features.shape[:-1], self.n_output_dims).float()
        if "color_activation" in self.config:
            act = get_activation(self.config["color_activation"])
            if self.dual:
                color = act(color) + act(features[..., 1:4])     # Appendix C-10: may exceed 1
            else:
                color = act(color)
        return color

    # ---- fused-head path: the input row [feature(Fd+3) | dirs_embd | normal] is assembled by ops.sdf_head ----------
    def supports_fused_head(self, geometry) -> bool:
        return (self.n_dir_dims == 3 and
                self.n_input_dims == geometry.n_output_dims + 3 + self.encoding.n_output_dims + 3)

    def forward_fused_head(self, h, w_last, b_last, pts01, dirs, normals):
        """forward(cat[geometry_out, pts01*2-1], dirs, normals) with geometry_out = h @ w_last.T + b_last applied inside
        the assembly.  Returns (color [S,3], sdf [S])."""
        dirs_embd = self.encoding(((dirs + 1.0) / 2.0).view(-1, self.n_dir_dims))
        net = self.network
        n_feat = w_last.shape[0]
        if _FOLD_HEAD and n_feat >= 4 and net.n_hidden_layers == 2 and net.precision == _L.IA_MLP_TC_F16 and h.shape[0] > 0:
            # Fold the geometry output layer into the colour network's first layer: with out = Wl h + bl,
            #   Wc0 [out | rest] + bc0 = (Wc0[:, :n_feat] Wl) h + Wc0[:, n_feat:] rest + (bc0 + Wc0[:, :n_feat] bl),
            # so the colour network reads h (64 columns) instead of the n_feat-wide geometry output and the 64 -> n_feat
            # projection of every sample (forward and backward) disappears; the compositions are 64 x n_feat x 64 products of
            # parameters, done once per step as element-wise kernels.
            flat = net.flat_params()
            n_in, wd = net.n_input_dims, net.n_neurons
            n_tail = n_in - n_feat                                                   # pts (3) | dirs_embd | normal (3)
            ld = (wd + n_tail + 3) // 4 * 4                                          # 16-byte aligned rows
            if wd == 64 and os.environ.get("IA_NO_FOLD_KERNEL") is None:
                flat_eff = ops.fold_head(flat, w_last, b_last, n_in, n_feat, ld)     # one launch (and one in backward)
            else:
                wc0 = flat[:wd * n_in].view(wd, n_in)
                bc0 = flat[wd * n_in:wd * n_in + wd]
                rest = flat[wd * n_in + wd:]
                head = wc0[:, :n_feat]
                m = (head[:, :, None] * w_last[None, :, :]).sum(1)                    # [wd, 64]
                b_eff = bc0 + (head * b_last[None, :]).sum(1)
                pad = wc0.new_zeros(wd, ld - wd - n_tail)
                w_eff = torch.cat([m, wc0[:, n_feat:], pad], dim=1)
                flat_eff = torch.cat([w_eff.reshape(-1), b_eff, rest])
            tin, sdf, rgb_raw = ops.colour_in(h, w_last[:4], b_last[:4], pts01.reshape(-1, 3), dirs_embd, normals.reshape(-1, 3), ld)
            desc = ops.make_mlp_desc(0, ld, net.n_hidden_layers, net.n_output_dims, net.hidden_act, 1.0, 0.0, net.precision)
            color = net._post(ops.mlp_apply(None, tin, flat_eff, desc)).view(*dirs.shape[:-1], self.n_output_dims).float()
        else:
            tin, sdf, rgb_raw = ops.sdf_head(h, w_last, b_last, pts01.reshape(-1, 3), dirs_embd, normals.reshape(-1, 3))
            color = net(tin).view(*dirs.shape[:-1], self.n_output_dims).float()
        if "color_activation" in self.config:
            act = get_activation(self.config["color_activation"])
            color = act(color) + act(rgb_raw.view(*dirs.shape[:-1], 3)) if self.dual else act(color)
        return color, sdf.view(*dirs.shape[:-1])

    def update_step(self, epoch, global_step):
        update_module_step(self.encoding, epoch, global_step)

    def regularizations(self, out):
        return {}


@models.register("volume-dual-color")
class VolumeDualColor(VolumeRadiance):
    """reference models/texture.py:38-64."""
    dual = True


@models.register("volume-dual-colorV3")
class VolumeDualColorV3(nn.Module):
    """reference models/texture.py:113-149 (UniSDF camera / reflected-direction blend)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.n_dir_dims = self.config.get("n_dir_dims", 3)
        self.n_output_dims = 3
        self.encoding = get_encoding(self.n_dir_dims, self.config["dir_encoding_config"])
        self.n_input_dims = self.config["input_feature_dim"] + self.encoding.n_output_dims
        self.cam_network = get_mlp(self.n_input_dims, self.n_output_dims, self.config["mlp_network_config"])
        self.ref_network = get_mlp(self.n_input_dims, self.n_output_dims, self.config["mlp_network_config"])
        self.weight_network = get_mlp(self.config["input_feature_dim"], 1, self.config["weitht_network_config"])

    def forward(self, features, viewdirs, normals):
        dirs = (viewdirs + 1.0) / 2.0
        dirs_embd = self.encoding(dirs.view(-1, self.n_dir_dims))
        VdotN = (-viewdirs * normals).sum(-1, keepdim=True)
        refdirs = 2 * VdotN * normals + viewdirs
        refdirs = (refdirs + 1.0) / 2.0
        refdirs_embd = self.encoding(refdirs.view(-1, self.n_dir_dims))
        network_inp = torch.cat([features.view(-1, features.shape[-1]), normals.view(-1, normals.shape[-1])], dim=-1)
        ref_weight = self.weight_network(network_inp)
        cam_color = self.cam_network(torch.cat([network_inp, dirs_embd], dim=-1)).view(*features.shape[:-1], 3).float()
        ref_color = self.ref_network(torch.cat([network_inp, refdirs_embd], dim=-1)).view(*features.shape[:-1], 3).float()
        act = get_activation(self.config["color_activation"])
        return ref_weight * act(ref_color) + (1 - ref_weight) * act(cam_color)

    def update_step(self, epoch, global_step):
        pass

    def regularizations(self, out):
        return {}
