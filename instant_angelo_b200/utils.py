"""Small helpers of reference models/utils.py:54-114 that the hot path needs."""
from __future__ import annotations

import torch
import torch.nn.functional as F


class _TruncExp(torch.autograd.Function):
    """reference models/utils.py:54-69: exp forward, gradient clamps the exponent at 15."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(torch.clamp(x, max=15))


trunc_exp = _TruncExp.apply


def get_activation(name):
    """reference models/utils.py:72-98."""
    if name is None:
        return lambda x: x
    name = str(name).lower()
    if name == "none":
        return lambda x: x
    if name.startswith("scale"):
        s = float(name[5:])
        return lambda x: x.clamp(0.0, s) / s
    if name.startswith("clamp"):
        s = float(name[5:])
        return lambda x: x.clamp(0.0, s)
    if name.startswith("mul"):
        s = float(name[3:])
        return lambda x: x * s
    if name == "lin2srgb":
        return lambda x: torch.where(x > 0.0031308, torch.pow(torch.clamp(x, min=0.0031308), 1.0 / 2.4) * 1.055 - 0.055,
                                     12.92 * x).clamp(0.0, 1.0)
    if name == "trunc_exp":
        return trunc_exp
    if name.startswith("+") or name.startswith("-"):
        s = float(name)
        return lambda x: x + s
    if name == "sigmoid":
        return torch.sigmoid
    if name == "tanh":
        return torch.tanh
    return getattr(F, name)


def scale_anything(dat, inp_scale, tgt_scale):
    """reference models/utils.py:109-114."""
    if inp_scale is None:
        inp_scale = [dat.min(), dat.max()]
    dat = (dat - inp_scale[0]) / (inp_scale[1] - inp_scale[0])
    dat = dat * (tgt_scale[1] - tgt_scale[0]) + tgt_scale[0]
    return dat
