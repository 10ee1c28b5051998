"""ORACLE (test infrastructure only) -- CPU restatement of the tiny-cuda-nn pieces the
Instant-angelo hot path calls.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (instant_angelo_b200/) never does.

What is restated, and where the reference calls it:
  * tcnn.Encoding(otype=HashGrid)            <- models/network_utils.py:40-59 (ProgressiveBandHashGrid)
  * tcnn.Encoding(otype=SphericalHarmonics)  <- models/network_utils.py:83-93 via models/texture.py:15-25

tiny-cuda-nn is an un-vendored, unpinned (git master, README.md:25) dependency that is absent from
/root/reference and not installable here, so the algorithm below follows its published semantics
(SURVEY.md Appendix A.1 / A.2).  PARITY UNPINNED: the reference ships no tests or golden vectors
for this boundary; the known-answer tests in tests/test_oracle_hashgrid.py are authored here.

Everything is plain PyTorch on CPU, differentiable through autograd w.r.t. the table AND the
input positions (the reference needs d enc / d x because curvature tap positions depend on
parameters, models/geometry.py:238-246).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

PRIME_Y = 2654435761
PRIME_Z = 805459861


def _f32(v) -> np.float32:
    return np.float32(v)


@dataclass
class GridPlan:
    """Per-level geometry of a tcnn HashGrid (float32 arithmetic)."""
    n_levels: int
    n_features: int
    log2_hashmap_size: int
    base_resolution: int
    per_level_scale: float
    scale: List[float] = field(default_factory=list)      # float32 values
    res: List[int] = field(default_factory=list)
    size: List[int] = field(default_factory=list)          # entries in the level
    offset: List[int] = field(default_factory=list)        # entry offsets, len L+1
    hashed: List[bool] = field(default_factory=list)

    @property
    def n_entries(self) -> int:
        return self.offset[-1]

    @property
    def n_params(self) -> int:
        return self.offset[-1] * self.n_features

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_features


def grid_plan(n_levels: int, n_features: int, log2_hashmap_size: int, base_resolution: int,
              per_level_scale: float) -> GridPlan:
    """tcnn GridEncoding constructor arithmetic.

    scale_l = exp2f(l * log2f(pls)) * base - 1   (every step rounded to float32)
    res_l   = ceilf(scale_l) + 1
    n_l     = min(next_multiple(res_l^3 saturating at 2^31-ish, 8), 2^log2T)
    Correctly-rounded float32 functions are emulated by evaluating in float64 and rounding.
    """
    plan = GridPlan(n_levels, n_features, log2_hashmap_size, base_resolution, per_level_scale)
    pls32 = _f32(per_level_scale)
    log2_pls = _f32(math.log2(float(pls32)))
    offset = 0
    plan.offset.append(0)
    for l in range(n_levels):
        arg = _f32(_f32(l) * log2_pls)
        e = _f32(2.0 ** float(arg))
        scale = _f32(_f32(e * _f32(base_resolution)) - _f32(1.0))
        res = int(math.ceil(float(scale))) + 1
        max_params = (2 ** 32 - 1) // 2
        dense = res ** 3
        n = max_params if float(res) ** 3 > float(max_params) else dense
        n = ((n + 7) // 8) * 8
        n = min(n, 1 << log2_hashmap_size)
        # a level is addressed by the hash iff the dense stride overflows its entry count
        stride = 1
        for _ in range(3):
            if stride > n:
                break
            stride *= res
        plan.scale.append(float(scale))
        plan.res.append(res)
        plan.size.append(n)
        plan.hashed.append(n < stride)
        offset += n
        plan.offset.append(offset)
    return plan


def _corner_index(plan: GridPlan, level: int, cx: torch.Tensor, cy: torch.Tensor, cz: torch.Tensor) -> torch.Tensor:
    """tcnn grid_index(): uint32 wrap-around arithmetic done in int64 with explicit masks."""
    M = 0xFFFFFFFF
    res, n = plan.res[level], plan.size[level]
    cx, cy, cz = cx & M, cy & M, cz & M
    coords = (cx, cy, cz)
    stride, idx = 1, torch.zeros_like(cx)
    for d in range(3):
        if stride > n:
            break
        idx = (idx + coords[d] * stride) & M
        stride = (stride * res) & M
    if plan.hashed[level]:
        idx = (cx * 1) ^ ((cy * PRIME_Y) & M) ^ ((cz * PRIME_Z) & M)
    return idx % n


def hashgrid_forward(x: torch.Tensor, table: torch.Tensor, plan: GridPlan, active_levels: int | None = None) -> torch.Tensor:
    """x: [N,3] fp32 in [0,1]; table: flat fp32 [n_params] (level-major, entry-major, feature-minor).

    Returns [N, L*F] fp32, level-major feature order.  Levels >= active_levels are returned as exact
    zeros, which is bit-identical to the reference's multiply by the 0/1 progressive mask
    (models/network_utils.py:56-59) for finite table values.
    """
    assert x.dim() == 2 and x.shape[1] == 3
    L, F = plan.n_levels, plan.n_features
    if active_levels is None:
        active_levels = L
    tab = table.view(-1, F)
    outs = []
    x64 = x.double()
    for l in range(L):
        if l >= active_levels:
            outs.append(torch.zeros(x.shape[0], F, dtype=x.dtype))
            continue
        # pos = fmaf(scale, x, 0.5f): single rounding, emulated through float64
        pos = x64 * float(plan.scale[l]) + 0.5
        if x.dtype != torch.float64:          # float64 inputs: the high-precision arbiter of the parity tests, no fp32 rounding
            pos = pos.float()
        g = torch.floor(pos.detach())
        w = pos - g                                   # d w / d x = scale_l
        gi = g.long()
        # all 8 corners of a level are gathered at once from the level's own slice of the table, so that
        # autograd builds one level-sized gradient per level (not one table-sized gradient per corner)
        tl = tab[plan.offset[l]: plan.offset[l + 1]]
        idx8, w8 = [], []
        for corner in range(8):
            wgt = torch.ones(x.shape[0], dtype=x.dtype)
            c = []
            for d in range(3):
                if corner & (1 << d):
                    wgt = wgt * w[:, d]
                    c.append(gi[:, d] + 1)
                else:
                    wgt = wgt * (1.0 - w[:, d])
                    c.append(gi[:, d])
            idx8.append(_corner_index(plan, l, c[0], c[1], c[2]))
            w8.append(wgt)
        idx8 = torch.stack(idx8, dim=1)                  # [N,8]
        w8 = torch.stack(w8, dim=1)                      # [N,8]
        acc = (w8[:, :, None] * tl[idx8]).sum(dim=1)     # [N,F]
        outs.append(acc)
    return torch.cat(outs, dim=1)


def corner_indices(x: torch.Tensor, plan: GridPlan, level: int) -> torch.Tensor:
    """Debug/KAT helper: the 8 table entry indices (level-local) for each point. [N,8] int64."""
    pos = (x.double() * float(plan.scale[level]) + 0.5).float()
    gi = torch.floor(pos).long()
    cols = []
    for corner in range(8):
        c = [gi[:, d] + (1 if corner & (1 << d) else 0) for d in range(3)]
        cols.append(_corner_index(plan, level, c[0], c[1], c[2]))
    return torch.stack(cols, dim=1)


# ---------------------------------------------------------------------------------------------
# Spherical harmonics (tcnn SphericalHarmonics encoding)
# ---------------------------------------------------------------------------------------------

def sh_forward(d01: torch.Tensor, degree: int) -> torch.Tensor:
    """d01: [N,3] in [0,1] (the reference maps unit dirs with (d+1)/2, models/texture.py:24).
    Returns [N, degree^2] fp32 real SH of (x,y,z)=2*d01-1."""
    assert 1 <= degree <= 4
    v = d01 * 2.0 - 1.0
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    out = [torch.full_like(x, 0.28209479177387814)]
    if degree > 1:
        out += [-0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x]
    if degree > 2:
        out += [1.0925484305920792 * xy, -1.0925484305920792 * yz,
                0.94617469575755997 * z2 - 0.31539156525251999,
                -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2]
    if degree > 3:
        out += [0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
                0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
                0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
                0.59004358992664352 * x * (-x2 + 3.0 * y2)]
    return torch.stack(out, dim=1)
