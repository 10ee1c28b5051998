"""Loss terms of reference systems/neus.py:130-194 and the scalar schedules C() of systems/base.py:28-45."""
from __future__ import annotations

import os
from typing import Dict

import torch
import torch.nn.functional as F

from . import ops


def C(value, global_step: int, current_epoch: int = 0) -> float:
    """reference systems/base.py:28-45."""
    if isinstance(value, (int, float)):
        return value
    value = list(value)
    if len(value) == 3:
        value = [0] + value
    assert len(value) == 4
    start_step, start_value, end_value, end_step = value
    cur = global_step if isinstance(end_step, int) else current_epoch
    return start_value + (end_value - start_value) * max(min(1.0, (cur - start_step) / (end_step - start_step)), 0.0)


def binary_cross_entropy(inp, target):
    """reference systems/criterions.py:155-159."""
    return -(target * torch.log(inp) + (1 - target) * torch.log(1 - inp)).mean()


def flatten_eff_distloss(w: torch.Tensor, m: torch.Tensor, interval: torch.Tensor, ray_id: torch.Tensor) -> torch.Tensor:
    """Mip-NeRF 360 distortion loss over packed samples, the O(S) prefix-sum form of `torch_efficient_distloss`
    (an un-vendored pip dependency of the reference, imported at systems/neus.py:15, used at :163-171):

        loss = ( sum_i w_i^2 interval_i / 3  +  2 sum_i w_i (m_i W_i - WM_i) ) / n_rays,
        W_i / WM_i = sums of w_j / w_j m_j over the samples j < i of the same ray, n_rays = ray_id.max() + 1.

    Samples are sorted by ray (nerfacc's packed order).  Every shipped config has lambda_distortion = 0, so this is off
    the hot path: plain tensor operations, float64 prefix sums, gradients by autograd."""
    if w.numel() == 0:
        return w.sum() * 0.0
    w, m, interval = w.reshape(-1), m.reshape(-1), interval.reshape(-1)
    ray_id = ray_id.reshape(-1)
    n_rays = (ray_id.max() + 1).to(w.dtype)
    wd, md = w.double(), m.double()
    wm = wd * md
    w_ex, wm_ex = torch.cumsum(wd, 0) - wd, torch.cumsum(wm, 0) - wm          # exclusive prefix over the whole batch
    first = torch.ones_like(ray_id, dtype=torch.bool)
    first[1:] = ray_id[1:] != ray_id[:-1]
    idx = torch.arange(ray_id.numel(), device=ray_id.device)
    start = torch.cummax(torch.where(first, idx, torch.zeros_like(idx)), 0).values   # index of the ray's first sample
    w_prefix, wm_prefix = w_ex - w_ex[start], wm_ex - wm_ex[start]
    loss_uni = (interval.double() * wd * wd).sum() / 3.0
    loss_bi = 2.0 * (wd * (md * w_prefix - wm_prefix)).sum()
    return ((loss_uni + loss_bi) / n_rays.double()).to(w.dtype)


def training_loss(model, out: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], loss_cfg, global_step: int,
                  has_mask: bool = False, current_epoch: int = 0) -> Dict[str, torch.Tensor]:
    """reference systems/neus.py:130-194.  Returns every term plus 'loss'."""
    c = lambda v: C(v, global_step, current_epoch)
    terms = {}
    valid = out["rays_valid_full"][..., 0]
    use_mask = has_mask and "fg_mask" in batch
    if out["comp_rgb_full"].is_cuda and os.environ.get("IA_NO_FUSED_LOSSES") is None:
        # the per-ray / per-sample terms and their weighted sum as one kernel forward and one backward (ops.neus_losses):
        # as tensor expressions they are ~120 launches per step
        want_curv = c(loss_cfg["lambda_curvature"]) > 0
        if want_curv:
            assert "sdf_laplace_samples" in out, "Need geometry.grad_type='finite_difference' to get SDF Laplace samples"
        lambdas = {"rgb_mse": c(loss_cfg["lambda_rgb_mse"]), "rgb_l1": c(loss_cfg["lambda_rgb_l1"]),
                   "eikonal": c(loss_cfg["lambda_eikonal"]), "mask": c(loss_cfg["lambda_mask"]) if use_mask else 0.0,
                   "opaque": c(loss_cfg["lambda_opaque"]), "sparsity": c(loss_cfg["lambda_sparsity"]),
                   "curvature": c(loss_cfg["lambda_curvature"])}
        loss, named = ops.neus_losses(out["comp_rgb_full"], batch["rgb"], valid, out["opacity"],
                                      batch["fg_mask"].float() if use_mask else None, out["sdf_grad_samples"], out["sdf_samples"],
                                      out["sdf_laplace_samples"] if want_curv else None, lambdas, float(loss_cfg["sparsity_scale"]))
        terms.update(named)
    else:
        # mean over the valid rays' channels, as F.mse_loss / F.l1_loss on comp_rgb_full[valid] (systems/neus.py:134-138),
        # written as a masked sum so that no boolean compaction (a host read-back per use) sits inside the step;
        # with no valid ray both forms give 0/0 = nan
        diff = torch.where(valid[:, None], out["comp_rgb_full"] - batch["rgb"], 0.0)
        n_valid = valid.sum().to(diff.dtype) * diff.shape[-1]
        terms["rgb_mse"] = (diff * diff).sum() / n_valid
        loss = terms["rgb_mse"] * c(loss_cfg["lambda_rgb_mse"])
        terms["rgb_l1"] = diff.abs().sum() / n_valid
        loss = loss + terms["rgb_l1"] * c(loss_cfg["lambda_rgb_l1"])
        terms["eikonal"] = ((torch.linalg.norm(out["sdf_grad_samples"], ord=2, dim=-1) - 1.0) ** 2).mean()
        loss = loss + terms["eikonal"] * c(loss_cfg["lambda_eikonal"])
        opacity = torch.clamp(out["opacity"].squeeze(-1), 1.0e-3, 1.0 - 1.0e-3)
        if use_mask:
            terms["mask"] = binary_cross_entropy(opacity, batch["fg_mask"].float())
            loss = loss + terms["mask"] * c(loss_cfg["lambda_mask"])
        terms["opaque"] = binary_cross_entropy(opacity, opacity)
        loss = loss + terms["opaque"] * c(loss_cfg["lambda_opaque"])
        terms["sparsity"] = torch.exp(-loss_cfg["sparsity_scale"] * out["sdf_samples"].abs()).mean()
        loss = loss + terms["sparsity"] * c(loss_cfg["lambda_sparsity"])
        if c(loss_cfg["lambda_curvature"]) > 0:
            assert "sdf_laplace_samples" in out, "Need geometry.grad_type='finite_difference' to get SDF Laplace samples"
            terms["curvature"] = out["sdf_laplace_samples"].abs().mean()
            loss = loss + terms["curvature"] * c(loss_cfg["lambda_curvature"])
    # systems/neus.py:161-171 (inactive in every shipped config: lambda_distortion = lambda_distortion_bg = 0)
    if c(loss_cfg.get("lambda_distortion", 0.0)) > 0:
        terms["distortion"] = flatten_eff_distloss(out["weights"], out["points"], out["intervals"], out["ray_indices"])
        loss = loss + terms["distortion"] * c(loss_cfg["lambda_distortion"])
    if getattr(model, "learned_background", False) and c(loss_cfg.get("lambda_distortion_bg", 0.0)) > 0:
        terms["distortion_bg"] = flatten_eff_distloss(out["weights_bg"], out["points_bg"], out["intervals_bg"], out["ray_indices_bg"])
        loss = loss + terms["distortion_bg"] * c(loss_cfg["lambda_distortion_bg"])
    if c(loss_cfg["lambda_sdf_l1"]) > 0 and batch.get("pts") is not None:
        sdf_p, grad_p = model.geometry(batch["pts"], with_grad=True, with_feature=False)
        if grad_p.is_cuda and os.environ.get("IA_NO_FUSED_LOSSES") is None and sdf_p.numel() > 0:
            part, named = ops.point_losses(sdf_p, grad_p, batch["pts_normal"], batch["pts_weights"], c(loss_cfg["lambda_sdf_l1"]),
                                           c(loss_cfg.get("lambda_normal", loss_cfg["lambda_sdf_l1"])))
            terms.update(named)
            terms["loss"] = loss + part
            return terms
        terms["sdf_l1"] = (F.l1_loss(sdf_p, torch.zeros_like(sdf_p)) * batch["pts_weights"]).mean(dim=0)   # Appendix C-11
        if grad_p.is_cuda:
            n_gt, n_pr = ops.normalize3(batch["pts_normal"]), ops.normalize3(grad_p)
        else:
            n_gt, n_pr = F.normalize(batch["pts_normal"], p=2, dim=-1), F.normalize(grad_p, p=2, dim=-1)
        terms["normal_cos"] = (1.0 - torch.sum(n_pr * n_gt, dim=-1)).mean()
        loss = loss + terms["sdf_l1"] * c(loss_cfg["lambda_sdf_l1"])
        loss = loss + terms["normal_cos"] * c(loss_cfg.get("lambda_normal", loss_cfg["lambda_sdf_l1"]))
    terms["loss"] = loss
    return terms
