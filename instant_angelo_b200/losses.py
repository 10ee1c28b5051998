"""Loss terms of reference systems/neus.py:130-194 and the scalar schedules C() of systems/base.py:28-45."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


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


def training_loss(model, out: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], loss_cfg, global_step: int,
                  has_mask: bool = False) -> Dict[str, torch.Tensor]:
    """reference systems/neus.py:130-194.  Returns every term plus 'loss'."""
    c = lambda v: C(v, global_step)
    terms = {}
    # mean over the valid rays' channels, as F.mse_loss / F.l1_loss on comp_rgb_full[valid] (systems/neus.py:134-138),
    # written as a masked sum so that no boolean compaction (a host read-back per use) sits inside the step;
    # with no valid ray both forms give 0/0 = nan
    valid = out["rays_valid_full"][..., 0]
    diff = torch.where(valid[:, None], out["comp_rgb_full"] - batch["rgb"], 0.0)
    n_valid = valid.sum().to(diff.dtype) * diff.shape[-1]
    terms["rgb_mse"] = (diff * diff).sum() / n_valid
    loss = terms["rgb_mse"] * c(loss_cfg["lambda_rgb_mse"])
    terms["rgb_l1"] = diff.abs().sum() / n_valid
    loss = loss + terms["rgb_l1"] * c(loss_cfg["lambda_rgb_l1"])
    terms["eikonal"] = ((torch.linalg.norm(out["sdf_grad_samples"], ord=2, dim=-1) - 1.0) ** 2).mean()
    loss = loss + terms["eikonal"] * c(loss_cfg["lambda_eikonal"])
    opacity = torch.clamp(out["opacity"].squeeze(-1), 1.0e-3, 1.0 - 1.0e-3)
    if has_mask and "fg_mask" in batch:
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
    if c(loss_cfg["lambda_sdf_l1"]) > 0 and batch.get("pts") is not None:
        sdf_p, grad_p = model.geometry(batch["pts"], with_grad=True, with_feature=False)
        terms["sdf_l1"] = (F.l1_loss(sdf_p, torch.zeros_like(sdf_p)) * batch["pts_weights"]).mean(dim=0)   # Appendix C-11
        n_gt = F.normalize(batch["pts_normal"], p=2, dim=-1)
        n_pr = F.normalize(grad_p, p=2, dim=-1)
        terms["normal_cos"] = (1.0 - torch.sum(n_pr * n_gt, dim=-1)).mean()
        loss = loss + terms["sdf_l1"] * c(loss_cfg["lambda_sdf_l1"])
        loss = loss + terms["normal_cos"] * c(loss_cfg.get("lambda_normal", loss_cfg["lambda_sdf_l1"]))
    terms["loss"] = loss
    return terms
