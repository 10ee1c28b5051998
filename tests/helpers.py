"""Shared helpers for the parity tests."""
from __future__ import annotations

import os

import numpy as np
import torch

from tests.golden.scenes import golden_loss_config, golden_model_config, sphere_shell_binary

GOLDEN_CASES = {
    "neus_dualcolor_bg": dict(texture="volume-dual-color", learned_background=True),
    "neus_v3_nobg": dict(texture="volume-dual-colorV3", learned_background=False),
    # the shipped default of the reference configs: autograd normals (create_graph) + FD curvature taps
    "neus_analytic_bg": dict(texture="volume-dual-color", learned_background=True, grad_type="analytic"),
}


def load_golden(golden_dir: str, name: str) -> dict:
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_state_dict(fx: dict) -> dict:
    return {k[len("param."):]: torch.from_numpy(v.copy()) for k, v in fx.items() if k.startswith("param.")}


def golden_batch(fx: dict, device="cpu") -> dict:
    t = lambda k: torch.from_numpy(fx[k].copy()).to(device)
    return {"rays": t("rays"), "rgb": t("rgb"), "pts": t("pts"), "pts_normal": t("pts_normal"),
            "pts_weights": t("pts_weights"), "fg_mask": torch.ones(fx["rays"].shape[0], device=device)}


def assert_close(actual, expected, rtol, atol, name=""):
    a = actual.detach().cpu().double().numpy() if isinstance(actual, torch.Tensor) else np.asarray(actual, dtype=np.float64)
    e = expected.detach().cpu().double().numpy() if isinstance(expected, torch.Tensor) else np.asarray(expected, dtype=np.float64)
    assert a.shape == e.shape, f"{name}: shape {a.shape} vs {e.shape}"
    if a.size == 0:
        return
    err = np.abs(a - e)
    tol = atol + rtol * np.abs(e)
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{name}: {bad.sum()}/{a.size} elements out of tolerance (rtol={rtol}, atol={atol}); worst at {i}: "
                             f"actual={a[i]:.8g} expected={e[i]:.8g} |err|={err[i]:.3g}; max|expected|={np.abs(e).max():.3g}")


def grad_tol(expected, rtol=1e-3, floor=1e-6):
    """tolerances for gradient tensors: 1e-3 relative (north_star) to the tensor's scale."""
    e = expected.detach().cpu().double().numpy() if isinstance(expected, torch.Tensor) else np.asarray(expected, np.float64)
    scale = float(np.abs(e).max()) if e.size else 0.0
    return rtol, max(floor, rtol * scale)
