"""Generates tests/golden/*.npz by running the REFERENCE's own Python (models/network_utils.py,
models/geometry.py, models/neus.py, models/texture.py, systems/neus.py under /root/reference) on CPU.

The reference imports tinycudann, nerfacc, pytorch_lightning, omegaconf, trimesh, imageio, matplotlib
and torch_efficient_distloss, none of which exist in this image.  They are replaced by stubs:
  * tinycudann.Encoding, nerfacc.*  -> the third-party restatements in oracle/tcnn_ref.py / oracle/nerfacc_ref.py
    (the only way to execute the reference here; PARITY UNPINNED at that boundary, see oracle/__init__.py);
  * everything else                 -> inert shims (logging no-ops, OmegaConf.to_container, LightningModule = object).
So the fixtures pin everything the reference itself implements: VanillaMLP + weight-norm + sphere init,
ProgressiveBandHashGrid masking, CompositeEncoding, VolumeSDF finite-difference gradient and curvature
(with its quirks), VolumeDensity, the colour heads, get_alpha, forward_/forward_bg_ and the loss terms of
training_step.  Runs only in the build container (needs /root/reference); the .npz files are committed.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import nerfacc_ref as nf  # noqa: E402
from oracle import tcnn_ref as tc  # noqa: E402
from oracle import model_ref as mr  # noqa: E402
from instant_angelo_b200.config import Config, to_config, to_primitive  # noqa: E402
from tests.golden.scenes import golden_model_config, golden_loss_config, sphere_shell_binary, make_rays  # noqa: E402


# ---------------------------------------------------------------------------------------------
# stubs for the packages the reference imports
# ---------------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _StubTcnnEncoding(torch.nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        self.inner = mr.RefTcnnEncoding(n_input_dims, dict(encoding_config), seed)
        self.params = self.inner.params          # same key as tcnn: <...>.params
        del self.inner._parameters["params"]
        self.inner.params = self.params
        self.n_input_dims = n_input_dims
        self.n_output_dims = self.inner.n_output_dims

    def forward(self, x):
        return self.inner(x)


class _StubOccupancyGrid(torch.nn.Module):
    def __init__(self, roi_aabb, resolution=128, contraction_type=nf.ContractionType.AABB):
        super().__init__()
        self.inner = nf.OccupancyGrid(roi_aabb, resolution, contraction_type)

    @property
    def binary(self):
        return self.inner.binary

    @property
    def roi_aabb(self):
        return self.inner.roi_aabb

    @property
    def contraction_type(self):
        return self.inner.contraction_type

    def every_n_step(self, step, occ_eval_fn, occ_thre=1e-2, ema_decay=0.95, warmup_steps=256, n=16):
        raise RuntimeError("golden runs set the grids by hand")


_RNG = {}


def _stub_ray_marching(rays_o, rays_d, **kw):
    grid = kw.pop("grid", None)
    key = "u_bg" if kw.get("cone_angle", 0.0) > 0 else "u_fg"
    return nf.ray_marching(rays_o, rays_d, grid=grid.inner if grid is not None else None, stratified_u=_RNG[key], **kw)


def install_stubs():
    _mod("tinycudann", Encoding=_StubTcnnEncoding, Network=None, NetworkWithInputEncoding=None,
         free_temporary_memory=lambda: None)
    _mod("nerfacc", ContractionType=nf.ContractionType, OccupancyGrid=_StubOccupancyGrid, ray_marching=_stub_ray_marching,
         render_weight_from_density=nf.render_weight_from_density, render_weight_from_alpha=nf.render_weight_from_alpha,
         accumulate_along_rays=nf.accumulate_along_rays)
    _mod("nerfacc.intersection", ray_aabb_intersect=lambda o, d, aabb: nf.ray_aabb_intersect(o, d, aabb))
    rz = _mod("pytorch_lightning.utilities.rank_zero", rank_zero_info=lambda *a, **k: None,
              rank_zero_debug=lambda *a, **k: None, rank_zero_warn=lambda *a, **k: None)
    ut = _mod("pytorch_lightning.utilities", rank_zero=rz)
    _mod("pytorch_lightning", utilities=ut, LightningModule=type("LightningModule", (), {"__init__": lambda self: None}))

    class OmegaConf:
        @staticmethod
        def register_new_resolver(*a, **k):
            pass

        @staticmethod
        def to_container(cfg, resolve=True):
            return to_primitive(cfg)

    _mod("omegaconf", OmegaConf=OmegaConf)
    _mod("trimesh")
    _mod("imageio")
    _mod("matplotlib", cm=None)
    _mod("matplotlib.cm")
    _mod("matplotlib.colors", LinearSegmentedColormap=None)
    _mod("torch_efficient_distloss", flatten_eff_distloss=None)
    # `systems` as a namespace so that systems/__init__.py (which pulls Lightning-bound code) is not executed
    pkg = types.ModuleType("systems")
    pkg.__path__ = [os.path.join(REF, "systems")]
    pkg.register = lambda name: (lambda cls: cls)
    sys.modules["systems"] = pkg
    sys.path.insert(0, REF)
    torch.cuda.device = lambda *a, **k: contextlib.nullcontext()
    misc = importlib.import_module("utils.misc")
    misc.get_rank = lambda: "cpu"


def load_reference():
    install_stubs()
    models = importlib.import_module("models")
    neus_sys = importlib.import_module("systems.neus")
    return models, neus_sys


# ---------------------------------------------------------------------------------------------
def perturb_(model: torch.nn.Module, seed: int):
    """'trained-like' weights: tables N(0, 0.05), every MLP weight + N(0, 0.05) so that hash features matter."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".params") and p.numel() > 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
            elif "weight_v" in name or name.endswith(".weight"):
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
            elif name.endswith(".bias"):
                p.add_(torch.randn(p.shape, generator=g) * 0.02)


class _FakeSystem:
    """Just enough of NeuSSystem for the reference's training_step (systems/neus.py:120-206) to run."""

    def __init__(self, neus_sys, model, config, global_step):
        self._cls = neus_sys.NeuSSystem
        self.model, self.config = model, config
        self.global_step, self.current_epoch = global_step, 0
        self.train_num_rays = config.model.train_num_rays
        self.train_num_samples = config.model.train_num_rays * (config.model.num_samples_per_ray + config.model.get("num_samples_per_ray_bg", 0))
        self.dataset = types.SimpleNamespace(has_mask=False)
        self.logged = {}

    def __call__(self, batch):
        self.out = self.model(batch["rays"])
        return self.out

    def log(self, name, value, **kw):
        self.logged[name] = value

    def C(self, value):
        from systems.base import BaseSystem
        return BaseSystem.C(self, value)

    def training_step(self, batch):
        return self._cls.training_step(self, batch, 0)


def run_case(models, neus_sys, name: str, texture: str, learned_background: bool, global_step: int, seed: int,
             grad_type: str = "finite_difference"):
    torch.manual_seed(seed)
    mcfg = to_config(golden_model_config(texture=texture, learned_background=learned_background, grad_type=grad_type))
    cfg = to_config({"model": mcfg, "system": {"loss": golden_loss_config()}})
    cfg.model.dynamic_ray_sampling = False
    model = models.make("neus", cfg.model)
    perturb_(model, seed + 1)
    model.train()
    model.occupancy_grid.inner.binary = sphere_shell_binary(128, cfg.model.radius)
    if learned_background:
        model.occupancy_grid_bg.inner.binary = torch.ones(256, 256, 256, dtype=torch.bool)
    from systems.utils import update_module_step
    # NeuSModel.update_step minus the occupancy refresh (grids are set by hand)
    model.config.grid_prune = False
    model.update_step(0, global_step)
    model.config.grid_prune = True

    g = torch.Generator().manual_seed(seed + 2)
    n_rays = 48
    rays, rgb = make_rays(n_rays, g)
    _RNG["u_fg"] = torch.rand(n_rays, generator=g)
    _RNG["u_bg"] = torch.rand(n_rays, generator=g)
    model.background_color = torch.rand(3, generator=g)
    # the reference draws randn_like(points) inside VolumeSDF.forward (geometry.py:238); capture it
    drawn = {}
    orig_randn_like = torch.randn_like

    def randn_like(t, *a, **k):
        r = orig_randn_like(t, *a, **k)
        drawn.setdefault("rand_directions", r.clone())
        return r

    torch.randn_like = randn_like
    pts = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2
    batch = {"rays": rays, "rgb": rgb, "fg_mask": torch.ones(n_rays), "pts": pts,
             "pts_normal": torch.nn.functional.normalize(pts, dim=-1), "pts_weights": torch.rand(n_rays, generator=g)}
    system = _FakeSystem(neus_sys, model, cfg, global_step)
    try:
        loss = system.training_step(batch)["loss"]
    finally:
        torch.randn_like = orig_randn_like
    loss.backward()
    out = system.out

    fx = {"global_step": np.int64(global_step), "rays": rays.numpy(), "rgb": rgb.numpy(), "u_fg": _RNG["u_fg"].numpy(),
          "u_bg": _RNG["u_bg"].numpy(), "background_color": model.background_color.numpy(),
          "rand_directions": drawn["rand_directions"].numpy(), "pts": pts.detach().numpy(),
          "pts_normal": batch["pts_normal"].numpy(), "pts_weights": batch["pts_weights"].numpy(),
          "loss": loss.detach().numpy()}
    for k, v in out.items():
        if isinstance(v, torch.Tensor):
            fx["out." + k] = v.detach().numpy()
    for k, v in system.logged.items():
        if isinstance(v, torch.Tensor):
            fx["log." + k] = v.detach().numpy()
    for k, p in model.named_parameters():
        fx["param." + k] = p.detach().numpy()
        if p.grad is not None:
            fx["grad." + k] = p.grad.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **fx)
    n_s = int(out["num_samples"].item())
    print(f"{name}: loss={float(loss):.6f} samples fg={n_s} full={int(out['num_samples_full'].item())} -> "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def main():
    models, neus_sys = load_reference()
    run_case(models, neus_sys, "neus_dualcolor_bg", "volume-dual-color", True, 25, 100)
    run_case(models, neus_sys, "neus_v3_nobg", "volume-dual-colorV3", False, 60, 200)
    run_case(models, neus_sys, "neus_analytic_bg", "volume-dual-color", True, 40, 300, grad_type="analytic")


if __name__ == "__main__":
    main()
