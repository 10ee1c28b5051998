"""Generates tests/golden/ref_checkpoint.ckpt.gz: a Lightning-layout `.ckpt` written by the REFERENCE's own objects.

  model            = reference models.make('neus', ...) (models/neus.py, geometry.py, texture.py, network_utils.py; tcnn and
                     nerfacc stubbed by the oracle as in make_golden.py, so the state_dict keys, the flat tcnn `params`
                     layout, torch's weight-norm parameters and nerfacc's occupancy buffers are the reference's own)
  optimizer        = reference systems/utils.py:314-326 parse_optimizer  -> a real torch.optim.AdamW over the reference's
                     parameter groups, in the order the reference's module tree yields its parameters
  scheduler        = reference systems/utils.py:329-346 parse_scheduler -> a real torch SequentialLR[LinearLR, ExponentialLR]
  three steps      = reference NeuSSystem.training_step -> backward -> optimizer.step -> scheduler.step

and then the dict pytorch_lightning 1.7 saves (pytorch_lightning itself is not installable here):
`epoch, global_step, pytorch-lightning_version, state_dict ('model.' prefix), optimizer_states, lr_schedulers`.
Also stored: the reference's parameter names per optimizer index (`_param_names`, for the test's bookkeeping only) and
the learning rates after the three steps.  Occupancy buffers compress to almost nothing, hence the gzip.

    python tests/golden/make_golden_checkpoint.py
"""
from __future__ import annotations

import gzip
import importlib
import io
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from tests.golden.make_golden import _FakeSystem, _RNG, load_reference, perturb_  # noqa: E402
from tests.golden.scenes import golden_loss_config, golden_model_config, make_rays, sphere_shell_binary  # noqa: E402
from instant_angelo_b200.config import to_config  # noqa: E402
from instant_angelo_b200 import configs  # noqa: E402

N_STEPS = 3


def checkpoint_config():
    """Golden (small-table) model with the shipped optimizer / scheduler sections of neuralangelo-colmap_sparse.yaml."""
    shipped = configs.neuralangelo_colmap_sparse()
    mcfg = golden_model_config(texture="volume-dual-color", learned_background=True)
    return to_config({"model": mcfg, "system": {"loss": golden_loss_config(), "optimizer": shipped.system.optimizer,
                                                "scheduler": shipped.system.scheduler}})


def main():
    models, neus_sys = load_reference()
    sys_utils = importlib.import_module("systems.utils")
    torch.manual_seed(7)
    cfg = checkpoint_config()
    cfg.model.dynamic_ray_sampling = False
    model = models.make("neus", cfg.model)
    perturb_(model, 8)
    model.train()
    model.occupancy_grid.inner.binary = sphere_shell_binary(128, cfg.model.radius)
    model.occupancy_grid_bg.inner.binary = torch.ones(256, 256, 256, dtype=torch.bool)
    optim = sys_utils.parse_optimizer(cfg.system.optimizer, model)
    sched = sys_utils.parse_scheduler(cfg.system.scheduler, optim)["scheduler"]
    names = {id(p): n for n, p in model.named_parameters()}
    param_names = [names[id(p)] for g in optim.param_groups for p in g["params"]]
    g = torch.Generator().manual_seed(9)
    system = _FakeSystem(neus_sys, model, cfg, 0)
    for step in range(N_STEPS):
        system.global_step = step
        model.config.grid_prune = False
        model.update_step(0, step)
        model.config.grid_prune = True
        n_rays = 24
        rays, rgb = make_rays(n_rays, g)
        _RNG["u_fg"], _RNG["u_bg"] = torch.rand(n_rays, generator=g), torch.rand(n_rays, generator=g)
        model.background_color = torch.rand(3, generator=g)
        pts = (torch.rand(n_rays, 3, generator=g) - 0.5) * 1.2
        batch = {"rays": rays, "rgb": rgb, "fg_mask": torch.ones(n_rays), "pts": pts,
                 "pts_normal": torch.nn.functional.normalize(pts, dim=-1), "pts_weights": torch.rand(n_rays, generator=g)}
        optim.zero_grad()
        loss = system.training_step(batch)["loss"]
        loss.backward()
        optim.step()
        sched.step()
        print(f"step {step}: loss {float(loss):.6f} lr {[round(pg['lr'], 8) for pg in optim.param_groups]}")
    sd = {}
    for k, v in model.state_dict().items():
        # the stub wraps nerfacc's grid as `.inner`; nerfacc's own buffer names sit directly under occupancy_grid*
        sd["model." + k.replace(".inner.", ".")] = v.detach().clone()
    # nerfacc 0.3.3's OccupancyGrid persists four buffers; the oracle's grid is not an nn.Module, so they
    # are written here under nerfacc's names from the grids' state (occs as the refresh would leave a binary grid)
    for name in ("occupancy_grid", "occupancy_grid_bg"):
        inner = getattr(model, name).inner
        res = int(inner.binary.shape[0])
        sd[f"model.{name}._roi_aabb"] = torch.as_tensor(inner.roi_aabb, dtype=torch.float32).reshape(-1).clone()
        sd[f"model.{name}.resolution"] = torch.tensor([res, res, res], dtype=torch.int32)
        sd[f"model.{name}.occs"] = inner.binary.reshape(-1).float() * 0.5
        sd[f"model.{name}._binary"] = inner.binary.clone()
    ckpt = {"epoch": 0, "global_step": N_STEPS, "pytorch-lightning_version": "1.7.7", "state_dict": sd,
            "optimizer_states": [optim.state_dict()], "lr_schedulers": [sched.state_dict()],
            "_param_names": param_names, "_lrs": [pg["lr"] for pg in optim.param_groups]}
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    path = os.path.join(HERE, "ref_checkpoint.ckpt.gz")
    with gzip.open(path, "wb", compresslevel=9) as f:
        f.write(buf.getvalue())
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KiB ({len(buf.getvalue()) / 1e6:.1f} MB raw), "
          f"{len(param_names)} optimizer parameters, keys {sorted(k for k in sd if 'occupancy' in k)}")


if __name__ == "__main__":
    main()
