"""Generates tests/golden/system_preprocess.npz and system_schedule.npz with the REFERENCE's own code:

  * `NeuSSystem.preprocess_data` (systems/neus.py:35-118) + `get_rays` / `get_ray_directions` (models/ray_utils.py) on a
    tiny random dataset, torch's global generator seeded before each call -- pins the order of the random draws, the
    (image, y, x) indexing, ray construction, background colour and mask blending of instant_angelo_b200.systems;
  * `parse_scheduler` (systems/utils.py:329-346) on the shipped SequentialLR[LinearLR, ExponentialLR] config, stepped
    by a real torch optimizer -- pins the closed-form schedule factors;
  * `NeuSSystem.training_step`'s ray-count adaptation (systems/neus.py:125-128) over a series of sample counts.

Third-party imports of the reference are stubbed exactly as in make_golden.py.  Runs only in the build container.

    python tests/golden/make_golden_system.py
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from tests.golden.make_golden import load_reference  # noqa: E402
from instant_angelo_b200.config import to_config  # noqa: E402
from instant_angelo_b200 import configs  # noqa: E402


def tiny_dataset(seed: int, directions_4d: bool, with_normals: bool, apply_mask: bool):
    g = torch.Generator().manual_seed(seed)
    n, h, w, p = 3, 5, 7, 11
    ray_utils = importlib.import_module("models.ray_utils")
    d = ray_utils.get_ray_directions(w, h, 6.0, 6.5, w / 2, h / 2)
    ds = types.SimpleNamespace(
        w=w, h=h, img_wh=(w, h), has_mask=True, apply_mask=apply_mask,
        all_c2w=torch.randn(n, 3, 4, generator=g), all_images=torch.rand(n, h, w, 3, generator=g),
        all_fg_masks=(torch.rand(n, h, w, generator=g) > 0.4).float(),
        directions=torch.stack([d, d * 1.5, d * 0.5]) if directions_4d else d,
        all_points=torch.randn(p, 3, generator=g), all_points_confidence=torch.rand(p, generator=g),
        pts3d_normal=torch.randn(p, 3, generator=g) if with_normals else None)
    fg = ds.all_fg_masks > 0.5
    ds.all_fg_indexs, ds.all_bg_indexs = torch.nonzero(fg), torch.nonzero(~fg)
    return ds


CASES = {
    # name: (stage, batch_image_sampling, sample_foreground_ratio, background_color, directions_4d, with_normals, apply_mask, n_rays, seed)
    "train_batch_image": ("train", True, 1.0, "random", False, True, True, 13, 1),
    "train_single_image": ("train", False, 1.0, "white", True, False, False, 9, 2),
    "train_fg_ratio": ("train", True, 0.5, "random", False, True, False, 10, 3),
    "validation": ("validation", True, 1.0, "random", False, True, True, 4, 4),
    "test_4d": ("test", True, 1.0, "white", True, True, False, 4, 5),
}


def main():
    models, neus_sys = load_reference()
    fx = {}
    for name, (stage, bis, ratio, bgc, d4, wn, am, n_rays, seed) in CASES.items():
        ds = tiny_dataset(100 + seed, d4, wn, am)
        self = types.SimpleNamespace(
            dataset=ds, train_num_rays=n_rays, sample_foreground_ratio=ratio, rank="cpu", device="cpu",
            config=to_config({"model": {"batch_image_sampling": bis, "background_color": bgc}}),
            model=types.SimpleNamespace(background_color=None))
        batch = {"index": torch.tensor([1])} if stage != "train" else {}
        torch.manual_seed(seed)
        neus_sys.NeuSSystem.preprocess_data(self, batch, stage)
        for k, v in vars(ds).items():
            if isinstance(v, torch.Tensor):
                fx[f"{name}.ds.{k}"] = v.numpy()
        for k, v in batch.items():
            fx[f"{name}.out.{k}"] = v.numpy()
        fx[f"{name}.out.background_color"] = self.model.background_color.numpy()
        print(name, {k: tuple(v.shape) for k, v in batch.items()})
    np.savez_compressed(os.path.join(HERE, "system_preprocess.npz"), **fx)

    # schedule: the shipped config through the reference's parse_scheduler and real torch schedulers
    sys_utils = importlib.import_module("systems.utils")
    cfg = configs.neuralangelo_colmap_sparse()
    p = torch.nn.Parameter(torch.zeros(1))
    optim = torch.optim.AdamW([{"params": [p], "lr": 0.01}, {"params": [torch.nn.Parameter(torch.zeros(1))], "lr": 0.001}],
                              lr=0.01, betas=(0.9, 0.99), eps=1e-15)
    sched = sys_utils.parse_scheduler(cfg.system.scheduler, optim)["scheduler"]
    lrs = []
    for _ in range(1500):
        lrs.append([g["lr"] for g in optim.param_groups])
        optim.step()
        sched.step()
    # ray-count adaptation, as written at systems/neus.py:125-128 (dynamic_ray_sampling), driven through training_step's
    # own lines by a fake forward that returns only num_samples_full and then stops the step
    counts = [90000, 150000, 60000, 30000, 400000, 12000, 5000, 777, 196608, 196608]
    rays = []
    sysobj = types.SimpleNamespace(
        config=to_config({"model": {"dynamic_ray_sampling": True, "max_train_num_rays": 8192}}),
        train_num_rays=256, train_num_samples=256 * (512 + 256))

    class _Sys:
        def __call__(self, batch):
            return {"num_samples_full": torch.tensor([batch["n"]], dtype=torch.int32)}

        def __getattr__(self, k):
            return getattr(sysobj, k)

        def __setattr__(self, k, v):
            setattr(sysobj, k, v)

    s = _Sys()
    for n in counts:
        try:
            neus_sys.NeuSSystem.training_step(s, {"n": n}, 0)
        except Exception:        # the step continues into the losses, which the fake output does not carry
            pass
        rays.append(sysobj.train_num_rays)
    np.savez_compressed(os.path.join(HERE, "system_schedule.npz"), lrs=np.array(lrs, dtype=np.float64),
                        counts=np.array(counts), train_num_rays=np.array(rays))
    print("lr[0], lr[499], lr[500], lr[1499]:", lrs[0], lrs[499], lrs[500], lrs[1499])
    print("train_num_rays:", rays)


if __name__ == "__main__":
    main()
