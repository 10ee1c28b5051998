"""The caller side of the hot path: a Lightning-free mirror of reference systems/base.py:9-75 (BaseSystem) and
systems/neus.py:21-206 (NeuSSystem), plus the optimizer / scheduler parsing of systems/utils.py:314-346.

What the reference gets from pytorch_lightning (the loop that calls on_train_batch_start -> training_step ->
backward -> optimizer.step -> scheduler.step and keeps global_step) is `NeuSSystem.fit_step` here.  Everything else
keeps the reference's names, arguments and order of random draws:

  preprocess_data(batch, stage)   systems/neus.py:35-118   per-step (image, y, x) / sparse-point sampling, get_rays,
                                                            background colour, mask blending
  training_step(batch, batch_idx) systems/neus.py:120-206  forward, dynamic ray-count adaptation, loss terms
  validation_step / test_step     systems/neus.py:208-227, 256-270 (PSNR only: image/video writers are out of scope)
  C(value)                        systems/base.py:28-45
  configure_optimizers()          systems/neus.py:312-318 -> parse_optimizer / parse_scheduler

B200-side differences, none of which change results:
  * parameters of all groups that share hyper-parameters live in one flat arena (dp.ParamArena) stepped by one fused
    AdamW launch (csrc/adam.cu); the schedulers are closed-form factors of the step count (no torch scheduler objects);
  * dynamic ray sampling reads the marched sample count the marcher already brought to the host
    (NeuSModel.last_num_samples_full) instead of `out['num_samples_full'].sum().item()`, which would drain the queue;
  * with `device_sampling=True` the per-step index draws happen on the dataset's device (no CPU
    randint + index H2D per step).  The default draws from torch's global CPU generator in the reference's order, which
    is what the golden fixture (tests/golden/system_preprocess.npz, produced by the reference's own preprocess_data)
    pins bit for bit.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import registry as models
from .config import to_primitive
from .dp import FusedAdamW, ParamArena
from .losses import C as _C
from .losses import training_loss
from .network_utils import update_module_step


# ---------------------------------------------------------------------------------------------------------------------
# models/ray_utils.py:9-43
# ---------------------------------------------------------------------------------------------------------------------
def get_ray_directions(W, H, fx, fy, cx, cy, use_pixel_centers=True) -> torch.Tensor:
    """(H, W, 3) camera-space directions: x right, y up, looking along -z (models/ray_utils.py:9-21)."""
    pixel_center = 0.5 if use_pixel_centers else 0.0
    i = torch.arange(W, dtype=torch.float32) + pixel_center
    j = torch.arange(H, dtype=torch.float32) + pixel_center
    j, i = torch.meshgrid(j, i, indexing="ij")
    return torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)


def get_rays(directions: torch.Tensor, c2w: torch.Tensor, keepdim: bool = False):
    """Rotate camera-space directions into the world and attach the camera centres (models/ray_utils.py:24-43)."""
    assert directions.shape[-1] == 3
    if directions.ndim == 2:                      # (N_rays, 3) with (N_rays | 1, 3|4, 4)
        assert c2w.ndim == 3
        rays_d = (directions[:, None, :] * c2w[:, :3, :3]).sum(-1)
        rays_o = c2w[:, :, 3].expand(rays_d.shape)
    elif directions.ndim == 3:                    # (H, W, 3)
        if c2w.ndim == 2:
            rays_d = (directions[:, :, None, :] * c2w[None, None, :3, :3]).sum(-1)
            rays_o = c2w[None, None, :, 3].expand(rays_d.shape)
        elif c2w.ndim == 3:
            rays_d = (directions[None, :, :, None, :] * c2w[:, None, None, :3, :3]).sum(-1)
            rays_o = c2w[:, None, None, :, 3].expand(rays_d.shape)
        else:
            raise ValueError(f"c2w with {c2w.ndim} dimensions")
    else:
        raise ValueError(f"directions with {directions.ndim} dimensions")
    if not keepdim:
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    return rays_o, rays_d


# ---------------------------------------------------------------------------------------------------------------------
# systems/utils.py:314-346 -- schedulers as closed-form factors of the step count
# ---------------------------------------------------------------------------------------------------------------------
def parse_scheduler(config) -> Dict[str, object]:
    """Returns {'factor': f, 'interval': 'step' | 'epoch'} where lr(t) = base_lr * f(t) reproduces the torch scheduler
    the reference would build (`SequentialLR`, `Chained`, or any of LinearLR / ExponentialLR / ConstantLR / StepLR /
    MultiStepLR by name) after t scheduler steps."""
    interval = config.get("interval", "epoch")
    assert interval in ["epoch", "step"]
    name = config["name"]
    if name == "SequentialLR":
        subs = [parse_scheduler(c)["factor"] for c in config["schedulers"]]
        milestones = [int(m) for m in config["milestones"]]
        assert len(milestones) == len(subs) - 1, "SequentialLR expects one milestone fewer than schedulers"

        def factor(t: int) -> float:
            k = sum(1 for m in milestones if t >= m)
            return subs[k](t - (milestones[k - 1] if k > 0 else 0))
    elif name == "Chained":
        subs = [parse_scheduler(c)["factor"] for c in config["schedulers"]]

        def factor(t: int) -> float:
            return math.prod(f(t) for f in subs)
    else:
        args = dict(config.get("args", {}))
        if name == "LinearLR":
            s, e, n = float(args.get("start_factor", 1.0 / 3)), float(args.get("end_factor", 1.0)), int(args.get("total_iters", 5))
            factor = lambda t: s + (e - s) * min(t, n) / n
        elif name == "ExponentialLR":
            g = float(args["gamma"])
            factor = lambda t: g ** t
        elif name == "ConstantLR":
            f0, n = float(args.get("factor", 1.0 / 3)), int(args.get("total_iters", 5))
            factor = lambda t: f0 if t < n else 1.0
        elif name == "StepLR":
            n, g = int(args["step_size"]), float(args.get("gamma", 0.1))
            factor = lambda t: g ** (t // n)
        elif name == "MultiStepLR":
            ms, g = sorted(int(m) for m in args["milestones"]), float(args.get("gamma", 0.1))
            factor = lambda t: g ** sum(1 for m in ms if t >= m)
        else:
            raise NotImplementedError(f"scheduler {name!r}")
    return {"factor": factor, "interval": interval}


def _getattr_recursive(m, attr: str):
    for name in attr.split("."):
        m = getattr(m, name)
    return m


def get_parameters(model, name: str) -> List[nn.Parameter]:
    """systems/utils.py:305-311."""
    module = _getattr_recursive(model, name)
    if isinstance(module, nn.Module):
        return list(module.parameters())
    if isinstance(module, nn.Parameter):
        return [module]
    return []


class OptimizerGroups:
    """What parse_optimizer returns here: one (ParamArena, FusedAdamW) per distinct hyper-parameter set.  `param_groups`
    lists the reference's groups (name, lr, ...) for logging / checkpoint parity."""

    def __init__(self, arenas: List[ParamArena], optimizers: List[FusedAdamW], param_groups: List[dict]):
        self.arenas, self.optimizers, self.param_groups = arenas, optimizers, param_groups
        self._where = {}            # id(param) -> (optimizer, offset into its arena)
        for a, o in zip(arenas, optimizers):
            for p, off in zip(a.params, a.offsets):
                self._where[id(p)] = (o, off)

    # ---- torch.optim.AdamW.state_dict() layout, so that Lightning checkpoints of the reference resume here and back ----
    def state_dict(self, lr_factor: float = 1.0) -> dict:
        state, groups, idx = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                hit = self._where.get(id(p))
                if hit is not None and hit[0].t > 0:
                    o, off = hit
                    n = p.numel()
                    state[idx] = {"step": torch.tensor(float(o.t)), "exp_avg": o.m[off:off + n].view_as(p).clone(),
                                  "exp_avg_sq": o.v[off:off + n].view_as(p).clone()}
                ids.append(idx)
                idx += 1
            # every key torch.optim.AdamW keeps in a param group (torch 1.13 ... 2.x): Optimizer.load_state_dict REPLACES the
            # live groups with the saved ones, so a missing key is a KeyError in the reference's next optimizer.step().
            # `lr` is the scheduled rate at the saved step and `initial_lr` the base rate, as torch's schedulers leave them.
            groups.append({"name": g["name"], "lr": g["lr"] * float(lr_factor), "betas": g["betas"], "eps": g["eps"],
                           "weight_decay": g["weight_decay"], "amsgrad": False, "maximize": False, "foreach": None,
                           "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": True,
                           "initial_lr": g["lr"], "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        flat = [p for g in self.param_groups for p in g["params"]]
        n_saved = sum(len(g["params"]) for g in sd["param_groups"])
        if n_saved != len(flat):
            raise ValueError(f"loaded state dict contains {n_saved} parameters, the optimizer has {len(flat)}")
        steps = set()
        for idx, st in sd["state"].items():
            p = flat[int(idx)]
            hit = self._where.get(id(p))
            if hit is None:
                continue
            o, off = hit
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {idx}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
            o.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            o.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add((id(o), int(float(st["step"]))))
        for o in self.optimizers:
            mine = {t for oid, t in steps if oid == id(o)}
            if len(mine) > 1:
                raise ValueError(f"parameters of one fused group were saved at different step counts {sorted(mine)}")
            if mine:
                o.t = mine.pop()

    def zero_grad(self) -> None:
        for a in self.arenas:
            a.zero_grad()

    def all_reduce(self, group=None) -> None:
        for a in self.arenas:
            a.all_reduce(group)

    def broadcast_params(self, src: int = 0, group=None) -> None:
        for a in self.arenas:
            a.broadcast_params(src, group)

    def step(self, scheduler_step: int, grad_scale: float = 1.0) -> None:
        for o in self.optimizers:
            o.step(scheduler_step, grad_scale)

    def lr(self, scheduler_step: int) -> List[float]:
        return [o.lr_at(scheduler_step) for o in self.optimizers]


def parse_optimizer(config, model, schedule: Optional[Callable[[int], float]] = None) -> OptimizerGroups:
    """systems/utils.py:314-326 for `AdamW` / `Adam` (the names the shipped configs use).  Parameters of modules not
    named under `config.params` are not optimised, as in the reference."""
    name = config["name"]
    if name not in ("AdamW", "Adam", "FusedAdam"):
        raise NotImplementedError(f"optimizer {name!r}: the fused kernel implements Adam/AdamW")
    base = dict(to_primitive(config.get("args", {})))
    if "params" in config:
        named = [(n, get_parameters(model, n), dict(to_primitive(a))) for n, a in config["params"].items()]
    else:
        named = [("model", list(model.parameters()), {})]
    buckets: Dict[tuple, List[nn.Parameter]] = {}
    param_groups, seen = [], set()
    for gname, params, gargs in named:
        hp = {**base, **gargs}
        if hp.get("amsgrad", False):
            raise NotImplementedError("amsgrad")
        wd = float(hp.get("weight_decay", 0.01 if name == "AdamW" else 0.0))
        if name != "AdamW" and wd != 0.0:
            raise NotImplementedError("Adam with L2 weight decay (only decoupled AdamW decay is fused)")
        betas = tuple(float(b) for b in hp.get("betas", (0.9, 0.999)))
        key = (float(hp.get("lr", 1e-3)), betas, float(hp.get("eps", 1e-8)), wd)
        fresh = []
        for p in params:
            if id(p) in seen:
                raise ValueError("some parameters appear in more than one parameter group")    # torch's own message
            seen.add(id(p))
            if p.requires_grad and p.numel() > 0:
                fresh.append(p)
        buckets.setdefault(key, []).extend(fresh)
        param_groups.append({"name": gname, "lr": key[0], "betas": betas, "eps": key[2], "weight_decay": wd,
                             "numel": sum(p.numel() for p in fresh), "params": list(params)})
    arenas, opts = [], []
    for (lr, betas, eps, wd), params in buckets.items():
        if not params:
            continue
        arena = ParamArena(params)
        arenas.append(arena)
        opts.append(FusedAdamW(arena, lr=lr, betas=betas, eps=eps, weight_decay=wd, schedule=schedule))
    return OptimizerGroups(arenas, opts, param_groups)


def torch_scheduler_state(config, base_lrs: List[float], n_steps: int) -> dict:
    """`scheduler.state_dict()` of the torch scheduler the reference's parse_scheduler (systems/utils.py:329-346) builds,
    after `n_steps` scheduler steps -- the entry Lightning stores under `lr_schedulers` in a `.ckpt`.  Built by stepping
    real torch scheduler objects over a throw-away optimizer with the same groups, so the layout is whatever the installed
    torch writes (nested `_schedulers`, `_milestones`, `last_epoch`, `_last_lr`)."""
    import warnings
    from torch.optim import lr_scheduler as S
    dummy = torch.optim.SGD([{"params": [nn.Parameter(torch.zeros(1))], "lr": float(lr)} for lr in base_lrs], lr=0.1)

    def build(c):
        name = c["name"]
        if name == "SequentialLR":
            return S.SequentialLR(dummy, [build(x) for x in c["schedulers"]], milestones=[int(m) for m in c["milestones"]])
        if name == "Chained":
            return S.ChainedScheduler([build(x) for x in c["schedulers"]])
        return getattr(S, name)(dummy, **dict(to_primitive(c.get("args", {}))))

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        sched = build(config)
        for _ in range(int(n_steps)):
            dummy.step()
            sched.step()
        return sched.state_dict()


class PSNR(nn.Module):
    """systems/criterions.py:40-53."""

    def forward(self, inputs, targets, valid_mask=None, reduction="mean"):
        assert reduction in ["mean", "none"]
        value = (inputs - targets) ** 2
        if valid_mask is not None:
            value = value[valid_mask]
        if reduction == "mean":
            return -10 * torch.log10(torch.mean(value))
        return -10 * torch.log10(torch.mean(value, dim=tuple(range(value.ndim)[1:])))


# ---------------------------------------------------------------------------------------------------------------------
# systems/base.py + systems/neus.py
# ---------------------------------------------------------------------------------------------------------------------
class NeuSSystem:
    """`NeuSSystem(config)` builds `models.make(config.model.name, config.model)` exactly as BaseSystem.__init__
    (systems/base.py:16-21).  `dataset` is any object with the attributes of the reference's ColmapDatasetBase that
    preprocess_data reads: all_c2w [N,3,4], all_images [N,H,W,C], all_fg_masks [N,H,W], directions [H,W,3] or [N,H,W,3],
    all_points [P,3], all_points_confidence [P], pts3d_normal [P,3] or None, w, h, apply_mask, has_mask and (only when
    dataset.sample_foreground_ratio < 1) all_fg_indexs / all_bg_indexs [K,3]."""

    def __init__(self, config, dataset=None, device=None, model: Optional[nn.Module] = None, device_sampling: bool = False):
        self.config = config
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.rank = self.device
        self.dataset = dataset
        self.global_step, self.current_epoch = 0, 0
        self.logged: Dict[str, object] = {}
        self.device_sampling = device_sampling
        self._sampling_generator: Optional[torch.Generator] = None
        self.prepare()
        self.model = model if model is not None else models.make(config.model.name, config.model).to(self.device)
        self.optimizers: Optional[OptimizerGroups] = None
        self._scheduler = None

    # ---- systems/neus.py:27-30 ------------------------------------------------------------------------------------
    def prepare(self):
        m = self.config.model
        self.criterions = {"psnr": PSNR()}
        self.train_num_samples = m.train_num_rays * (m.num_samples_per_ray + m.get("num_samples_per_ray_bg", 0))
        self.train_num_rays = m.train_num_rays
        ds_cfg = self.config.get("dataset", {}) or {}
        self.sample_foreground_ratio = ds_cfg.get("sample_foreground_ratio", 1.0)

    def forward(self, batch):
        return self.model(batch["rays"])

    __call__ = forward

    def C(self, value):
        return _C(to_primitive(value), self.global_step, self.current_epoch)

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    # ---- index draws ------------------------------------------------------------------------------------------------
    def seed_sampling(self, seed: int) -> None:
        """Device-side sampling stream (device_sampling=True); per-rank seeds give per-rank ray streams."""
        dev = self._dataset_device()
        self._sampling_generator = torch.Generator(device=dev).manual_seed(seed)

    def _dataset_device(self):
        return self.dataset.all_c2w.device

    def _randint(self, high: int, n: int) -> torch.Tensor:
        if self.device_sampling:
            dev = self._dataset_device()
            if self._sampling_generator is None:
                self.seed_sampling(42)
            return torch.randint(0, high, size=(n,), device=dev, generator=self._sampling_generator)
        return torch.randint(0, high, size=(n,))

    # ---- systems/neus.py:35-118 -------------------------------------------------------------------------------------
    def preprocess_data(self, batch, stage):
        ds, m = self.dataset, self.config.model
        n = self.train_num_rays
        x = y = None
        if "index" in batch:                       # validation / testing
            index = batch["index"].cpu() if not self.device_sampling else batch["index"].to(self._dataset_device())
        else:
            if m.batch_image_sampling:
                if self.sample_foreground_ratio < 1:
                    n_fg = int(n * 0.8)
                    fg_ray_index = ds.all_fg_indexs[self._randint(len(ds.all_fg_indexs), n_fg)]
                    bg_ray_index = ds.all_bg_indexs[self._randint(len(ds.all_bg_indexs), n - n_fg)]
                    ray_index = torch.cat([fg_ray_index, bg_ray_index], dim=0)
                    index, y, x = ray_index[:, 0], ray_index[:, 1], ray_index[:, 2]
                else:
                    index = self._randint(len(ds.all_images), n)
                    x = self._randint(ds.w, n)
                    y = self._randint(ds.h, n)
            else:
                index = self._randint(len(ds.all_images), 1)
                x = self._randint(ds.w, n)
                y = self._randint(ds.h, n)
        if stage in ["train"]:
            c2w = ds.all_c2w[index]
            pts_index = self._randint(len(ds.all_points), n)      # as many sparse points as rays
            pts = ds.all_points[pts_index]
            pts_weights = ds.all_points_confidence[pts_index]
            pts_normal = ds.pts3d_normal[pts_index] if ds.pts3d_normal is not None else torch.tensor([])
            if ds.directions.ndim == 3:
                directions = ds.directions[y, x]
            else:
                directions = ds.directions[index, y, x]
            rays_o, rays_d = get_rays(directions, c2w)
            rgb = ds.all_images[index, y, x].view(-1, ds.all_images.shape[-1]).to(self.device)
            fg_mask = ds.all_fg_masks[index, y, x].view(-1).to(self.device)
        else:
            c2w = ds.all_c2w[index][0]
            pts, pts_weights, pts_normal = torch.tensor([]), torch.tensor([]), torch.tensor([])
            directions = ds.directions if ds.directions.ndim == 3 else ds.directions[index][0]
            rays_o, rays_d = get_rays(directions, c2w)
            rgb = ds.all_images[index].view(-1, ds.all_images.shape[-1]).to(self.device)
            fg_mask = ds.all_fg_masks[index].view(-1).to(self.device)

        rays = torch.cat([rays_o, F.normalize(rays_d, p=2, dim=-1)], dim=-1)

        if stage in ["train"]:
            if m.background_color == "white":
                bg = torch.ones((3,), dtype=torch.float32)
            elif m.background_color == "random":
                if self.device_sampling:
                    bg = torch.rand((3,), dtype=torch.float32, device=self._dataset_device(), generator=self._sampling_generator)
                else:
                    bg = torch.rand((3,), dtype=torch.float32)
            else:
                raise NotImplementedError
        else:
            bg = torch.ones((3,), dtype=torch.float32)
        self.model.background_color = bg.to(self.device)
        if ds.apply_mask:
            rgb = rgb * fg_mask[..., None] + self.model.background_color * (1 - fg_mask[..., None])

        batch.update({"rays": rays.to(self.device), "rgb": rgb.to(self.device), "fg_mask": fg_mask.to(self.device),
                      "pts": pts.to(self.device), "pts_normal": pts_normal.to(self.device),
                      "pts_weights": pts_weights.to(self.device)})

    # ---- systems/base.py:54-72 --------------------------------------------------------------------------------------
    def on_train_batch_start(self, batch, batch_idx=0, unused=0):
        self.preprocess_data(batch, "train")
        update_module_step(self.model, self.current_epoch, self.global_step)

    def on_validation_batch_start(self, batch, batch_idx=0, dataloader_idx=0):
        self.preprocess_data(batch, "validation")
        update_module_step(self.model, self.current_epoch, self.global_step)

    on_test_batch_start = on_validation_batch_start

    # ---- systems/neus.py:120-206 ------------------------------------------------------------------------------------
    def update_train_num_rays(self, num_samples_full: int) -> None:
        """systems/neus.py:125-128: steer the ray count towards a constant number of marched samples per step."""
        m = self.config.model
        train_num_rays = int(self.train_num_rays * (self.train_num_samples / num_samples_full))
        self.train_num_rays = min(int(self.train_num_rays * 0.9 + train_num_rays * 0.1), m.max_train_num_rays)

    def training_step(self, batch, batch_idx=0):
        out = self(batch)
        if self.config.model.dynamic_ray_sampling:
            n_full = getattr(self.model, "last_num_samples_full", None)
            if n_full is None:
                n_full = int(out["num_samples_full"].sum().item())
            self.update_train_num_rays(n_full)
        sys_cfg = self.config.system
        pts = batch.get("pts")
        loss_batch = dict(batch)
        if pts is None or pts.numel() == 0:
            loss_batch["pts"] = None
        terms = training_loss(self.model, out, loss_batch, sys_cfg.loss, self.global_step,
                              has_mask=bool(getattr(self.dataset, "has_mask", False)), current_epoch=self.current_epoch)
        for k, v in terms.items():
            if k != "loss":
                self.log(f"train/loss_{k}", v)
        for k, v in self.model.regularizations(out).items():
            self.log(f"train/loss_{k}", v)
            terms["loss"] = terms["loss"] + v * self.C(sys_cfg.loss[f"lambda_{k}"])
        self.log("train/inv_s", out["inv_s"], prog_bar=True)
        for name, value in sys_cfg.loss.items():
            if name.startswith("lambda"):
                self.log(f"train_params/{name}", self.C(value))
        self.log("train/num_rays", float(self.train_num_rays), prog_bar=True)
        self.out = out
        return {"loss": terms["loss"]}

    @torch.no_grad()
    def validation_step(self, batch, batch_idx=0):
        out = self(batch)
        psnr = self.criterions["psnr"](out["comp_rgb_full"].to(batch["rgb"]), batch["rgb"])
        self.out = out
        return {"psnr": psnr, "index": batch["index"]}

    test_step = validation_step

    # ---- systems/neus.py:312-318 ------------------------------------------------------------------------------------
    def configure_optimizers(self) -> OptimizerGroups:
        sys_cfg = self.config.system
        sched = parse_scheduler(sys_cfg.scheduler) if "scheduler" in sys_cfg else None
        self._scheduler = sched
        self.optimizers = parse_optimizer(sys_cfg.optimizer, self.model, schedule=sched["factor"] if sched else None)
        return self.optimizers

    def scheduler_step_count(self) -> int:
        if self._scheduler is None or self._scheduler["interval"] == "step":
            return self.global_step
        return self.current_epoch

    # ---- what Lightning's fit loop does around one batch ------------------------------------------------------------
    def fit_step(self, batch: Optional[dict] = None, world_size: int = 1, group=None) -> torch.Tensor:
        """on_train_batch_start -> training_step -> backward -> [all-reduce] -> optimizer + scheduler step.  Returns
        the (detached, on-device) loss; nothing in here reads it back."""
        if self.optimizers is None:
            self.configure_optimizers()
        batch = {} if batch is None else batch
        self.model.train()
        self.on_train_batch_start(batch)
        self.optimizers.zero_grad()
        loss = self.training_step(batch)["loss"]
        loss.backward()
        if world_size > 1:
            self.optimizers.all_reduce(group)
        self.optimizers.step(self.scheduler_step_count(), grad_scale=1.0 / world_size)
        self.global_step += 1
        return loss.detach()

    # ---- Lightning checkpoint layout (reference: ModelCheckpoint of launch.py:86-90, loaded by trainer.fit(ckpt_path=...)) ----
    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Keys as LightningModule.state_dict() of the reference system: 'model.' + the model's own keys."""
        return {"model." + k: v for k, v in self.model.state_dict().items()}

    def save_checkpoint(self, path: str) -> None:
        """Lightning 1.7 `.ckpt` layout: `state_dict` under the `model.` prefix, `optimizer_states[0]` as
        torch.optim.AdamW.state_dict() with every param-group key torch expects (incl. `initial_lr` and the scheduled
        `lr`), `lr_schedulers[0]` as the SequentialLR / ChainedScheduler state after the steps taken so far, and the fit
        loop's step counters."""
        ckpt = {"epoch": self.current_epoch, "global_step": self.global_step, "pytorch-lightning_version": "1.7.7",
                "state_dict": {k: v.detach().cpu().clone() for k, v in self.state_dict().items()},
                "optimizer_states": [], "lr_schedulers": [], "train_num_rays": self.train_num_rays,
                "loops": {"fit_loop": {"epoch_loop.batch_progress": {"total": {"ready": self.global_step, "completed": self.global_step}},
                                       "epoch_loop.state_dict": {"_batches_that_stepped": self.global_step}}}}
        if self.optimizers is not None:
            t = self.scheduler_step_count()
            factor = float(self._scheduler["factor"](t)) if self._scheduler is not None else 1.0
            osd = self.optimizers.state_dict(lr_factor=factor)
            osd["state"] = {i: {k: v.detach().cpu() for k, v in st.items()} for i, st in osd["state"].items()}
            ckpt["optimizer_states"] = [osd]
            if self._scheduler is not None:
                base = [g["lr"] for g in self.optimizers.param_groups]
                ckpt["lr_schedulers"] = [torch_scheduler_state(self.config.system.scheduler, base, t)]
        torch.save(ckpt, path)

    def load_checkpoint(self, path_or_dict, strict: bool = True, load_optimizer: bool = True) -> None:
        """Accepts a checkpoint written by the reference (Lightning `.ckpt`: tcnn's flat `...encoding.params`, weight-norm
        `weight_g/weight_v`, nerfacc's occupancy buffers) or by save_checkpoint.  Derived device state (the marching
        bitfields) is rebuilt; parameters keep living in their optimizer arenas (copied in place)."""
        ckpt = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu", weights_only=False)
        sd = {k[len("model."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("model.")}
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        if strict and (missing or unexpected):
            raise RuntimeError(f"checkpoint does not match the model: missing {list(missing)}, unexpected {list(unexpected)}")
        for name in ("occupancy_grid", "occupancy_grid_bg"):
            grid = getattr(self.model, name, None)
            if grid is not None and grid.occs.is_cuda:
                grid.repack()
        self.global_step = int(ckpt.get("global_step", 0))
        self.current_epoch = int(ckpt.get("epoch", 0))
        if "train_num_rays" in ckpt:
            self.train_num_rays = int(ckpt["train_num_rays"])
        if load_optimizer and ckpt.get("optimizer_states"):
            if self.optimizers is None:
                self.configure_optimizers()
            self.optimizers.load_state_dict(ckpt["optimizer_states"][0])

    # ---- systems/neus.py:305-310 ------------------------------------------------------------------------------------
    def export(self, path: Optional[str] = None) -> Dict[str, torch.Tensor]:
        """model.export(config.export) and, when `path` is given, the mesh as a Wavefront OBJ with vertex colours."""
        from .isosurface import save_obj
        export_cfg = self.config.get("export", None) or {"chunk_size": 2097152, "export_vertex_color": True}
        mesh = self.model.export(export_cfg)
        if path is not None:
            save_obj(path, **mesh)
        return mesh

    def seed_everything(self, seed: int, rank: int = 0) -> None:
        """Data-parallel replicas: the occupancy grids refresh from an identically seeded device generator on every rank
        (so they stay replica-identical without the reference's per-forward buffer broadcast), while the
        ray / point sampling stream is per rank."""
        dev = self.device
        for name in ("occupancy_grid", "occupancy_grid_bg"):
            grid = getattr(self.model, name, None)
            if grid is not None and dev.type == "cuda":
                grid.generator = torch.Generator(device=dev).manual_seed(seed + 4200)
        if self.device_sampling and self.dataset is not None:
            self.seed_sampling(seed + 1000 * rank)
