"""CPU: the caller side of the hot path (instant_angelo_b200/systems.py) against fixtures produced by the reference's own
systems/neus.py, systems/utils.py and models/ray_utils.py (tests/golden/make_golden_system.py)."""
import os
import types

import numpy as np
import pytest
import torch

from instant_angelo_b200 import configs
from instant_angelo_b200.config import to_config
from instant_angelo_b200.systems import (NeuSSystem, get_ray_directions, get_rays, parse_optimizer, parse_scheduler)

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    # name: (stage, batch_image_sampling, sample_foreground_ratio, background_color, apply_mask, n_rays, seed)
    "train_batch_image": ("train", True, 1.0, "random", True, 13, 1),
    "train_single_image": ("train", False, 1.0, "white", False, 9, 2),
    "train_fg_ratio": ("train", True, 0.5, "random", False, 10, 3),
    "validation": ("validation", True, 1.0, "random", True, 4, 4),
    "test_4d": ("test", True, 1.0, "white", False, 4, 5),
}


def _system(name, fx):
    stage, bis, ratio, bgc, apply_mask, n_rays, seed = CASES[name]
    ds = types.SimpleNamespace(w=7, h=5, img_wh=(7, 5), has_mask=True, apply_mask=apply_mask, pts3d_normal=None)
    for k in fx.files:
        if k.startswith(name + ".ds."):
            setattr(ds, k.split(".ds.")[1], torch.from_numpy(fx[k]))
    cfg = to_config({"model": {"name": "neus", "batch_image_sampling": bis, "background_color": bgc, "train_num_rays": n_rays,
                               "num_samples_per_ray": 8, "max_train_num_rays": 64, "dynamic_ray_sampling": True},
                     "dataset": {"sample_foreground_ratio": ratio}})
    system = NeuSSystem(cfg, dataset=ds, model=types.SimpleNamespace(background_color=None))
    return system, stage, seed


@pytest.mark.parametrize("name", list(CASES))
def test_preprocess_data_matches_reference_bit_for_bit(name):
    fx = np.load(os.path.join(GOLD, "system_preprocess.npz"))
    system, stage, seed = _system(name, fx)
    batch = {"index": torch.tensor([1])} if stage != "train" else {}
    torch.manual_seed(seed)
    system.preprocess_data(batch, stage)
    keys = [k.split(".out.")[1] for k in fx.files if k.startswith(name + ".out.")]
    assert set(keys) == set(batch) | {"background_color"}
    for k in keys:
        got = system.model.background_color if k == "background_color" else batch[k]
        want = fx[f"{name}.out.{k}"]
        assert tuple(got.shape) == want.shape, k
        assert np.array_equal(got.numpy(), want), f"{name}.{k}"


def test_preprocess_device_sampling_is_seeded_and_in_range():
    fx = np.load(os.path.join(GOLD, "system_preprocess.npz"))
    system, _, _ = _system("train_batch_image", fx)
    system.device_sampling = True
    runs = []
    for _ in range(2):
        system.seed_sampling(7)
        b = {}
        system.preprocess_data(b, "train")
        runs.append(b)
    for k in runs[0]:
        assert torch.equal(runs[0][k], runs[1][k]), k
    assert runs[0]["rays"].shape == (13, 6)
    assert torch.allclose(runs[0]["rays"][:, 3:].norm(dim=-1), torch.ones(13), atol=1e-6)


def test_get_rays_shapes_and_conventions():
    d = get_ray_directions(4, 3, 2.0, 2.0, 2.0, 1.5)
    assert d.shape == (3, 4, 3) and torch.all(d[..., 2] == -1)
    assert d[0, 0, 0] < 0 < d[0, 3, 0] and d[0, 0, 1] > 0 > d[2, 0, 1]            # x right, y up, pixel centres
    c2w = torch.eye(4)[:3][None].repeat(2, 1, 1)
    c2w[1, :, 3] = torch.tensor([1.0, 2.0, 3.0])
    o, r = get_rays(d, c2w)                                                       # (B, H, W) flattened
    assert o.shape == (24, 3) and torch.equal(r[:12], d.reshape(-1, 3)) and torch.equal(o[12:], c2w[1, :, 3].expand(12, 3))
    o, r = get_rays(d.reshape(-1, 3)[:2], c2w)
    assert torch.equal(o[1], c2w[1, :, 3])
    o, r = get_rays(d, c2w[0], keepdim=True)
    assert r.shape == (3, 4, 3)


def test_scheduler_factors_match_torch_schedulers_stepped_by_the_reference():
    fx = np.load(os.path.join(GOLD, "system_schedule.npz"))
    cfg = configs.neuralangelo_colmap_sparse()
    sched = parse_scheduler(cfg.system.scheduler)
    assert sched["interval"] == "step"
    lrs = fx["lrs"]
    for t in range(lrs.shape[0]):
        for base, want in zip((0.01, 0.001), lrs[t]):
            assert abs(base * sched["factor"](t) - want) <= 1e-12 * max(1.0, abs(want)) + 1e-15, (t, base)
    # other scheduler names against torch directly
    for conf in ({"name": "StepLR", "args": {"step_size": 7, "gamma": 0.5}},
                 {"name": "MultiStepLR", "args": {"milestones": [3, 11], "gamma": 0.3}},
                 {"name": "ConstantLR", "args": {"factor": 0.25, "total_iters": 4}},
                 {"name": "Chained", "schedulers": [{"name": "ExponentialLR", "args": {"gamma": 0.9}},
                                                    {"name": "ConstantLR", "args": {"factor": 0.5, "total_iters": 6}}]}):
        f = parse_scheduler(to_config(conf))["factor"]
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
        if conf["name"] == "Chained":
            ts = torch.optim.lr_scheduler.ChainedScheduler([torch.optim.lr_scheduler.ExponentialLR(opt, 0.9),
                                                            torch.optim.lr_scheduler.ConstantLR(opt, 0.5, 6)])
        else:
            ts = getattr(torch.optim.lr_scheduler, conf["name"])(opt, **conf["args"])
        for t in range(20):
            assert abs(f(t) - opt.param_groups[0]["lr"]) < 1e-12, (conf["name"], t)
            opt.step()
            ts.step()
    with pytest.raises(NotImplementedError):
        parse_scheduler(to_config({"name": "CosineAnnealingLR", "args": {}}))


def test_dynamic_ray_sampling_matches_reference_training_step():
    fx = np.load(os.path.join(GOLD, "system_schedule.npz"))
    cfg = configs.neuralangelo_colmap_sparse()
    system = NeuSSystem(cfg, model=types.SimpleNamespace())
    assert system.train_num_rays == 256 and system.train_num_samples == 256 * (512 + 256)
    got = []
    for n in fx["counts"]:
        system.update_train_num_rays(int(n))
        got.append(system.train_num_rays)
    assert got == fx["train_num_rays"].tolist()
    with pytest.raises(ZeroDivisionError):
        system.update_train_num_rays(0)


def test_parse_optimizer_groups_parameters_like_the_reference_config():
    cfg = configs.neuralangelo_colmap_sparse("finite_difference")
    for blk in (cfg.model.geometry, cfg.model.geometry_bg):
        blk.xyz_encoding_config["log2_hashmap_size"] = 8          # keep the CPU test tiny
    system = NeuSSystem(cfg)
    groups = system.configure_optimizers()
    names = [g["name"] for g in groups.param_groups]
    assert names == ["geometry", "texture", "geometry_bg", "texture_bg", "variance"]
    assert [g["lr"] for g in groups.param_groups] == [0.01, 0.01, 0.01, 0.01, 0.001]
    assert all(g["betas"] == (0.9, 0.99) and g["eps"] == 1e-15 and g["weight_decay"] == 0.01 for g in groups.param_groups)
    # two distinct hyper-parameter sets -> two arenas (one fused AdamW launch each); every trainable parameter is in one
    assert len(groups.arenas) == 2 and [o.lr for o in groups.optimizers] == [0.01, 0.001]
    n_model = sum(p.numel() for p in system.model.parameters() if p.requires_grad)
    assert sum(g["numel"] for g in groups.param_groups) == n_model
    owned = {id(p) for a in groups.arenas for p in a.params}
    assert owned == {id(p) for p in system.model.parameters() if p.requires_grad and p.numel() > 0}
    for a in groups.arenas:                                      # parameters are views into the arena
        for p, off in zip(a.params, a.offsets):
            assert p.data_ptr() == a.data[off:].data_ptr()
    assert groups.lr(0) == [0.01 * 0.01, 0.001 * 0.01] and abs(groups.lr(500)[0] - 0.01) < 1e-15
    # the product path has no CPU fallback: stepping on CPU tensors raises, as every other operator does
    with pytest.raises(NotImplementedError):
        groups.step(0)
    with pytest.raises(ValueError):
        parse_optimizer(to_config({"name": "AdamW", "args": {}, "params": {"geometry": {}, "geometry.network": {}}}), system.model)
    with pytest.raises(NotImplementedError):
        parse_optimizer(to_config({"name": "SGD", "args": {}}), system.model)


def test_sphere_dataset_surface():
    from instant_angelo_b200.synthetic import SphereDataset
    ds = SphereDataset(n_cameras=3, width=16, height=12, focal=10.0, n_points=50)
    assert ds.all_images.shape == (3, 12, 16, 3) and ds.all_fg_masks.shape == (3, 12, 16) and ds.directions.shape == (12, 16, 3)
    assert 0 < ds.all_fg_masks.mean() < 1 and len(ds.all_fg_indexs) + len(ds.all_bg_indexs) == 3 * 12 * 16
    assert torch.all(ds.all_images[ds.all_fg_masks == 0] == 1.0)
    assert torch.allclose(ds.all_points.norm(dim=-1), torch.full((50,), 0.5), atol=1e-6)
    cfg = configs.neuralangelo_colmap_sparse()
    system = NeuSSystem(cfg, dataset=ds, model=types.SimpleNamespace(background_color=None))
    b = {}
    torch.manual_seed(0)
    system.preprocess_data(b, "train")
    assert b["rays"].shape == (256, 6) and b["rgb"].shape == (256, 3) and b["pts"].shape == (256, 3)
    v = {"index": torch.tensor([2])}
    system.preprocess_data(v, "validation")
    assert v["rays"].shape == (12 * 16, 6) and torch.equal(system.model.background_color, torch.ones(3))


@pytest.mark.parametrize("case", ["neus_dualcolor_bg", "neus_v3_nobg"])
def test_training_step_matches_reference_training_step(golden_dir, case):
    """NeuSSystem.training_step (wiring, schedules C(), loss terms, dynamic ray count) against the loss and the logged terms
    of the reference's own training_step (tests/golden/*.npz), with the CPU oracle standing in for the CUDA model."""
    from tests.helpers import GOLDEN_CASES, assert_close, golden_batch, load_golden
    from tests.golden.scenes import golden_loss_config, golden_model_config
    from tests.test_oracle_golden import build_oracle
    fx = load_golden(golden_dir, case)
    oracle = build_oracle(fx, case)

    class Model:
        geometry = oracle.geometry
        background_color = oracle.background_color

        def __call__(self, rays):
            out = oracle.forward_(rays, stratified_u=torch.from_numpy(fx["u_fg"]), rand_directions=torch.from_numpy(fx["rand_directions"]),
                                  stratified_u_bg=torch.from_numpy(fx["u_bg"]))
            return {**out, "inv_s": oracle.variance.inv_s}

        def regularizations(self, out):
            return {}

    mcfg = golden_model_config(**GOLDEN_CASES[case])
    mcfg["dynamic_ray_sampling"] = True
    cfg = to_config({"model": mcfg, "system": {"loss": golden_loss_config()}})
    system = NeuSSystem(cfg, dataset=types.SimpleNamespace(has_mask=False), model=Model())
    system.global_step = int(fx["global_step"])
    res = system.training_step(golden_batch(fx))
    assert_close(res["loss"], fx["loss"], rtol=1e-4, atol=1e-6, name="loss")
    for k, ref_k in [("rgb_mse", "train/loss_rgb_mse"), ("eikonal", "train/loss_eikonal"), ("curvature", "train/loss_curvature"),
                     ("sdf_l1", "train/loss_sdf_l1"), ("normal_cos", "train/loss_normal_cos"), ("sparsity", "train/loss_sparsity")]:
        if "log." + ref_k in fx:
            assert_close(system.logged["train/loss_" + k], fx["log." + ref_k], rtol=1e-4, atol=1e-6, name=k)
    assert_close(system.logged["train/inv_s"], fx["log.train/inv_s"], rtol=1e-6, atol=0, name="inv_s")
    # systems/neus.py:125-128 with the fixture's sample count
    n_full = int(fx["out.num_samples_full"].reshape(-1)[0])
    n0 = mcfg["train_num_rays"]
    budget = n0 * (mcfg["num_samples_per_ray"] + (mcfg["num_samples_per_ray_bg"] if mcfg["learned_background"] else mcfg.get("num_samples_per_ray_bg", 0)))
    want = min(int(n0 * 0.9 + int(n0 * (budget / n_full)) * 0.1), mcfg["max_train_num_rays"])
    assert system.train_num_rays == want and system.logged["train/num_rays"] == float(want)
    res["loss"].backward()
    g = oracle.geometry.encoding.encoding.encoding.params.grad
    assert_close(g, fx["grad.geometry.encoding.encoding.encoding.params"], rtol=1e-3,
                 atol=1e-4 * float(np.abs(fx["grad.geometry.encoding.encoding.encoding.params"]).max()), name="table grad")


def _tiny_system():
    cfg = configs.neuralangelo_colmap_sparse("finite_difference")
    for blk in (cfg.model.geometry, cfg.model.geometry_bg):
        blk.xyz_encoding_config["log2_hashmap_size"] = 8
    torch.manual_seed(3)
    system = NeuSSystem(cfg)
    system.configure_optimizers()
    return cfg, system


def _reference_adamw(cfg, model):
    """torch.optim.AdamW exactly as the reference's parse_optimizer builds it (systems/utils.py:314-326)."""
    from instant_angelo_b200.systems import get_parameters
    oc = cfg.system.optimizer
    groups = [{"params": get_parameters(model, n), "name": n, **a} for n, a in oc.params.items()]
    return torch.optim.AdamW(groups, **oc.args)


def test_optimizer_state_dict_is_interchangeable_with_torch_adamw():
    cfg, system = _tiny_system()
    groups = system.optimizers
    ref_opt = _reference_adamw(cfg, system.model)
    g = torch.Generator().manual_seed(5)
    for a in groups.arenas:
        a.grad.copy_(torch.randn(a.grad.shape, generator=g))
    before = [a.data.clone() for a in groups.arenas]
    ref_opt.step()                                                  # one real AdamW step on the arena-backed parameters
    for a, b in zip(groups.arenas, before):
        assert not torch.equal(a.data, b)
    # torch -> fused: moments land at the parameters' arena offsets
    groups.load_state_dict(ref_opt.state_dict())
    assert all(o.t == 1 for o in groups.optimizers)
    for a, o in zip(groups.arenas, groups.optimizers):
        for p, off in zip(a.params, a.offsets):
            st = ref_opt.state[p]
            assert torch.equal(o.m[off:off + p.numel()].view_as(p), st["exp_avg"])
            assert torch.equal(o.v[off:off + p.numel()].view_as(p), st["exp_avg_sq"])
    # fused -> torch: a fresh reference optimizer accepts the layout and holds the same moments
    sd = groups.state_dict()
    assert [g_["name"] for g_ in sd["param_groups"]] == ["geometry", "texture", "geometry_bg", "texture_bg", "variance"]
    ref2 = _reference_adamw(cfg, system.model)
    ref2.load_state_dict(sd)
    for p in system.model.parameters():
        if p.numel() and p.requires_grad:
            assert torch.equal(ref2.state[p]["exp_avg"], ref_opt.state[p]["exp_avg"])
            assert float(ref2.state[p]["step"]) == 1.0
    assert [g_["lr"] for g_ in ref2.param_groups] == [0.01, 0.01, 0.01, 0.01, 0.001]
    with pytest.raises(ValueError):
        bad = ref_opt.state_dict()
        bad["param_groups"] = bad["param_groups"][:-1]
        groups.load_state_dict(bad)


def test_checkpoint_roundtrip_in_lightning_layout(tmp_path):
    cfg, system = _tiny_system()
    g = torch.Generator().manual_seed(9)
    for o in system.optimizers.optimizers:
        o.m.copy_(torch.randn(o.m.shape, generator=g))
        o.v.copy_(torch.rand(o.v.shape, generator=g))
        o.t = 37
    system.global_step, system.train_num_rays = 37, 1234
    system.model.occupancy_grid.occs.copy_(torch.rand(system.model.occupancy_grid.occs.shape, generator=g))
    system.model.occupancy_grid._binary.copy_(system.model.occupancy_grid.occs.view(128, 128, 128) > 0.5)
    path = str(tmp_path / "last.ckpt")
    system.save_checkpoint(path)
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert {"epoch", "global_step", "state_dict", "optimizer_states", "lr_schedulers"} <= set(ckpt)
    keys = set(ckpt["state_dict"])
    for k in ("model.geometry.encoding.encoding.encoding.params", "model.geometry.network.layers.0.weight_g",
              "model.geometry.network.layers.4.weight_v", "model.variance.variance", "model.occupancy_grid._binary",
              "model.occupancy_grid.occs", "model.occupancy_grid_bg.resolution", "model.scene_aabb"):
        assert k in keys, k
    assert not [k for k in keys if "bitfield" in k or "_fd_signs" in k], "derived state must not be persisted"

    _, fresh = _tiny_system()
    with torch.no_grad():
        for p in fresh.model.parameters():
            p.add_(1.0)
    arena_ptrs = [a.data.data_ptr() for a in fresh.optimizers.arenas]
    fresh.load_checkpoint(path)
    assert fresh.global_step == 37 and fresh.train_num_rays == 1234
    for (k, a), (_, b) in zip(system.model.state_dict().items(), fresh.model.state_dict().items()):
        assert torch.equal(a, b), k
    for a1, o1, o2 in zip(system.optimizers.arenas, system.optimizers.optimizers, fresh.optimizers.optimizers):
        assert o2.t == 37
        for p, off in zip(a1.params, a1.offsets):           # (alignment padding between segments is not state)
            sl = slice(off, off + p.numel())
            assert torch.equal(o1.m[sl], o2.m[sl]) and torch.equal(o1.v[sl], o2.v[sl])
    # parameters were copied in place: they still live in the optimizer arenas
    assert arena_ptrs == [a.data.data_ptr() for a in fresh.optimizers.arenas]
    for a in fresh.optimizers.arenas:
        for p, off in zip(a.params, a.offsets):
            assert p.data_ptr() == a.data[off:].data_ptr()
    # a checkpoint that does not fit is refused
    bad = dict(ckpt, state_dict={k: v for k, v in ckpt["state_dict"].items() if "variance" not in k})
    with pytest.raises(RuntimeError, match="missing"):
        fresh.load_checkpoint(bad)


def test_reference_checkpoint_keys_load(golden_dir):
    """state_dict keys written by the reference's own modules (tests/golden/*.npz 'param.*' = named_parameters of the
    reference NeuSModel with tcnn's flat table layout) load under the Lightning 'model.' prefix."""
    from tests.helpers import GOLDEN_CASES, golden_state_dict, load_golden
    from tests.golden.scenes import golden_model_config
    fx = load_golden(golden_dir, "neus_dualcolor_bg")
    cfg = to_config({"model": golden_model_config(**GOLDEN_CASES["neus_dualcolor_bg"])})
    system = NeuSSystem(cfg)
    ref_sd = golden_state_dict(fx)
    ckpt = {"state_dict": {"model." + k: v for k, v in ref_sd.items()}, "global_step": 25, "epoch": 0}
    system.load_checkpoint(ckpt, strict=False, load_optimizer=False)
    own = dict(system.model.named_parameters())
    assert set(ref_sd) == set(own), set(ref_sd) ^ set(own)
    for k, v in ref_sd.items():
        assert torch.equal(own[k].detach(), v), k
    assert system.global_step == 25


# ---------------------------------------------------------------------------------------------------------------------
# checkpoint written by the reference's own objects (tests/golden/make_golden_checkpoint.py)
# ---------------------------------------------------------------------------------------------------------------------
def _load_ref_checkpoint():
    import gzip
    import io
    with gzip.open(os.path.join(GOLD, "ref_checkpoint.ckpt.gz"), "rb") as f:
        return torch.load(io.BytesIO(f.read()), map_location="cpu", weights_only=False)


def _ref_checkpoint_system():
    from tests.golden.make_golden_checkpoint import checkpoint_config
    torch.manual_seed(11)
    system = NeuSSystem(checkpoint_config())
    system.configure_optimizers()
    return system


def test_parameter_order_is_torchs():
    """torch.optim state dicts index parameters by position: VanillaMLP must yield its parameters in the order of the
    reference's nn.Sequential of nn.Linear (weight, bias) / weight-normed nn.Linear (bias, weight_g, weight_v)
    (reference models/network_utils.py:115-134)."""
    import torch.nn as nn
    from instant_angelo_b200.network_utils import VanillaMLP
    for wn in (False, True):
        cfg = {"n_neurons": 64, "n_hidden_layers": 2, "sphere_init": wn, "weight_norm": wn}
        ours = [n for n, _ in VanillaMLP(7, 5, cfg).named_parameters()]
        lin = lambda i, o: nn.utils.weight_norm(nn.Linear(i, o)) if wn else nn.Linear(i, o)
        ref = nn.Module()
        ref.layers = nn.Sequential(lin(7, 64), nn.ReLU(), lin(64, 64), nn.ReLU(), lin(64, 5))
        assert ours == [n for n, _ in ref.named_parameters()], (wn, ours)


def test_reference_written_checkpoint_loads():
    """A Lightning-layout checkpoint whose state_dict / optimizer_states / lr_schedulers were produced by the reference's
    models, the reference's parse_optimizer (a real torch.optim.AdamW) and parse_scheduler after three training steps."""
    ckpt = _load_ref_checkpoint()
    system = _ref_checkpoint_system()
    system.load_checkpoint(ckpt)                                    # strict: every key of the reference must be consumed
    assert system.global_step == 3 and system.current_epoch == 0
    own = dict(system.model.named_parameters())
    ref_params = {k[len("model."):]: v for k, v in ckpt["state_dict"].items()}
    for k, p in own.items():
        assert torch.equal(p.detach(), ref_params[k]), k
    assert torch.equal(system.model.occupancy_grid._binary, ref_params["occupancy_grid._binary"])
    assert torch.equal(system.model.occupancy_grid_bg.occs, ref_params["occupancy_grid_bg.occs"])
    # the reference's parameter order (names recorded by the generator) is the order this model yields them in
    groups = system.optimizers
    flat = [p for g in groups.param_groups for p in g["params"]]
    names = {id(p): n for n, p in system.model.named_parameters()}
    assert [names[id(p)] for p in flat] == ckpt["_param_names"]
    # optimizer moments land on the tensors they belong to
    osd = ckpt["optimizer_states"][0]
    assert all(o.t == 3 for o in groups.optimizers)
    hit = 0
    for idx, st in osd["state"].items():
        o, off = groups._where[id(flat[idx])]
        n = flat[idx].numel()
        assert torch.equal(o.m[off:off + n].view_as(flat[idx]), st["exp_avg"]), ckpt["_param_names"][idx]
        assert torch.equal(o.v[off:off + n].view_as(flat[idx]), st["exp_avg_sq"]), ckpt["_param_names"][idx]
        hit += 1
    assert hit == len(osd["state"]) >= 25
    # the closed-form schedule continues where the reference's SequentialLR stood
    lrs = dict(zip([g["name"] for g in groups.param_groups], ckpt["_lrs"]))
    t = system.scheduler_step_count()
    for o in groups.optimizers:
        want = lrs["variance"] if abs(o.lr - 0.001) < 1e-12 else lrs["geometry"]
        assert abs(o.lr_at(t) - want) < 1e-12


def test_checkpoint_resumes_in_the_reference(tmp_path):
    """here -> reference: the optimizer_states / lr_schedulers entries written by save_checkpoint load into the REAL torch
    objects the reference builds (AdamW with the reference's groups, SequentialLR[LinearLR, ExponentialLR]), step without
    a KeyError, and continue the schedule from the saved step instead of restarting the warm-up."""
    from torch.optim.lr_scheduler import ExponentialLR, LinearLR, SequentialLR
    system = _ref_checkpoint_system()
    system.load_checkpoint(_load_ref_checkpoint())
    system.global_step = 700                                         # past the 500-step warm-up milestone
    for o in system.optimizers.optimizers:
        o.t = 700
    path = str(tmp_path / "resume.ckpt")
    system.save_checkpoint(path)
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    assert ckpt["loops"]["fit_loop"]["epoch_loop.state_dict"]["_batches_that_stepped"] == 700
    cfg = system.config
    ref_opt = _reference_adamw(cfg, system.model)
    sc = cfg.system.scheduler
    subs = [LinearLR(ref_opt, **dict(sc.schedulers[0].args)), ExponentialLR(ref_opt, **dict(sc.schedulers[1].args))]
    ref_sched = SequentialLR(ref_opt, subs, milestones=list(sc.milestones))
    ref_opt.load_state_dict(ckpt["optimizer_states"][0])
    ref_sched.load_state_dict(ckpt["lr_schedulers"][0])
    gamma = float(sc.schedulers[1].args.gamma)
    assert abs(ref_opt.param_groups[0]["lr"] - 0.01 * gamma ** 200) < 1e-12
    assert ref_opt.param_groups[0]["initial_lr"] == 0.01 and ref_opt.param_groups[4]["initial_lr"] == 0.001
    for p in system.model.parameters():
        if p.requires_grad and p.numel():
            p.grad = torch.zeros_like(p)
    ref_opt.step()                                                   # KeyError here if a param-group key were missing
    ref_sched.step()
    assert abs(ref_opt.param_groups[0]["lr"] - 0.01 * gamma ** 201) < 1e-12
    assert abs(ref_opt.param_groups[4]["lr"] - 0.001 * gamma ** 201) < 1e-13
    some = next(p for p in system.model.parameters() if p.requires_grad and p.numel())
    assert float(ref_opt.state[some]["step"]) == 701.0


def test_distortion_loss_matches_definition():
    """flatten_eff_distloss (reference systems/neus.py:163-171 via torch_efficient_distloss) against the O(S^2) definition
    sum_ij w_i w_j |m_i - m_j| + sum_i w_i^2 interval_i / 3 per ray, averaged over ray_id.max() + 1 rays."""
    from instant_angelo_b200.losses import flatten_eff_distloss, training_loss
    g = torch.Generator().manual_seed(0)
    counts = [5, 0, 1, 9, 3]
    ray_id = torch.cat([torch.full((c,), i) for i, c in enumerate(counts)])
    S = ray_id.numel()
    w = torch.rand(S, generator=g).requires_grad_(True)
    interval = torch.rand(S, generator=g) * 0.1
    m = torch.cat([torch.sort(torch.rand(c, generator=g)).values for c in counts])
    want = 0.0
    for i in range(len(counts)):
        sel = ray_id == i
        wi, mi, ii = w[sel].double(), m[sel].double(), interval[sel].double()
        want = want + (wi[:, None] * wi[None, :] * (mi[:, None] - mi[None, :]).abs()).sum() + (wi * wi * ii).sum() / 3.0
    want = want / len(counts)
    got = flatten_eff_distloss(w, m, interval, ray_id)
    assert abs(float(got) - float(want)) < 1e-6 * max(1.0, abs(float(want)))
    (gw,) = torch.autograd.grad(got, w)
    (gw_ref,) = torch.autograd.grad(want, w)
    assert torch.allclose(gw, gw_ref.float(), rtol=1e-5, atol=1e-7)
    assert float(flatten_eff_distloss(w[:0], m[:0], interval[:0], ray_id[:0])) == 0.0
