"""GPU: the caller side of the hot path (instant_angelo_b200.systems.NeuSSystem, mirror of reference systems/neus.py)
driving the kernels: data sampling on the device, training_step with dynamic ray sampling, parse_optimizer arenas + fused
AdamW with the parsed LR schedule, and the eval-mode (chunked, no-grad) validation step."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _system(n_cameras=8, size=128):
    from instant_angelo_b200 import configs
    from instant_angelo_b200.synthetic import SphereDataset
    from instant_angelo_b200.systems import NeuSSystem
    cfg = configs.neuralangelo_colmap_sparse("finite_difference", mlp_otype="FullyFusedMLP")
    ds = SphereDataset(n_cameras=n_cameras, width=size, height=size, focal=0.6 * size, n_points=4096, device="cuda")
    torch.manual_seed(42)
    system = NeuSSystem(cfg, dataset=ds, device="cuda", device_sampling=True)
    system.seed_everything(42)
    return system


def test_system_fit_steps(cuda_lib):
    system = _system()
    before = {n: p.detach().clone() for n, p in system.model.named_parameters()}
    losses, rays = [], []
    for _ in range(24):
        losses.append(system.fit_step())
        rays.append(system.train_num_rays)
    vals = [float(v) for v in losses]
    assert all(np.isfinite(v) for v in vals), vals
    assert system.global_step == 24
    # dynamic ray sampling (systems/neus.py:125-128): starts at train_num_rays=256 and steers towards
    # 256 * (512 + 256) marched samples per step, capped at max_train_num_rays
    assert rays[0] != 256 and all(1 <= r <= 8192 for r in rays), rays
    assert system.model.last_num_samples_full > 0
    assert system.logged["train/num_rays"] == float(rays[-1])
    # LR schedule of the shipped config: linear warm-up from 1 % over 500 steps; variance group at a tenth of the rest
    lr = system.optimizers.lr(23)
    assert lr[0] == pytest.approx(0.01 * (0.01 + 0.99 * 23 / 500), rel=1e-12) and lr[1] == pytest.approx(lr[0] / 10, rel=1e-12)
    moved = [n for n, p in system.model.named_parameters() if not torch.equal(p.detach(), before[n])]
    for key in ("geometry.encoding.encoding.encoding.params", "variance.variance"):
        assert key in moved, (key, moved)
    assert all(o.t == 24 for o in system.optimizers.optimizers)
    # every parameter still lives in (and trains through) its arena
    for a in system.optimizers.arenas:
        for p, off in zip(a.params, a.offsets):
            assert p.data_ptr() == a.data[off:].data_ptr() and p.grad.data_ptr() == a.grad[off:].data_ptr()


def test_system_validation_step_eval_mode(cuda_lib):
    system = _system(n_cameras=2, size=96)
    for _ in range(2):
        system.fit_step()
    system.model.eval()
    batch = {"index": torch.tensor([1])}
    system.on_validation_batch_start(batch)
    out = system.validation_step(batch)
    assert np.isfinite(float(out["psnr"])) and float(out["psnr"]) > 0
    o = system.out
    n = 96 * 96
    assert o["comp_rgb_full"].shape == (n, 3) and o["opacity"].shape == (n, 1) and o["comp_normal"].shape == (n, 3)
    assert not o["comp_rgb_full"].requires_grad and "sdf_samples" not in o
    assert torch.isfinite(o["comp_rgb_full"]).all() and torch.isfinite(o["depth"]).all()
    # eval mode is deterministic: no stratified jitter, white background (systems/neus.py:103)
    batch2 = {"index": torch.tensor([1])}
    system.on_validation_batch_start(batch2)
    out2 = system.validation_step(batch2)
    assert torch.equal(system.out["comp_rgb_full"], o["comp_rgb_full"]) and float(out2["psnr"]) == float(out["psnr"])
    assert torch.equal(system.model.background_color, torch.ones(3, device="cuda"))


def test_system_export_mesh(cuda_lib, tmp_path):
    """Export path (systems/neus.py:305-310 -> models/neus.py:308-318 -> models/geometry.py:80-113) on the kernels: the
    sphere-initialised SDF of a fresh model meshes into one closed, outward-oriented surface whose vertices sit on the
    0.001 level of the network that produced them."""
    from tests.test_isosurface import _signed_volume
    system = _system(n_cameras=2, size=64)
    system.config.model.geometry.isosurface["resolution"] = 40
    system.config.model.geometry.isosurface["block_res"] = 32         # several blocks
    for _ in range(2):
        system.fit_step()                                              # sets the progressive levels / finite-difference eps
    path = tmp_path / "mesh.obj"
    mesh = system.export(str(path))
    nv, nf = mesh["v_pos"].shape[0], mesh["t_pos_idx"].shape[0]
    assert nv > 100 and nf > 100 and mesh["v_rgb"].shape == (nv, 3) and mesh["v_norm"].shape == (nv, 3)
    assert _signed_volume(mesh["v_pos"], mesh["t_pos_idx"]) > 0
    system.model.eval()
    with torch.no_grad():
        sdf = system.model.geometry(mesh["v_pos"].cuda().contiguous(), with_grad=False, with_feature=False)
    assert float((sdf - 0.001).abs().max()) < 2e-2                     # linear interpolation on a 0.05 lattice
    assert torch.isfinite(mesh["v_rgb"]).all() and (mesh["v_norm"].norm(dim=-1) - 1).abs().max() < 1e-4
    assert sum(1 for l in path.read_text().splitlines() if l.startswith("f ")) == nf
