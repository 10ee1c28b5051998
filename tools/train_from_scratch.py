"""End-to-end run of the caller side on the kernels: instant_angelo_b200.systems.NeuSSystem (mirror of reference
systems/neus.py) trained FROM SCRATCH on the synthetic sphere scene with the shipped schedule of
neuralangelo-colmap_sparse.yaml (dynamic ray sampling from 256 rays, progressive levels from step 5000 on are not reached in a
short run, LR warm-up over 500 steps, occupancy refresh every 16 steps), then validated: PSNR of a held-in view rendered in
eval mode, |sdf| and normal agreement on the analytic surface, and the extracted mesh's radius statistics.

    python tools/train_from_scratch.py [steps=2000] [cameras=24] [size=256]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instant_angelo_b200 import configs
from instant_angelo_b200.synthetic import SphereDataset, SphereScene
from instant_angelo_b200.systems import NeuSSystem

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_cam = int(sys.argv[2]) if len(sys.argv) > 2 else 24
size = int(sys.argv[3]) if len(sys.argv) > 3 else 256
cfg = configs.neuralangelo_colmap_sparse("finite_difference", mlp_otype="FullyFusedMLP")
ds = SphereDataset(n_cameras=n_cam, width=size, height=size, focal=0.6 * size, n_points=8192, device="cuda")
torch.manual_seed(42)
system = NeuSSystem(cfg, dataset=ds, device="cuda", device_sampling=True)
system.seed_everything(42)


def validate(tag):
    system.model.eval()
    batch = {"index": torch.tensor([0])}
    system.on_validation_batch_start(batch)
    psnr = float(system.validation_step(batch)["psnr"])
    pts, nrm, _ = SphereScene(seed=42).surface_points(20000, torch.Generator().manual_seed(5))
    with torch.no_grad():
        sdf, grad = system.model.geometry(pts.cuda(), with_grad=True, with_feature=False)
    cosn = (torch.nn.functional.normalize(grad, dim=-1) * nrm.cuda()).sum(-1)
    print(f"[{tag}] step {system.global_step}: val PSNR {psnr:.2f} dB | on the analytic surface (r = 0.5): mean |sdf| {float(sdf.abs().mean()):.4f}, "
          f"mean ||grad|-1| {float((grad.norm(dim=-1) - 1).abs().mean()):.4f}, mean cos(normal, true) {float(cosn.mean()):.4f} | "
          f"inv_s {float(system.model.variance.inv_s):.1f}", flush=True)
    system.model.train()
    return psnr


validate("init")
torch.cuda.synchronize()
t0, rays_total = time.perf_counter(), 0
for i in range(steps):
    loss = system.fit_step()
    rays_total += system.train_num_rays
    if (i + 1) % 250 == 0:
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"step {i + 1:5d} loss {float(loss):.4f} rays/step {system.train_num_rays} samples/ray "
              f"{system.model.last_num_samples_full / max(system.train_num_rays, 1):.0f} | {dt:.1f} s, {rays_total / dt / 1e3:.0f} k rays/s wall", flush=True)
    if (i + 1) % 1000 == 0:
        validate("val")
final = validate("final")
mesh = system.export(os.path.join(os.environ.get("TMPDIR", "/tmp"), "sphere.obj"))
r = mesh["v_pos"].norm(dim=-1)
print(f"mesh: {mesh['v_pos'].shape[0]} vertices, {mesh['t_pos_idx'].shape[0]} faces; vertex radius mean {float(r.mean()):.4f} "
      f"std {float(r.std()):.4f} min {float(r.min()):.4f} max {float(r.max()):.4f} (analytic sphere: 0.5)")
assert final > 20.0, final
