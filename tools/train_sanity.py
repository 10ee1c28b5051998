"""End-to-end sanity: train the bench workload for a few hundred steps on the synthetic sphere scene and print the loss
terms -- the photometric loss must fall and nothing may turn non-finite.  usage: python tools/train_sanity.py [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import bench
from instant_angelo_b200.losses import training_loss

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
args = argparse.Namespace(mlp="tc", rays=8192, steps=3, warmup=3, grad_type=sys.argv[2] if len(sys.argv) > 2 else "finite_difference")
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
gs = bench.GLOBAL_STEP0
hist = []
for i in range(steps):
    (buf, bgc), = bench.make_batches(1, 8192, 1000 + i, pin=False)
    b, bg = bench.unpack_batch(buf.to(dev), bgc.to(dev))
    model.update_step(0, gs)
    model.background_color = bg
    arena.zero_grad(); var_arena.zero_grad()
    out = model(b["rays"])
    terms = training_loss(model, out, b, cfg.system.loss, gs)
    terms["loss"].backward()
    opt.step(gs); opt_var.step(gs)
    gs += 1
    if i % 25 == 0 or i == steps - 1:
        t = {k: float(v.detach()) for k, v in terms.items()}
        psnr = -10.0 * torch.log10(torch.tensor(max(t["rgb_mse"], 1e-12))).item()
        hist.append((i, t["rgb_mse"]))
        print(f"step {i:4d} loss {t['loss']:.5f} rgb_mse {t['rgb_mse']:.5f} (PSNR {psnr:5.2f}) eikonal {t['eikonal']:.5f} "
              f"curv {t.get('curvature', 0):.5f} samples/ray {int(out['num_samples_full']) / 8192:.1f} inv_s {float(out['inv_s']):.2f}", flush=True)
        assert all(torch.isfinite(torch.tensor(v)) for v in t.values()), t
assert hist[-1][1] < 0.6 * hist[0][1], f"photometric loss did not fall: {hist[0]} -> {hist[-1]}"
print("OK: rgb_mse fell from %.5f to %.5f" % (hist[0][1], hist[-1][1]))
