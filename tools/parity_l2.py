"""Relative L2 error of every output and parameter gradient of the CUDA path against the reference-generated fixtures
(tests/golden/*.npz): the aggregate companion of the entry-wise 1e-3 checks of tests/test_gpu_model.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import GOLDEN_CASES, golden_batch, golden_state_dict, load_golden
from tests.golden.scenes import golden_loss_config, golden_model_config
from tests.test_gpu_model import build_product
from instant_angelo_b200.losses import training_loss

gd = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for case in GOLDEN_CASES:
    for otype in ("VanillaMLP", "FullyFusedMLP"):
        fx = load_golden(gd, case)
        cfg = golden_model_config(**GOLDEN_CASES[case])
        gs = int(fx["global_step"])
        model = build_product(cfg, golden_state_dict(fx), gs, torch.from_numpy(fx["background_color"]), otype)
        batch = golden_batch(fx, "cuda")
        c = lambda k: torch.from_numpy(fx[k]).cuda() if k in fx else None
        out = model(batch["rays"], stratified_u=c("u_fg"), rand_directions=c("rand_directions"), stratified_u_bg=c("u_bg"))
        terms = training_loss(model, out, batch, golden_loss_config(), gs)
        terms["loss"].backward()
        worst = []
        for k in ("comp_rgb_full", "opacity", "depth", "sdf_samples", "sdf_grad_samples", "sdf_laplace_samples", "weights"):
            if "out." + k in fx:
                e = torch.from_numpy(fx["out." + k]).double().reshape(-1); a = out[k].detach().cpu().double().reshape(-1)
                worst.append((float((a - e).norm() / e.norm().clamp_min(1e-30)), "out " + k))
        for n, p in model.named_parameters():
            if "grad." + n in fx and p.grad is not None:
                e = torch.from_numpy(fx["grad." + n]).double().reshape(-1); a = p.grad.detach().cpu().double().reshape(-1)
                if float(e.norm()) > 0:
                    worst.append((float((a - e).norm() / e.norm()), "grad " + n))
        worst.sort(reverse=True)
        print(f"{case:20s} {otype:14s} loss {float(terms['loss']):.6f} (fixture {float(fx['loss']):.6f}) | worst relative L2: " +
              ", ".join(f"{n} {v:.1e}" for v, n in worst[:4]) + f" | median {np.median([v for v, _ in worst]):.1e} over {len(worst)} tensors")
