"""Error of ia_mlp_fwd_grad / ia_mlp_fwd_grad_bwd and of torch's own fp32 double backward against float64 autograd
(max abs error / max |want| per tensor), for the regimes of tests/test_gpu_ops.py::test_mlp_fwd_grad_second_order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instant_angelo_b200 import _lib as L, ops

dev = "cuda"


def run(n, cot_scale, regime, dtype):
    g = torch.Generator(device=dev).manual_seed(1234 + n)
    din, nout = 35, 65
    wscale = 1.5 if regime == "geometric" else 0.08
    dims = [din, 64, 64, nout]
    Ws = [(torch.randn(dims[i + 1], dims[i], device=dev, generator=g) * (wscale / dims[i] ** 0.5)) for i in range(3)]
    Ws[2] = torch.randn(nout, 64, device=dev, generator=g) * 0.3
    bs = [(torch.randn(dims[i + 1], device=dev, generator=g) * (0.1 if regime == "geometric" else 0.01)) for i in range(3)]
    a = torch.rand(n, 3, device=dev, generator=g)
    b = torch.randn(n, 32, device=dev, generator=g) * 0.3
    ch = torch.randn(n, 64, device=dev, generator=g) * cot_scale
    cg0 = torch.randn(n, 3, device=dev, generator=g) * cot_scale
    cg1 = torch.randn(n, 32, device=dev, generator=g) * cot_scale
    if dtype == "kernel":
        flat = torch.cat([t.reshape(-1) for pair in zip(Ws, bs) for t in pair]).requires_grad_(True)
        desc = ops.make_mlp_desc(3, 32, 2, nout, L.IA_ACT_SOFTPLUS100, 2.0, -1.0, L.IA_MLP_TC_F16)
        ag, bg = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        h, g0, g1 = ops.mlp_fwd_grad(ag, bg, flat, desc)
        ((h * ch).sum() + (g0 * cg0).sum() + (g1 * cg1).sum()).backward()
        return dict(h=h, g0=g0, g1=g1, da=ag.grad, db=bg.grad, dp=flat.grad)
    W = [w.to(dtype).requires_grad_(True) for w in Ws]
    B = [t.to(dtype).requires_grad_(True) for t in bs]
    a_, b_ = a.to(dtype).requires_grad_(True), b.to(dtype).requires_grad_(True)
    x = torch.cat([a_ * 2 - 1, b_], 1)
    h1 = torch.nn.functional.softplus(x @ W[0].t() + B[0], beta=100)
    h2 = torch.nn.functional.softplus(h1 @ W[1].t() + B[1], beta=100)
    y = h2 @ W[2][0] + B[2][0]
    g0, g1 = torch.autograd.grad(y.sum(), [a_, b_], create_graph=True)
    ((h2 * ch.to(dtype)).sum() + (g0 * cg0.to(dtype)).sum() + (g1 * cg1.to(dtype)).sum()).backward()
    dp = torch.cat([t.grad.reshape(-1) if t.grad is not None else torch.zeros_like(t).reshape(-1) for pair in zip(W, B) for t in pair])
    return dict(h=h2, g0=g0, g1=g1, da=a_.grad, db=b_.grad, dp=dp)


torch.backends.cuda.matmul.allow_tf32 = False
for regime in ("geometric", "soft"):
    for n in (1, 777, 50_003, 300_007):
        for cs in (1.0, 1e-4):
            ref = run(n, cs, regime, torch.float64)
            k = run(n, cs, regime, "kernel")
            f = run(n, cs, regime, torch.float32)
            line = f"{regime:9s} n={n:7d} cot={cs:g} "
            for key in ref:
                w = ref[key].detach().double()
                sc = float(w.abs().max()) + 1e-300
                ek = float((k[key].detach().double() - w).abs().max()) / sc
                ef = float((f[key].detach().double() - w).abs().max()) / sc
                l2k = float((k[key].detach().double() - w).norm() / w.norm())
                l2f = float((f[key].detach().double() - w).norm() / w.norm())
                line += f"| {key}: k {ek:.1e}/{l2k:.1e} f32 {ef:.1e}/{l2f:.1e} "
            print(line, flush=True)
