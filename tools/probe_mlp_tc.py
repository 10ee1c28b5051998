"""Diagnostic: tcgen05 MLP (IA_MLP_TC_F16) against the fp32 FFMA kernel and a float64 torch reference."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instant_angelo_b200 import _lib as L, ops

CASES = {"geometry_sdf": (3, 32, 2, 65, True, 1), "texture": (0, 87, 2, 3, False, 3), "v3_weight": (0, 71, 2, 1, False, 1),
         "bg_geometry": (3, 32, 1, 8, False, 8), "bg_texture": (0, 24, 2, 3, False, 3), "v3_cam": (0, 80, 2, 3, False, 3)}


ZMIN = {}


def ref64(a, b, Ws, bs, softplus, nou):
    """float64 reference; ZMIN['z'] = per row, the smallest |pre-activation| over all hidden units (distance to a ReLU kink)."""
    x = torch.cat([a * 2 - 1, b], 1) if a is not None else b
    h = x.double()
    zmin = torch.full((h.shape[0],), float("inf"), device=h.device, dtype=torch.float64)
    for i, (w, bias) in enumerate(zip(Ws, bs)):
        h = h @ w.double().t() + bias.double()
        if i < len(Ws) - 1:
            zmin = torch.minimum(zmin, h.detach().abs().min(dim=1).values)
            h = torch.nn.functional.softplus(h, beta=100) if softplus else torch.relu(h)
    ZMIN["z"] = zmin
    return h[:, :nou]


def main():
    torch.manual_seed(0)
    dev = "cuda"
    worst = 0.0
    for name, (n0, n1, nh, nout, softplus, nou) in CASES.items():
        for n in (1, 777, 50000):
            din = n0 + n1
            dims = [din] + [64] * nh + [nout]
            Ws = [(torch.randn(dims[i + 1], dims[i], device=dev) * (1.5 / dims[i] ** 0.5)).requires_grad_(True) for i in range(len(dims) - 1)]
            bs = [(torch.randn(dims[i + 1], device=dev) * 0.1).requires_grad_(True) for i in range(len(dims) - 1)]
            a = torch.rand(n, n0, device=dev).requires_grad_(True) if n0 else None
            b = (torch.randn(n, n1, device=dev) * 0.3).requires_grad_(True)
            go = torch.randn(n, nou, device=dev) * 1e-4
            y64 = ref64(a, b, Ws, bs, softplus, nou)
            y64.backward(go.double())
            g64 = {"in1": b.grad.clone(), "in0": a.grad.clone() if n0 else None,
                   "params": torch.cat([t.grad.reshape(-1) for pair in zip(Ws, bs) for t in pair])}
            flat = torch.cat([t.detach().reshape(-1) for pair in zip(Ws, bs) for t in pair])
            res = {}
            for prec, pname in ((L.IA_MLP_FP32, "fp32"), (L.IA_MLP_TC_F16, "tc")):
                desc = ops.make_mlp_desc(n0, n1, nh, nout, L.IA_ACT_SOFTPLUS100 if softplus else L.IA_ACT_RELU, 2.0, -1.0, prec)
                fp = flat.clone().requires_grad_(True)
                ag = a.detach().clone().requires_grad_(True) if n0 else None
                bg = b.detach().clone().requires_grad_(True)
                y = ops.mlp_apply(ag, bg, fp, desc, nou)
                y.backward(go)
                torch.cuda.synchronize()
                rel = lambda got, want: float((got.double() - want.double()).abs().max() / want.double().abs().max().clamp_min(1e-30))
                res[pname] = (rel(y, y64), rel(bg.grad, g64["in1"]), rel(ag.grad, g64["in0"]) if n0 else 0.0, rel(fp.grad, g64["params"]))
                if res[pname][1] > 1e-3:
                    # which row carries the d(in1) outlier, and how far its nearest hidden pre-activation is from the ReLU kink
                    row_err = (bg.grad.double() - g64["in1"].double()).abs().max(dim=1).values
                    r = int(row_err.argmax())
                    off = row_err.clone(); off[r] = 0
                    print(f"    {pname}: worst d(in1) row {r}: min |pre-activation| of that row in float64 = {float(ZMIN['z'][r]):.3e} "
                          f"(rows within 1e-5 of a kink: {int((ZMIN['z'] < 1e-5).sum())}); max rel err over all OTHER rows = "
                          f"{float(off.max() / g64['in1'].double().abs().max()):.1e}", flush=True)
            print(f"{name:13s} n={n:6d}  fp32: out {res['fp32'][0]:.1e} din1 {res['fp32'][1]:.1e} din0 {res['fp32'][2]:.1e} dpar {res['fp32'][3]:.1e}"
                  f"   tc: out {res['tc'][0]:.1e} din1 {res['tc'][1]:.1e} din0 {res['tc'][2]:.1e} dpar {res['tc'][3]:.1e}", flush=True)
            worst = max(worst, *res["tc"])
    print("worst tc rel err", worst)


if __name__ == "__main__":
    main()
