"""Small driver for ncu: one forward + backward of the geometry tap network (35 -> 64 -> 64 -> 1 of 65)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instant_angelo_b200 import _lib as L, ops

prec = L.IA_MLP_TC_F16 if (len(sys.argv) < 2 or sys.argv[1] == "tc") else L.IA_MLP_FP32
case = sys.argv[2] if len(sys.argv) > 2 else "geo"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
torch.manual_seed(0)
if case == "geo":
    n0, n1, nh, nout, act, nou = 3, 32, 2, 65, L.IA_ACT_SOFTPLUS100, 1
elif case == "tex88":       # colour head after the output-layer fold (mlp_tc_bwd_duo96_kernel)
    n0, n1, nh, nout, act, nou = 0, 88, 2, 3, L.IA_ACT_RELU, 3
else:
    n0, n1, nh, nout, act, nou = 0, 87, 2, 3, L.IA_ACT_RELU, 3
desc = ops.make_mlp_desc(n0, n1, nh, nout, act, 2.0, -1.0, prec)
npar = L.load().ia_mlp_param_count(desc)
flat = (torch.randn(npar, device="cuda") * 0.15).requires_grad_(True)
a = torch.rand(n, n0, device="cuda").requires_grad_(True) if n0 else None
b = (torch.randn(n, n1, device="cuda") * 0.3).requires_grad_(True)
go = torch.randn(n, nou, device="cuda") * 1e-4
for it in range(3):
    y = ops.mlp_apply(a, b, flat, desc, nou)
    y.backward(go)
torch.cuda.synchronize()
e0, e1, e2 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e0.record(); y = ops.mlp_apply(a, b, flat, desc, nou); e1.record(); y.backward(go); e2.record()
torch.cuda.synchronize()
print(f"{case} prec={prec} n={n}: fwd {e0.elapsed_time(e1):.3f} ms, bwd {e1.elapsed_time(e2):.3f} ms")

if "--timing" in sys.argv:   # needs `IA_TC_TIMING=1 python -m instant_angelo_b200.build` (force a rebuild of mlp_tc.cu)
    import ctypes as C
    lib = L.load()
    buf = (C.c_ulonglong * 8)()
    for name, fn in (("fwd", lambda: ops.mlp_apply(a, b, flat, desc, nou)), ("fwd+bwd", lambda: ops.mlp_apply(a, b, flat, desc, nou).backward(go))):
        lib.ia_debug_tc_timing(1, None)
        fn(); torch.cuda.synchronize()
        lib.ia_debug_tc_timing(0, buf)
        v = list(buf)
        tiles = max(v[4], 1)
        print(f"{name}: tiles {v[4]} | per tile (thread 0): total {v[3]/tiles:.0f} cyc, barrier-wait {v[0]/tiles:.0f}, mma-issue {v[1]/tiles:.0f}, "
              f"mma-wait(run_mma) {v[2]/tiles:.0f}, split waits m0 {v[5]/tiles:.0f} m1 {v[6]/tiles:.0f} m2 {v[7]/tiles:.0f}")
