"""Timing of the wide output layer kernels (linear64.cu): plain linear64 and the fused SDF head, n = 1.56 M rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from instant_angelo_b200 import ops, _lib as L
n = 1560000
dev = "cuda"
h = torch.randn(n, 64, device=dev, requires_grad=True)
W = (torch.randn(65, 64, device=dev) * 0.1).requires_grad_(True)
b = torch.zeros(65, device=dev, requires_grad=True)
pts, enc, nrm = torch.rand(n, 3, device=dev), torch.randn(n, 16, device=dev), torch.randn(n, 3, device=dev)
go = torch.randn(n, 65, device=dev)
lib, s = L.load(), L.stream()

def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); bb.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(bb))
    return sorted(ts[2:])[len(ts[2:]) // 2]

out65 = torch.empty(n, 65, device=dev); tin = torch.empty(n, 87, device=dev)
sdf = torch.empty(n, device=dev); raw = torch.empty(n, 3, device=dev)
hd, Wd, bd = h.detach(), W.detach(), b.detach()
print("linear64_fwd ld=65      %.3f ms" % timed(lambda: lib.ia_linear64_fwd(hd.data_ptr(), n, Wd.data_ptr(), bd.data_ptr(), 65, out65.data_ptr(), 65, s)))
print("linear64_fwd ld=87      %.3f ms" % timed(lambda: lib.ia_linear64_fwd(hd.data_ptr(), n, Wd.data_ptr(), bd.data_ptr(), 65, tin.data_ptr(), 87, s)))
print("sdf_head_fwd (87 cols)  %.3f ms" % timed(lambda: lib.ia_sdf_head_fwd(hd.data_ptr(), n, Wd.data_ptr(), bd.data_ptr(), 65, pts.data_ptr(), enc.data_ptr(), 16,
                                                                             nrm.data_ptr(), tin.data_ptr(), 87, sdf.data_ptr(), raw.data_ptr(), s)))
dtin = torch.randn(n, 87, device=dev); dex = torch.randn(n, 4, device=dev)
dh = torch.empty(n, 64, device=dev); dW = torch.zeros(65, 64, device=dev); db = torch.zeros(65, device=dev)
dp, de, dn = torch.empty(n, 3, device=dev), torch.empty(n, 16, device=dev), torch.empty(n, 3, device=dev)
print("linear64_bwd dh only    %.3f ms" % timed(lambda: lib.ia_linear64_bwd(hd.data_ptr(), n, Wd.data_ptr(), dtin.data_ptr(), 87, 65, dex.data_ptr(), 4, dh.data_ptr(), None, None, s)))
print("linear64_bwd dW only    %.3f ms" % timed(lambda: lib.ia_linear64_bwd(hd.data_ptr(), n, Wd.data_ptr(), dtin.data_ptr(), 87, 65, dex.data_ptr(), 4, None, dW.data_ptr(), db.data_ptr(), s)))
print("sdf_head_bwd (all)      %.3f ms" % timed(lambda: lib.ia_sdf_head_bwd(hd.data_ptr(), n, Wd.data_ptr(), dtin.data_ptr(), 87, 65, 16, dex.data_ptr(), 4, dh.data_ptr(), dW.data_ptr(),
                                                                             db.data_ptr(), dp.data_ptr(), de.data_ptr(), dn.data_ptr(), s)))
y = ops.linear64(h, W, b)
ref = torch.addmm(b, h, W.t())
print("max err", float((y - ref).abs().max().detach()))
