import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instant_angelo_b200 import ops
n = 1560000
h = torch.randn(n, 64, device="cuda", requires_grad=True)
W = (torch.randn(65, 64, device="cuda") * 0.1).requires_grad_(True)
b = torch.zeros(65, device="cuda", requires_grad=True)
go = torch.randn(n, 65, device="cuda")
for _ in range(3):
    y = ops.linear64(h, W, b); y.backward(go)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); y = ops.linear64(h, W, b); e[1].record(); y.backward(go); e[2].record(); torch.cuda.synchronize()
print("linear64 fwd", e[0].elapsed_time(e[1]), "ms; bwd", e[1].elapsed_time(e[2]), "ms")
ref = torch.addmm(b, h, W.t())
print("max err", float((y - ref).abs().max()))
