"""Ordered kernel timeline of one bench training step (short names, durations, gap before each launch) plus, for the torch
glue, the aten operator and the Python line that issued it -- the work list for fusing glue into the kernels."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=8192)
a = ap.parse_args()
args = argparse.Namespace(mlp="tc", rays=a.rays, steps=3, warmup=3, grad_type="finite_difference")
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
batches = [(b.to(dev), g.to(dev)) for b, g in bench.make_batches(10, a.rays, 0, pin=False)]
gs = bench.GLOBAL_STEP0 + 1
def step(i):
    global gs
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
for i in range(5): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step(5)
    torch.cuda.synchronize()
evs = prof.events()
kern = [e for e in evs if e.device_type == torch.autograd.DeviceType.CUDA]
kern.sort(key=lambda e: e.time_range.start)
# map each kernel to the innermost CPU op (aten::*) that covers its launch via correlation: use linked cpu parent
cpu_ops = [e for e in evs if e.device_type == torch.autograd.DeviceType.CPU]
by_corr = {}
for e in cpu_ops:
    for k in getattr(e, "kernels", []) or []:
        by_corr.setdefault(k.name + str(k.duration), e)
prev_end = None
agg = collections.OrderedDict()
for k in kern:
    name = k.name
    short = name.replace("void ", "").replace("at::native::", "").replace("(anonymous namespace)::", "").split("(")[0][:70]
    gap = 0 if prev_end is None else k.time_range.start - prev_end
    prev_end = k.time_range.end
    dur = k.time_range.end - k.time_range.start
    print(f"{dur:9.1f} us  gap {gap:7.1f}  {short}")
print()
# aten ops that launched kernels, with stack top inside the package
cnt = collections.Counter(); tim = collections.Counter()
for e in cpu_ops:
    ks = getattr(e, "kernels", []) or []
    if not ks: continue
    stack = [s for s in (e.stack or []) if "instant_angelo_b200" in s or "bench.py" in s]
    where = stack[0].split("/")[-1] if stack else "?"
    key = (e.name, where)
    cnt[key] += len(ks); tim[key] += sum(k.duration for k in ks)
print("== aten ops by launches (op, innermost package frame) ==")
for key, c in cnt.most_common(80):
    print(f"{c:4d} launches {tim[key]:9.1f} us  {key[0]:40s} {key[1]}")
