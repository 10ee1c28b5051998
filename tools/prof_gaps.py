"""Where the GPU idles inside one training step: kernel timeline of one profiled step, gaps between consecutive kernels
grouped by the kernel that FOLLOWS the gap (i.e. what the GPU was waiting for)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench
from torch.profiler import profile, ProfilerActivity

args = argparse.Namespace(mlp="tc", rays=8192, steps=3, warmup=3, grad_type="finite_difference")
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
batches = [(b.to(dev), g.to(dev)) for b, g in bench.make_batches(10, 8192, 0, pin=False)]
gs = bench.GLOBAL_STEP0 + 1
def step(i):
    global gs
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
for i in range(5): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(5, 8): step(i)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# middle step only: between the 1st and 2nd adamw pair boundaries
adam = [i for i, e in enumerate(ev) if "adamw" in e.name]
lo, hi = adam[1] + 1, adam[3] + 1          # two adamw launches per step
ev = ev[lo:hi]
t0, t1 = ev[0].time_range.start, ev[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in ev)
print(f"step span {(t1 - t0) / 1e3:.2f} ms, kernel busy {busy / 1e3:.2f} ms, idle {(t1 - t0 - busy) / 1e3:.2f} ms, {len(ev)} launches")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    g = b.time_range.start - a.time_range.end
    gaps.append((g, a.name[:60], b.name[:60], (a.time_range.end - t0) / 1e3))
hist = collections.Counter()
for g, *_ in gaps:
    hist["<2us" if g < 2 else "2-5us" if g < 5 else "5-20us" if g < 20 else "20-100us" if g < 100 else ">100us"] += g
print("idle by gap size (us):", {k: round(v) for k, v in hist.items()})
print("largest gaps: gap_us | at ms | after kernel -> before kernel")
for g, a, b, at in sorted(gaps, reverse=True)[:25]:
    print(f"{g:8.1f} | {at:6.2f} | {a}  ->  {b}")
