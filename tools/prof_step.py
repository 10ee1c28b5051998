"""torch.profiler view of the bench training step: how the step splits between libia_b200 kernels, torch glue
kernels, and GPU idle time (host-bound gaps).  Usage: python tools/prof_step.py [--rays 8192]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench
from torch.profiler import profile, ProfilerActivity

ap = argparse.ArgumentParser(); ap.add_argument("--rays", type=int, default=8192); ap.add_argument("--rows", type=int, default=40)
a = ap.parse_args()
args = argparse.Namespace(mlp="tc", rays=a.rays, steps=3, warmup=3)
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
batches = [(b.to(dev), g.to(dev)) for b, g in bench.make_batches(12, a.rays, 0, pin=False)]
gs = bench.GLOBAL_STEP0 + 1
def step(i):
    global gs
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
for i in range(4): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(4, 8): step(i)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 4 * 1e3
# host-only time to enqueue one step (no sync inside except the two sample-count reads)
NS = 2
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(8, 8 + NS): step(i)
    torch.cuda.synchronize()
ka = prof.key_averages()
mine = glue = 0.0; n_mine = n_glue = 0; rows = []
for e in ka:
    if e.device_type == torch.autograd.DeviceType.CUDA:
        t = e.self_device_time_total / 1e3 / NS
        name = e.key
        is_glue = ("at::" in name) or ("at_cuda_detail" in name) or ("cub::" in name) or name.startswith("Memcpy") or name.startswith("Memset") or "nccl" in name
        if is_glue: glue += t; n_glue += e.count / NS
        else: mine += t; n_mine += e.count / NS
        rows.append((t, e.count / NS, ("glue " if is_glue else "ia   ") + name[:110]))
rows.sort(reverse=True)
print(f"step wall (no profiler) {wall:.2f} ms | per step: libia kernels {mine:.2f} ms ({n_mine:.0f} launches), torch glue {glue:.2f} ms ({n_glue:.0f} launches), "
      f"GPU idle ~{wall - mine - glue:.2f} ms")
for t, c, n in rows[:a.rows]:
    print(f"{t:8.3f} ms {c:6.0f}x  {n}")

if os.environ.get("IA_PROF_SHAPES"):
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof2:
        step(10)
        torch.cuda.synchronize()
    ev = [e for e in prof2.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") or "Backward" in e.key]
    print("\n== torch ops by device time (one step), with shapes ==")
    for e in sorted(ev, key=lambda e: -e.self_device_time_total)[:60]:
        print(f"{e.self_device_time_total/1e3:8.3f} ms {e.count:4d}x cpu {e.self_cpu_time_total/1e3:7.3f} ms  {e.key:34s} {str(e.input_shapes)[:120]}")
    print("\n== torch ops by call count ==")
    for e in sorted(ev, key=lambda e: -e.count)[:40]:
        print(f"{e.count:4d}x dev {e.self_device_time_total/1e3:8.3f} ms cpu {e.self_cpu_time_total/1e3:7.3f} ms  {e.key:34s} {str(e.input_shapes)[:120]}")
