"""Which steps of the bench loop stall, and what changed on that step: CUDA-event step time, host time, marched samples,
caching-allocator segments / reserved bytes (growth = cudaMalloc inside the step)."""
import sys, os, time, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--config", default="sparse"); ap.add_argument("--steps", type=int, default=50)
a = ap.parse_args()
args = argparse.Namespace(mlp="tc", rays=8192, steps=a.steps, warmup=5, grad_type="finite_difference", config=a.config)
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
K, W = a.steps, 5
batches = [(b.to(dev), g.to(dev)) for b, g in bench.make_batches(K + W, 8192, 0, pin=True)]
gs = bench.GLOBAL_STEP0
big, big_bg = bench.make_batches(1, 8192 + 2048, 7919, pin=False)[0]
b, bg = bench.unpack_batch(big.to(dev), big_bg.to(dev))
bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs - 1, 1)
for i in range(W):
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
gc.collect(); gc.freeze(); gc.disable()
torch.cuda.synchronize()
rows = []
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
ev[0].record()
for i in range(W, W + K):
    t0 = time.perf_counter()
    b, bg = bench.unpack_batch(*batches[i])
    loss, out = bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
    ev[i - W + 1].record()
    st = torch.cuda.memory_stats()
    rows.append((time.perf_counter() - t0, model.last_num_samples, model.last_num_samples_full, st["segment.all.current"],
                 st["reserved_bytes.all.current"] >> 20, st["num_alloc_retries"], st["allocation.all.allocated"]))
torch.cuda.synchronize()
prev_alloc = None
for i, r in enumerate(rows):
    ms = ev[i].elapsed_time(ev[i + 1])
    flag = "  <<<<" if ms > 40 else ""
    print(f"step {i:3d} dev {ms:7.2f} ms host {1e3 * r[0]:7.2f} ms fg {r[1]:8d} full {r[2]:8d} segments {r[3]:4d} reserved {r[4]:7d} MB retries {r[5]} "
          f"allocs/step {r[6] - (prev_alloc or r[6])}{flag}")
    prev_alloc = r[6]
