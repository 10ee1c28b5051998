"""Host-side cost of one training step: wall time of the Python enqueue path (cProfile)."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench

RAYS = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
args = argparse.Namespace(mlp="tc", rays=RAYS, steps=3, warmup=3)
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
batches = [(b.to(dev), g.to(dev)) for b, g in bench.make_batches(12, RAYS, 0, pin=False)]
gs = bench.GLOBAL_STEP0 + 1
for i in range(4):
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
torch.cuda.synchronize()
# (1) enqueue time with the GPU kept busy (no sync between steps except the marching D2H)
t0 = time.perf_counter()
for i in range(4, 8):
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"4 steps: host loop {1e3*(t1-t0)/4:.1f} ms/step, + final sync {1e3*(t2-t1):.1f} ms total")
# (2) host-only cost: sync first so no call ever waits on the GPU queue... the marching sync still waits for the GPU
pr = cProfile.Profile()
pr.enable()
for i in range(8, 12):
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(40)
pstats.Stats(pr, stream=s).sort_stats("cumtime").print_stats(60)
print(s.getvalue()[:30000])
