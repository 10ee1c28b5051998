"""torch.profiler view of one training step of the bench workload (which torch ops make up the glue)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench
from torch.profiler import profile, ProfilerActivity

args = argparse.Namespace(mlp="tc", rays=8192, steps=3, warmup=3)
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
batches = [(b.to(dev), g.to(dev)) for b, g in bench.make_batches(6, 8192, 0, pin=False)]
gs = bench.GLOBAL_STEP0 + 1
for i in range(3):
    b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(3, 5):
        b, bg = bench.unpack_batch(*batches[i]); bench.train_step(cfg, model, arena, var_arena, opt, opt_var, b, bg, gs, 1); gs += 1
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=60))
