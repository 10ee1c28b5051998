import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, argparse
import bench
from torch.profiler import profile, ProfilerActivity
args = argparse.Namespace(mlp="tc", rays=8192, steps=3, warmup=3, grad_type="finite_difference")
dev = torch.device("cuda", 0)
cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, 0, 1, dev)
torch.cuda.synchronize()
for gs in (19008, 19024):
    t0 = time.perf_counter(); model.update_step(0, gs); torch.cuda.synchronize(); print("refresh", gs, 1e3 * (time.perf_counter() - t0), "ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    model.update_step(0, 19040); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=50))
