"""Ray-sharded data parallelism for the hot path.

Reference: Lightning DDP (`strategy='ddp_find_unused_parameters_false'`, launch.py:98) all-reduces every
parameter gradient (mean) in 25 MB buckets and broadcasts the occupancy buffers each forward.  Here: one
process per GPU, parameters AND gradients live in two flat fp32 arenas (hash tables first), each rank renders
its own shard of the ray batch, and ONE NCCL all-reduce(sum) over the gradient arena follows backward; the
1/world_size of DDP's mean is folded into the fused AdamW step.  Occupancy grids stay replica-identical
because every rank refreshes them with the same seed -- no broadcast."""
from __future__ import annotations

from typing import Iterable, List

import torch
import torch.distributed as dist
import torch.nn as nn


class ParamArena:
    """Re-homes the parameters of `modules` into one contiguous fp32 buffer (and their .grad into another).
    Works on any device, so the gloo/CPU tests exercise the same code as NCCL on B200."""

    def __init__(self, params: Iterable[nn.Parameter]):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad and p.numel() > 0]
        assert self.params, "no parameters"
        dev = self.params[0].device
        # 16-byte align every segment so the fused AdamW kernel can use float4 on the whole arena
        self.offsets, n = [], 0
        for p in self.params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.numel = n
        self.data = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.data[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off:off + p.numel()].view_as(p)
                # kernels whose gradient is a scatter (the hash tables) accumulate into this slice directly (ops._grad_sink)
                p._ia_grad_inplace = True

    def zero_grad(self) -> None:
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):   # autograd may have replaced .grad; re-point it
            g = self.grad[off:off + p.numel()].view_as(p)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                p.grad = g

    def pin_tables_in_l2(self, max_bytes: int = 96 << 20) -> dict:
        """Persisting L2 window over the leading run of large parameters (the hash tables come first in the arena): the
        gathers of a step then hit the L2 set-aside instead of re-fetching table lines that the [N, L*F] activation
        streams evicted.  The window is clipped to `max_bytes` and to what the device grants; the persisting fraction is
        scaled so that the set-aside is not over-subscribed."""
        from . import ops
        if not self.data.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        tables = [(off, p.numel()) for p, off in zip(self.params, self.offsets) if p.numel() >= (1 << 20)]
        if not tables:
            return {}
        lo = tables[0][0]
        hi = lo
        for off, n in tables:                 # contiguous prefix of tables that fits the budget
            if (off + n - lo) * 4 > max_bytes and hi > lo:
                break
            hi = off + n
        window = self.data[lo:hi]
        probe = ops.l2_persist(window, 1.0)
        ratio = min(1.0, probe["set_aside_bytes"] / max(probe["window_bytes"], 1))
        if ratio < 1.0:
            probe = ops.l2_persist(window, ratio)
        probe["hit_ratio"] = ratio
        return probe

    def all_reduce(self, group=None) -> None:
        """Sum over ranks (the mean's 1/G is applied by the optimizer)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)

    def broadcast_params(self, src: int = 0, group=None) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(self.data, src=src, group=group)


class FusedAdamW:
    """AdamW(lr, betas=(0.9, 0.99), eps=1e-15, weight_decay=0.01) over a ParamArena with the reference's LR
    schedule: LinearLR(0.01 -> 1, warmup_steps) then ExponentialLR(gamma) (configs/...sparse.yaml:134-165)."""

    def __init__(self, arena: ParamArena, lr: float = 0.01, betas=(0.9, 0.99), eps: float = 1e-15, weight_decay: float = 0.01,
                 warmup_steps: int = 500, max_steps: int = 20000, decay_factor: float = 0.1, schedule=None):
        """`schedule(t) -> factor` (systems.parse_scheduler) replaces the built-in warm-up + exponential decay."""
        self.arena, self.lr, self.betas, self.eps, self.wd = arena, lr, betas, eps, weight_decay
        self.schedule = schedule
        self.m = torch.zeros_like(arena.data)
        self.v = torch.zeros_like(arena.data)
        self.t = 0
        self.warmup_steps = warmup_steps
        self.gamma = decay_factor ** (1.0 / max(max_steps - warmup_steps, 1))

    def lr_at(self, step: int) -> float:
        if self.schedule is not None:
            return self.lr * float(self.schedule(step))
        if step < self.warmup_steps:
            return self.lr * (0.01 + (1.0 - 0.01) * step / self.warmup_steps)
        return self.lr * self.gamma ** (step - self.warmup_steps)

    def step(self, global_step: int, grad_scale: float = 1.0) -> None:
        self.t += 1
        lr = self.lr_at(global_step)
        a = self.arena
        if not a.data.is_cuda:
            raise NotImplementedError("Only support cuda inputs.")
        from . import ops
        ops.adamw_step(a.data, a.grad, self.m, self.v, lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t, grad_scale)


def shard_rays(n_global: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a global ray batch."""
    per = (n_global + world - 1) // world
    lo = min(rank * per, n_global)
    return lo, min(lo + per, n_global)
