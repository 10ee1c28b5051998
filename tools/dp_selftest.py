"""Multi-GPU correctness of the ray-sharded data-parallel path on the REAL kernels (reference contract: Lightning DDP,
launch.py:98 -- every rank renders its shard of the ray batch, parameter gradients are averaged over ranks, occupancy
buffers are identical on every rank).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_selftest.py

Checks, on the bench workload (neuralangelo-colmap_sparse, finite differences, tensor-core MLPs, full-size tables):
  1. gradient arena after `ParamArena.all_reduce()` x 1/G  ==  the mean over shards of the gradients ONE rank computes when it
     renders every shard itself (same weights, same random draws): DDP's mean-gradient semantics, to 1e-5 of the tensor scale
     (the scatter kernels use float atomics, so bit equality is not defined);
  2. every rank holds bit-identical parameters after the fused AdamW step that follows;
  3. occupancy grids (float occupancies, boolean grid, packed bitfield; foreground and background) are bit-identical on every
     rank after two refreshes driven by the per-rank, identically seeded generators (no broadcast on this path).
Rank 0 prints one JSON line {"dp_selftest": "ok", ...} and exits 0; any mismatch raises.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1024, help="rays per rank")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=device)
    from instant_angelo_b200.losses import training_loss

    args = argparse.Namespace(mlp="tc", rays=a.rays, grad_type="finite_difference", config="sparse")
    cfg, model, arena, var_arena, opt, opt_var = bench.build_b200(args, rank, world, device)
    gs = bench.GLOBAL_STEP0
    model.update_step(0, gs, update_occupancy=False)

    # ---- identical inputs on every rank: the G shards of one global batch, with every random draw explicit
    shards = []
    g = torch.Generator().manual_seed(1234)
    for r in range(world):
        (buf, bg), = bench.make_batches(1, a.rays, 100 + r, pin=False)
        shards.append((buf.to(device), bg.to(device), torch.rand(a.rays, generator=g).to(device), torch.rand(a.rays, generator=g).to(device)))

    def shard_grads(r, rnd=None):
        buf, bg, u_fg, u_bg = shards[r]
        batch, bgc = bench.unpack_batch(buf, bg)
        model.background_color = bgc
        arena.zero_grad()
        var_arena.zero_grad()
        if rnd is None:     # curvature directions are per marched sample: march once to learn the count, then draw
            with torch.no_grad():
                probe = model(batch["rays"], stratified_u=u_fg, stratified_u_bg=u_bg)
            n_s = int(probe["sdf_samples"].shape[0])
            rnd = torch.randn(n_s, 3, generator=torch.Generator().manual_seed(77 + r)).to(device)
        out = model(batch["rays"], stratified_u=u_fg, rand_directions=rnd, stratified_u_bg=u_bg)
        terms = training_loss(model, out, batch, cfg.system.loss, gs)
        terms["loss"].backward()
        torch.cuda.synchronize()
        return arena.grad.clone(), var_arena.grad.clone(), float(terms["loss"].detach()), rnd

    # ---- reference on every rank: all shards rendered locally, gradients averaged (DDP mean)
    ref_main = torch.zeros_like(arena.grad)
    ref_var = torch.zeros_like(var_arena.grad)
    for r in range(world):
        gm, gv, _, _ = shard_grads(r)
        ref_main += gm / world
        ref_var += gv / world

    # ---- the data-parallel path: own shard only, then one all-reduce over the arena
    shard_grads(rank)
    arena.all_reduce()
    var_arena.all_reduce()
    torch.cuda.synchronize()
    got_main, got_var = arena.grad / world, var_arena.grad / world

    def rel(got, want):
        return float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))

    e_main, e_var = rel(got_main, ref_main), rel(got_var, ref_var)
    assert float(ref_main.abs().max()) > 0 and e_main < 1e-5 and e_var < 1e-5, (rank, e_main, e_var)
    # per segment as well (tables dominate the arena's max): every parameter tensor to 1e-4 of its own scale
    worst = 0.0
    for p, off in zip(arena.params, arena.offsets):
        sl = slice(off, off + p.numel())
        if float(ref_main[sl].abs().max()) > 0:
            worst = max(worst, rel(got_main[sl], ref_main[sl]))
    assert worst < 1e-4, (rank, worst)

    # ---- optimizer step on the reduced gradients: replicas stay bit-identical
    opt.step(gs, grad_scale=1.0 / world)
    opt_var.step(gs, grad_scale=1.0 / world)
    torch.cuda.synchronize()

    def same_everywhere(t, what):
        t = t.contiguous()
        flat = t.view(torch.uint8).reshape(-1) if t.dtype != torch.bool else t.to(torch.uint8).reshape(-1)
        bufs = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(bufs, flat)
        for r, b in enumerate(bufs):
            assert torch.equal(b, bufs[0]), f"{what}: rank {r} differs from rank 0"

    same_everywhere(arena.data, "parameters after the AdamW step")
    same_everywhere(var_arena.data, "variance after the AdamW step")

    # ---- occupancy refreshes: identically seeded per-rank generators keep the grids replica-identical
    with torch.no_grad():
        model.update_step(0, gs + (16 - gs % 16) % 16)          # a refresh step (multiple of 16)
        model.update_step(0, gs + (16 - gs % 16) % 16 + 16)
    torch.cuda.synchronize()
    for name in ("occupancy_grid", "occupancy_grid_bg"):
        grid = getattr(model, name)
        same_everywhere(grid.occs, f"{name}.occs")
        same_everywhere(grid._binary, f"{name}._binary")
        same_everywhere(grid.bitfield, f"{name}.bitfield")
        assert 0.0 < float(grid._binary.float().mean()) < 1.0

    if rank == 0:
        print(json.dumps({"dp_selftest": "ok", "world_size": world, "rays_per_rank": a.rays,
                          "arena_numel": int(arena.numel), "grad_rel_err_vs_mean_of_shards": e_main,
                          "worst_per_parameter_rel_err": worst, "variance_grad_rel_err": e_var,
                          "checked": ["all-reduced arena gradient == mean of shard gradients", "parameters bit-identical after AdamW",
                                      "occupancy occs/_binary/bitfield bit-identical after 2 refreshes"]}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
