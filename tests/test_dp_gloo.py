"""CPU, world_size 2 over gloo: the data-parallel plumbing (instant_angelo_b200/dp.py) -- parameter/gradient arenas,
ray sharding, all-reduce(sum) + 1/G folding, parameter broadcast.  The compute kernels are not involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from instant_angelo_b200.dp import ParamArena, shard_rays
    torch.manual_seed(100 + rank)                     # deliberately different initial weights per rank
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    table = torch.nn.Parameter(torch.randn(1001))     # odd size: exercises the 16-byte segment padding
    arena = ParamArena([table] + list(net.parameters()))
    arena.broadcast_params(0)
    # every parameter is now a view into the arena and identical on all ranks
    assert all(p.data_ptr() >= arena.data.data_ptr() for p in arena.params)
    ref = [torch.zeros_like(arena.data) for _ in range(world)]
    dist.all_gather(ref, arena.data)
    assert torch.equal(ref[0], ref[1])
    # a global batch of 10 "rays", sharded contiguously; loss = sum over the shard
    g = torch.Generator().manual_seed(7)
    x = torch.randn(10, 5, generator=g)
    lo, hi = shard_rays(10, rank, world)
    arena.zero_grad()
    loss = (net(x[lo:hi]).sum() + (table[:10] * x[lo:hi].sum()).sum())
    loss.backward()
    assert table.grad.data_ptr() == arena.grad.data_ptr(), "autograd must accumulate into the arena view"
    arena.all_reduce()
    # single-process reference on the whole batch
    torch.manual_seed(100)
    net0 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    table0 = torch.nn.Parameter(torch.randn(1001))
    loss0 = sum((net0(x[a:b]).sum() + (table0[:10] * x[a:b].sum()).sum()) for a, b in [shard_rays(10, r, world) for r in range(world)])
    loss0.backward()
    ok = torch.allclose(table.grad, table0.grad, atol=1e-5) and all(
        torch.allclose(p.grad, p0.grad, atol=1e-5) for p, p0 in zip(net.parameters(), net0.parameters()))
    q.put((rank, bool(ok), (lo, hi)))
    dist.destroy_process_group()


def test_dp_arena_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert [r[1] for r in res] == [True, True]
    assert [r[2] for r in res] == [(0, 5), (5, 10)]


def _worker_system(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from instant_angelo_b200 import configs
    from instant_angelo_b200.synthetic import SphereDataset
    from instant_angelo_b200.systems import NeuSSystem
    cfg = configs.neuralangelo_colmap_sparse("finite_difference")
    for blk in (cfg.model.geometry, cfg.model.geometry_bg):
        blk.xyz_encoding_config["log2_hashmap_size"] = 8
    torch.manual_seed(100 + rank)                     # replicas start from different weights ...
    system = NeuSSystem(cfg, dataset=SphereDataset(3, 16, 12, focal=10.0, n_points=50), device_sampling=True)
    groups = system.configure_optimizers()
    groups.broadcast_params(0)                        # ... and agree after the broadcast of both arenas
    same = True
    for a in groups.arenas:
        got = [torch.zeros_like(a.data) for _ in range(world)]
        dist.all_gather(got, a.data)
        same &= torch.equal(got[0], got[1])
    # per-rank ray streams (seed + 1000 * rank), so the ranks render different shards of the image set
    system.seed_everything(42, rank)
    b = {}
    system.preprocess_data(b, "train")
    rays = [torch.zeros_like(b["rays"]) for _ in range(world)]
    dist.all_gather(rays, b["rays"].contiguous())
    distinct = not torch.equal(rays[0], rays[1])
    # one all-reduce(sum) per arena; the mean's 1/world is the optimizer's grad_scale
    groups.zero_grad()
    for a in groups.arenas:
        a.grad.fill_(float(rank + 1))
    groups.all_reduce()
    summed = all(bool((a.grad == 3.0).all()) for a in groups.arenas)
    views = all(p.grad.data_ptr() == a.grad[off:].data_ptr() for a in groups.arenas for p, off in zip(a.params, a.offsets))
    q.put((rank, bool(same), bool(distinct), bool(summed), bool(views), len(groups.arenas)))
    dist.destroy_process_group()


def test_system_optimizer_groups_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_system, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=10) for _ in range(2))
    assert [r[1:] for r in res] == [(True, True, True, True, 2)] * 2


def test_shard_and_lr_schedule():
    from instant_angelo_b200.dp import FusedAdamW, ParamArena, shard_rays
    assert [shard_rays(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    arena = ParamArena([torch.nn.Parameter(torch.zeros(8))])
    opt = FusedAdamW(arena, lr=0.01, warmup_steps=500, max_steps=20000)
    assert abs(opt.lr_at(0) - 1e-4) < 1e-12 and abs(opt.lr_at(500) - 0.01) < 1e-12
    assert abs(opt.lr_at(20000) - 0.001) < 1e-9          # ExponentialLR decays by 0.1 over max_steps - warmup
    with pytest.raises(NotImplementedError, match="cuda"):
        opt.step(0)                                      # no CPU fallback for the fused optimizer
