#!/usr/bin/env python
"""bench.py -- train rays/s of the Instant-angelo hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one pass of the training hot path over one batch of synthetic rays: occupancy refresh when due
(every 16th step), occupancy-grid ray marching, VolumeSDF (1 + 6 finite-difference + 6 curvature taps through
the hash grid and the width-64 MLP), colour head, NeuS alpha compositing, background NeRF++ branch, the losses
of systems/neus.py:130-194, backward, gradient all-reduce over NCCL when N > 1 and a fused AdamW step.

Workload (config.workload): BASELINE.json configs[1] -- configs/neuralangelo-colmap_sparse.yaml with
model.geometry.grad_type=finite_difference on synthetic 512x512 cameras around an analytic sphere, 8192 rays per
GPU per step (max_train_num_rays), 512 (+256 background) samples/ray budget, all 16 hash levels active.

Rank 0 prints ONE JSON line.  `value` is measured with the ray batches already resident in HBM; `e2e` repeats
the measurement with every step's rays copied from pinned host memory and the loss read back.
`--impl reference`: the reference has no CPU implementation and its CUDA dependencies (tinycudann, nerfacc) are
not installable in this image, so the reference arm times the CPU oracle restatement (oracle/) of the same
workload on all host cores, on a bounded sample of rays per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# Marched sample counts change every step, so every activation tensor has a new size every step: with exact-size blocks a
# record-high count finds no cached block and the pool grows inside the timed region -- cudaMalloc / cuMemMap stalls of
# 30-600 ms on this (virtualised) box, at deterministic step indices (tools/diag_stalls.py).  Sizes rounded to eighths of a
# power of two land in the same bucket step after step (measured: 0 stalls in 6 runs, against 3 in 6 with expandable segments).
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "roundup_power2_divisions:8")
import torch  # noqa: E402

GLOBAL_STEP0 = 19000          # all 16 levels active (start_level 4 + (19000-5000)//1000 >= 16), curvature weight 0.5
RAYS_PER_GPU = 8192           # max_train_num_rays of the config
OCC_WARMUP_UPDATES = 16
L2_WINDOW = None              # what ParamArena.pin_tables_in_l2 was granted (--l2-persist), for the bench line
AR_EVENTS = None              # list of (start, end) CUDA events around the gradient all-reduce while the timed region runs


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe), through NVML in a
    background thread (the nvidia-smi CLI in loop mode re-enumerates every GPU on each tick and measurably stalls
    kernel launches of the process being measured)."""

    def __init__(self, gpu_index: int, period_s: float = float(os.environ.get("IA_CLOCK_PERIOD_S", "0.2"))):
        self.idx, self.period = gpu_index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = None
        self._thread = None

    def _loop(self, nv, handle):
        masks = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(handle)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, m in masks.items():
                    if r & m:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def start(self):
        import threading
        try:
            import pynvml as nv
            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES when mapping the CUDA ordinal to an NVML index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            nvml_idx = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].strip().isdigit() else self.idx
            handle = nv.nvmlDeviceGetHandleByIndex(nvml_idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self._stop = threading.Event()
            self._thread = threading.Thread(target=self._loop, args=(nv, handle), daemon=True)
            self._thread.start()
        except Exception as e:       # NVML missing: report that no clock record exists rather than guessing
            self._thread = None
            self.error = str(e)

    def mark(self):
        """Start of the timed region: forget what was sampled during warm-up (NVML initialisation and its first
        queries stall the driver for ~100 ms, so the sampler is started before the warm-up steps)."""
        self.samples, self.reasons = [], set()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread is None:
            out["error"] = getattr(self, "error", "sampler not started")
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            sm = sorted(self.samples)
            out.update({"sm_mhz": sm[len(sm) // 2], "reasons": sorted(self.reasons), "samples": len(sm)})
        return out


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
WORKLOADS = {
    # --config: (BASELINE.json configs[] index, reference YAML, what differs from configs[1])
    "sparse": (1, "neuralangelo-colmap_sparse.yaml", "2^19-entry tables, 512 (+256 background) samples/ray budget"),
    "dense21": (2, "neuralangelo-colmap_dense.yaml", "2^21-entry tables (198 MB per grid: not L2 resident), 256 samples/ray, "
                                                      "dual-colour background head"),
    "wreflection": (3, "neuralangelo-colmap_sparse-wreflection.yaml", "UniSDF V3 colour heads (two 87->64->64->3 heads + 71->64->1 "
                                                                      "weight net, SH degree 3), 1024 samples/ray"),
}


def _set_mlp_otype(node, otype):
    for k in list(node.keys()):
        v = node[k]
        if hasattr(v, "keys"):
            if "n_neurons" in v and "otype" in v:
                v["otype"] = otype
            else:
                _set_mlp_otype(v, otype)


def workload_config(args):
    """The reference configuration `--config` names, with model.geometry.grad_type and the MLP arithmetic of the arm."""
    from instant_angelo_b200 import configs
    name = getattr(args, "config", "sparse")
    grad_type = getattr(args, "grad_type", "finite_difference")
    if name == "sparse":
        cfg = configs.neuralangelo_colmap_sparse(grad_type)
    elif name == "dense21":
        cfg = configs.neuralangelo_colmap_dense(grad_type, log2_hashmap_size=21)
        cfg.model.num_samples_per_ray = 256
    elif name == "wreflection":
        cfg = configs.neuralangelo_colmap_sparse_wreflection(grad_type)
    else:
        raise ValueError(name)
    _set_mlp_otype(cfg.model, "FullyFusedMLP" if getattr(args, "mlp", "tc") == "tc" else "VanillaMLP")
    return cfg


def workload_description(args, cfg, world):
    """`config` of the JSON line: static facts of the workload only, identical for the B200 arm and the reference arm."""
    name = getattr(args, "config", "sparse")
    idx, yaml, what = WORKLOADS[name]
    return {"workload": f"{yaml}, grad_type={args.grad_type}, synthetic 512x512 cameras around an analytic sphere "
                        f"(BASELINE.json configs[{idx}]: {what})",
            "config_name": name, "rays_per_gpu_per_step": args.rays, "global_rays_per_step": args.rays * world,
            "num_samples_per_ray": int(cfg.model.num_samples_per_ray), "num_samples_per_ray_bg": int(cfg.model.num_samples_per_ray_bg),
            "hash_levels_active": 16, "log2_hashmap_size": int(cfg.model.geometry.xyz_encoding_config.log2_hashmap_size),
            "global_step": GLOBAL_STEP0, "grad_type": args.grad_type,
            "l2": "per-step working set (hash tables + >1 GB of per-sample activations) exceeds the 126 MB L2; no explicit flush"}


def build_b200(args, rank, world, device):
    from instant_angelo_b200 import make
    from instant_angelo_b200.dp import FusedAdamW, ParamArena

    torch.manual_seed(42)
    cfg = workload_config(args)
    model = make("neus", cfg.model).to(device)
    model.train()
    for grid in (model.occupancy_grid, model.occupancy_grid_bg):
        grid.generator = torch.Generator(device=device).manual_seed(4242)      # identical on every rank
    with torch.no_grad():
        for i in range(OCC_WARMUP_UPDATES):
            model.update_step(0, 16 * i)
    main_params = [p for n, p in model.named_parameters() if not n.startswith("variance.")]
    arena = ParamArena(main_params)
    var_arena = ParamArena(list(model.variance.parameters()))
    if world > 1:
        arena.broadcast_params(0)
        var_arena.broadcast_params(0)
    opt = FusedAdamW(arena, lr=0.01)
    opt_var = FusedAdamW(var_arena, lr=0.001)
    global L2_WINDOW
    L2_WINDOW = arena.pin_tables_in_l2(args.l2_persist << 20) if getattr(args, "l2_persist", 0) > 0 else None
    return cfg, model, arena, var_arena, opt, opt_var


def make_batches(n_batches, n_rays, rank, pin):
    from instant_angelo_b200.synthetic import SphereScene
    scene = SphereScene(seed=42)
    gen = torch.Generator().manual_seed(42 + 1000 * rank)     # per-rank ray stream (Appendix C-13 neutralised)
    out = []
    for _ in range(n_batches):
        rays, rgb = scene.sample(n_rays, gen)
        pts, nrm, conf = scene.surface_points(n_rays, gen)
        bg = torch.rand(3, generator=gen)
        # one pinned staging buffer per batch: rays(6) rgb(3) pts(3) nrm(3) conf(1) = 16 floats per ray, + bg colour
        buf = torch.cat([rays, rgb, pts, nrm, conf[:, None]], dim=1).contiguous()
        if pin:
            buf, bg = buf.pin_memory(), bg.pin_memory()
        out.append((buf, bg))
    return out


def unpack_batch(buf, bg):
    return {"rays": buf[:, 0:6], "rgb": buf[:, 6:9], "pts": buf[:, 9:12].contiguous(), "pts_normal": buf[:, 12:15],
            "pts_weights": buf[:, 15]}, bg


def train_step(cfg, model, arena, var_arena, opt, opt_var, batch, bg, gs, world):
    from instant_angelo_b200.losses import training_loss
    model.update_step(0, gs)                                   # levels/eps/cos-anneal + occupancy refresh when gs % 16 == 0
    model.background_color = bg
    arena.zero_grad()
    var_arena.zero_grad()
    out = model(batch["rays"])
    terms = training_loss(model, out, batch, cfg.system.loss, gs)
    terms["loss"].backward()
    if world > 1:
        if AR_EVENTS is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        arena.all_reduce()
        var_arena.all_reduce()
        if AR_EVENTS is not None:
            b.record()
            AR_EVENTS.append((a, b))
    opt.step(gs, grad_scale=1.0 / world)
    opt_var.step(gs, grad_scale=1.0 / world)
    return terms["loss"], out


def mlp_flops_of(module) -> float:
    """2*MAC of one forward evaluation of every width-64 network under `module` (unpadded)."""
    from instant_angelo_b200.network_utils import VanillaMLP
    total = 0.0
    for m in module.modules():
        if isinstance(m, VanillaMLP):
            total += 2.0 * (m.n_input_dims * 64 + (64 * 64 if m.n_hidden_layers == 2 else 0) + 64 * m.n_output_dims)
    return total


def step_roofline(fg_per_ray, bg_per_ray, rays_per_s, peaks, grad_type="finite_difference", colour_flop=None, bg_flop=None,
                  tables_l2_resident=True):
    """Whole-step roofline of SURVEY.md section 8d: rays/s <= min(HBM bytes/s / algorithmic bytes per ray,
    tensor FLOP/s / algorithmic FLOP per ray), from the measured samples per ray.  Algorithmic work per sample (fp32
    tables and encodings, 16 levels x 2 features; unpadded 2*MAC of the networks as this path evaluates them: the 12 tap
    evaluations compute the SDF column only; backward = 2x forward):
      foreground sample, finite differences: 13 encodes (1164 B forward + 1164 B table gradient each) + 6 input gradients
        (1176 B: the curvature taps move with the parameters); 35>64>64>65 centre + 12 x 35>64>64>1 taps + 87>64>64>3 colour
      foreground sample, analytic: 7 encodes, 1 input gradient + its two second-order adjoints, centre network twice over
      background sample: 1 encode (forward + table gradient); 35>64>8 density + 24>64>64>3 colour."""
    enc_f, enc_bt, enc_bi = 1164.0, 1164.0, 1176.0
    mlp = lambda din, nh, nout: 2.0 * (din * 64 + (64 * 64 if nh == 2 else 0) + 64 * nout)
    centre, tap, colour = mlp(35, 2, 65), mlp(35, 2, 1), (mlp(87, 2, 3) if colour_flop is None else colour_flop)
    if grad_type == "finite_difference":
        fg_bytes = 13 * (enc_f + enc_bt) + 6 * enc_bi
        fg_flop = 3.0 * (centre + 12 * tap + colour)
    else:
        fg_bytes = 7 * (enc_f + enc_bt) + 6 * enc_bi + (enc_bi + enc_f + enc_bt)
        fg_flop = 3.0 * (2 * centre + 6 * tap + colour)
    bg_bytes = enc_f + enc_bt
    bg_flop = 3.0 * ((mlp(35, 1, 8) + mlp(24, 2, 3)) if bg_flop is None else bg_flop)
    bytes_per_ray = fg_per_ray * fg_bytes + bg_per_ray * bg_bytes + 24 + 32
    flop_per_ray = fg_per_ray * fg_flop + bg_per_ray * bg_flop
    by_hbm = peaks["hbm_gbs"] * 1e9 / bytes_per_ray
    by_tensor = peaks["bf16_tflops_sustained"] * 1e12 / flop_per_ray
    bound = "hbm" if by_hbm <= by_tensor else "tensor"
    roof = min(by_hbm, by_tensor)
    return {"bound": bound, "rays_per_s_roofline": roof, "frac": rays_per_s / roof, "algorithmic_bytes_per_ray": bytes_per_ray,
            "algorithmic_flop_per_ray": flop_per_ray, "rays_per_s_by_hbm": by_hbm, "rays_per_s_by_tensor": by_tensor,
            "note": ("per GPU; the 2^19-entry tables are L2 resident, so the HBM line is the SURVEY 8d reporting convention, not a hard bound"
                     if tables_l2_resident else "per GPU; the 2^21-entry tables (198 MB per grid) do not fit the 126 MB L2: the HBM line binds")}


def hashgrid_microbench(device, peaks):
    """Secondary BASELINE metric: hash-grid G point-evals/s (16 levels, F=2, 2^19 entries/level, 2^22 incoherent points)."""
    from instant_angelo_b200 import ops
    plan = ops.make_grid_plan(16, 2, 19, 32, 1.3195079107728942)
    n = 1 << 22
    g = torch.Generator(device=device).manual_seed(7)
    x = torch.rand(n, 3, device=device, generator=g)
    table = torch.randn(plan.n_params, device=device, generator=g) * 0.1
    dy = torch.randn(n, 32, device=device, generator=g)
    dtab = torch.zeros_like(table)
    import ctypes as C
    from instant_angelo_b200 import _lib as L
    lib, s = L.load(), L.stream()
    out = torch.empty(n, 32, device=device)
    dx = torch.empty_like(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2

    def timed(fn, reps=5):
        ts = []
        for _ in range(reps + 2):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts = sorted(ts[2:])
        return ts[len(ts) // 2]

    res = {}
    ms = timed(lambda: lib.ia_hashgrid_fwd(x.data_ptr(), n, table.data_ptr(), C.byref(plan), 16, out.data_ptr(), s))
    b = ops.hashgrid_bytes_per_point(plan, 16, "fwd")
    res["fwd"] = {"ms": ms, "gevals_per_s": n / ms / 1e6, "algorithmic_GBps": n * b / ms / 1e6, "frac_of_hbm": n * b / ms / 1e6 / peaks["hbm_gbs"]}
    ms = timed(lambda: lib.ia_hashgrid_bwd(x.data_ptr(), n, table.data_ptr(), dy.data_ptr(), C.byref(plan), 16, dtab.data_ptr(), None, s))
    b = ops.hashgrid_bytes_per_point(plan, 16, "bwd_table")
    res["bwd_table"] = {"ms": ms, "gevals_per_s": n / ms / 1e6, "algorithmic_GBps": n * b / ms / 1e6, "frac_of_hbm": n * b / ms / 1e6 / peaks["hbm_gbs"]}
    ms = timed(lambda: lib.ia_hashgrid_bwd(x.data_ptr(), n, table.data_ptr(), dy.data_ptr(), C.byref(plan), 16, None, dx.data_ptr(), s))
    b = ops.hashgrid_bytes_per_point(plan, 16, "bwd_input")
    res["bwd_input"] = {"ms": ms, "gevals_per_s": n / ms / 1e6, "algorithmic_GBps": n * b / ms / 1e6, "frac_of_hbm": n * b / ms / 1e6 / peaks["hbm_gbs"]}
    # SURVEY 8d grid: T in {2^19, 2^21} x active levels in {4, 16} x {incoherent, ray-coherent} points, forward
    grid = []
    n_c = 8192 * 512
    o = torch.nn.functional.normalize(torch.randn(8192, 3, device=device, generator=g), dim=-1) * 1.0
    tgt = (torch.rand(8192, 3, device=device, generator=g) - 0.5) * 0.6
    d = torch.nn.functional.normalize(tgt - o, dim=-1)
    tt = torch.linspace(0.0, 2.0, 512, device=device)[None, :, None]
    x_coh = (((o[:, None, :] + d[:, None, :] * tt) / 3.0 + 0.5).clamp(0.0, 1.0)).reshape(-1, 3).contiguous()   # 512 steps along 8192 rays
    out_c = torch.empty(n_c, 32, device=device)
    for log2_t in (19, 21):
        plan_t = ops.make_grid_plan(16, 2, log2_t, 32, 1.3195079107728942)
        table_t = torch.randn(plan_t.n_params, device=device, generator=g) * 0.1
        table_h = ops.table_to_half(table_t)
        for active in (4, 16):
            for label, pts, dst in (("incoherent 2^22 uniform points", x, out), ("ray-coherent 8192 rays x 512 steps", x_coh, out_c)):
                npts = pts.shape[0]
                ms = timed(lambda: lib.ia_hashgrid_fwd(pts.data_ptr(), npts, table_t.data_ptr(), C.byref(plan_t), active, dst.data_ptr(), s))
                bpp = ops.hashgrid_bytes_per_point(plan_t, active, "fwd")
                grid.append({"log2_T": log2_t, "active_levels": active, "points": label, "table": "fp32", "ms": ms, "gevals_per_s": npts / ms / 1e6,
                             "algorithmic_GBps": npts * bpp / ms / 1e6, "frac_of_hbm": npts * bpp / ms / 1e6 / peaks["hbm_gbs"]})
                # the same gathers from the fp16 shadow table (opt-in `table_precision: fp16`), against ITS algorithmic bytes
                ms = timed(lambda: lib.ia_hashgrid_fwd_h(pts.data_ptr(), npts, table_h.data_ptr(), C.byref(plan_t), active, dst.data_ptr(), s))
                bpp = ops.hashgrid_bytes_per_point(plan_t, active, "fwd", param_bytes=2)
                grid.append({"log2_T": log2_t, "active_levels": active, "points": label, "table": "fp16", "ms": ms, "gevals_per_s": npts / ms / 1e6,
                             "algorithmic_GBps": npts * bpp / ms / 1e6, "frac_of_hbm": npts * bpp / ms / 1e6 / peaks["hbm_gbs"]})
        del table_t, table_h
    res["fwd_grid"] = grid
    # L2 / HBM random 32-byte-sector gather peaks: table resident in the 126 MB L2 vs. larger than it
    big = torch.randn((1 << 30) // 4, device=device, generator=g)            # 1 GiB
    sink = torch.empty(1 << 22, device=device)
    gather = {}
    for label, nbytes in (("l2_resident_48MB", 48 << 20), ("hbm_1GiB", 1 << 30)):
        nsec, nthr, iters = nbytes // 32, 1 << 22, 64

        def run_gather():
            assert lib.ia_debug_sector_gather(big.data_ptr(), nsec, nthr, iters, sink.data_ptr(), s) == 0
        run_gather()                                                           # warm the L2 for the resident case
        ts = []
        for _ in range(5):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run_gather(); b_.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b_))
        ms = sorted(ts)[2]
        gather[label] = {"ms": ms, "gsectors_per_s": nthr * iters / ms / 1e6, "sector_GBps": nthr * iters * 32 / ms / 1e6}
    res["sector_gather_peaks"] = gather
    res["config"] = "N=2^22 uniform random points, L=16 F=2 T=2^19 fp32 tables, L2 flushed between launches; bytes/point fwd=%d" % ops.hashgrid_bytes_per_point(plan, 16, "fwd")
    return res


def run_b200(args):
    import torch.distributed as dist
    from instant_angelo_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run"
    peaks = measured_peaks()
    K, W = args.steps, args.warmup
    n_rays = args.rays

    cfg, model, arena, var_arena, opt, opt_var = build_b200(args, rank, world, device)
    host_batches = make_batches(K + W, n_rays, rank, pin=True)
    dev_batches = [(b.to(device), g.to(device)) for b, g in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    def sum_over_ranks(v):
        if world > 1:
            t = torch.tensor([float(v)], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return float(v)

    # ---- kernel-resident measurement (`value`) --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    gs = GLOBAL_STEP0
    # allocator head-room: untimed steps on 25 % and 12.5 % larger ray batches (the next two size buckets of the caching
    # allocator, see PYTORCH_CUDA_ALLOC_CONF above), so that a timed step whose marched sample count exceeds everything seen
    # during warm-up does not grow the CUDA memory pool
    def headroom(step):
        for extra in (n_rays // 4, n_rays // 8):
            big, big_bg = make_batches(1, n_rays + extra, rank + 7919, pin=False)[0]
            batch, bg = unpack_batch(big.to(device), big_bg.to(device))
            train_step(cfg, model, arena, var_arena, opt, opt_var, batch, bg, step, world)
            del big, big_bg, batch, bg

    headroom(gs - 1)
    for i in range(W):
        batch, bg = unpack_batch(*dev_batches[i])
        train_step(cfg, model, arena, var_arena, opt, opt_var, batch, bg, gs, world)
        gs += 1
    # ... and spare 2 MB blocks for the allocator's small pool (per-step scalars, event-sized tensors): its growth inside the
    # timed region is a cudaMalloc too (tools/diag_stalls.py: +2 MB at the second timed step, a 30-80 ms stall on this box)
    spare = [torch.empty(1 << 20, dtype=torch.uint8, device=device) for _ in range(32)] + \
            [torch.empty(8 << 20, dtype=torch.uint8, device=device) for _ in range(16)]      # 1-10 MB requests: 20 MB blocks
    del spare
    prof = ops.PROFILER
    prof.reset()
    prof.enabled, prof.timing = True, rank == 0          # launch counting everywhere, CUDA-event timing on rank 0 only
    PROF_STEPS = min(K, 4)                               # ... and only for the first steps of the timed region
    n_samples_fg = n_samples_full = 0
    import gc
    gc.collect()
    gc.freeze()          # the training loop allocates thousands of short-lived Python objects per step: keep the cyclic
    gc.disable()         # collector from pausing the launch thread inside the timed regions
    barrier()
    sampler.mark()
    global AR_EVENTS
    AR_EVENTS = [] if world > 1 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    counts = []
    step_events = [e0]
    host_t = [time.perf_counter()]
    for i in range(W, W + K):
        batch, bg = unpack_batch(*dev_batches[i])
        loss, out = train_step(cfg, model, arena, var_arena, opt, opt_var, batch, bg, gs, world)
        counts.append((out["num_samples"], out["num_samples_full"]))
        gs += 1
        host_t.append(time.perf_counter())
        if i - W + 1 == PROF_STEPS:
            prof.timing = False
        if i < W + K - 1:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            step_events.append(ev)
    e1.record()
    step_events.append(e1)
    barrier()
    ar_events, AR_EVENTS = AR_EVENTS, None
    # per-step device times of this rank, then (min, median, max) over steps of the per-step MAX over ranks, the spread
    # between ranks, and the all-reduce's share: a straggler, unequal sample counts and a real collective cost look different
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in zip(step_events[:-1], step_events[1:])], device=device, dtype=torch.float64)
    ar_ms = torch.tensor([a.elapsed_time(b) for a, b in (ar_events or [])], device=device, dtype=torch.float64)
    if world > 1:
        all_steps = [torch.empty_like(step_ms) for _ in range(world)]
        dist.all_gather(all_steps, step_ms)
        all_steps = torch.stack(all_steps)                        # [world, K]
        all_ar = [torch.empty_like(ar_ms) for _ in range(world)]
        dist.all_gather(all_ar, ar_ms)
        all_ar = torch.stack(all_ar)
    else:
        all_steps, all_ar = step_ms[None], ar_ms[None]
    per_step_max = all_steps.max(dim=0).values
    step_stats = {"min_ms": float(per_step_max.min()), "median_ms": float(per_step_max.median()), "max_ms": float(per_step_max.max()),
                  "per_rank_mean_ms": [float(v) for v in all_steps.mean(dim=1)],
                  "slowest_steps": [[int(i), round(float(per_step_max[i]), 2)] for i in torch.argsort(per_step_max, descending=True)[:3]],
                  "note": "device time between consecutive step boundaries (CUDA events); min/median/max over the K timed steps "
                          "of the per-step maximum over ranks"}
    if world > 1 and all_ar.numel():
        step_stats["allreduce_ms"] = {"mean_over_ranks_and_steps": float(all_ar.mean()), "min_rank_mean": float(all_ar.mean(dim=1).min()),
                                      "max_rank_mean": float(all_ar.mean(dim=1).max()),
                                      "note": "events around the gradient all-reduce on each rank: includes the wait for the slowest "
                                              "rank to arrive, so min_rank_mean is the collective's own cost"}
    if rank == 0 and os.environ.get("IA_BENCH_VERBOSE"):
        print("host ms per step:", [round(1e3 * (b - a), 1) for a, b in zip(host_t[:-1], host_t[1:])], file=sys.stderr)
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    if rank == 0 and os.environ.get("IA_BENCH_VERBOSE"):
        print(f"device ms (events) {ms_total:.1f}; host loop ms {1e3 * (host_t[-1] - host_t[0]):.1f}", file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None
    prof.enabled = False
    launches = prof.launches
    summary = prof.summary()
    for a, b in counts:
        n_samples_fg += int(a.item()); n_samples_full += int(b.item())
    total_rays = n_rays * world * K
    value = total_rays / (ms_total / 1e3)
    fg_total = sum_over_ranks(n_samples_fg)
    full_total = sum_over_ranks(n_samples_full)

    # ---- end-to-end measurement (`e2e`): host pinned rays -> H2D -> step -> loss D2H, every step -------
    headroom(gs - 1)            # the model has trained K steps since the first head-room steps: its sample counts have drifted
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    h2d = d2h = 0
    for i in range(W, W + K):
        hb, hbg = host_batches[i]
        db, dbg = hb.to(device, non_blocking=True), hbg.to(device, non_blocking=True)
        h2d = hb.numel() * 4 + hbg.numel() * 4
        batch, bg = unpack_batch(db, dbg)
        loss, out = train_step(cfg, model, arena, var_arena, opt, opt_var, batch, bg, gs, world)
        _ = float(loss.item())                                   # D2H read of the step result
        d2h = 4
        gs += 1
        if rank == 0 and os.environ.get("IA_BENCH_VERBOSE"):
            print("e2e step done at", round(1e3 * time.perf_counter(), 1), file=sys.stderr)
    t1.record()
    barrier()
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = total_rays / (e2e_ms / 1e3)
    gc.enable()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event durations from the timed region) ------------
    prof_ms = (ms_total / K) * PROF_STEPS              # the per-kernel events cover the first PROF_STEPS timed steps
    per_kernel = {k: {"calls_per_step": c / PROF_STEPS, "ms_per_step": t / PROF_STEPS, "share_of_step": t / prof_ms}
                  for k, (c, t, w) in summary.items()}
    dom = max(summary, key=lambda k: summary[k][1])     # the (entry point, shape) with the largest total device time
    calls, ms, work = summary[dom]
    traffic = None
    tpath = next((p_ for p_ in (os.path.join(ROOT, "profiles", "r02_traffic.json"), os.path.join(ROOT, "profiles", "r01_traffic.json"))
                  if os.path.exists(p_)), None)
    if tpath is not None:
        with open(tpath) as f:
            tj = json.load(f)
        if dom in tj:       # DRAM bytes of the ncu capture, rescaled from its row count to this run's average launch
            unit_per_row = tj[dom].get("work_per_row", 2.0 * 2 * (35 * 64 + 64 * 64 + 64 * 1))
            traffic = tj[dom]["dram_bytes_per_row"] * (work / max(calls, 1)) / unit_per_row
    if dom.startswith("ia_mlp"):
        achieved = work / (ms / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        roof = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peaks["source"] + " (sustained bf16 GEMM)", "launches": calls,
                "avg_launch_ms": ms / max(calls, 1), "algorithmic_flop_per_launch": work / max(calls, 1),
                "note": "algorithmic FLOP = 2*MAC of the unpadded fp32 network (x2 for backward).  The kernel issues 3 f16 MMAs "
                        "per product (fp32-equivalent hi/lo split), pads K to 16 and recomputes the forward inside backward: for "
                        "35->64->64->1 that is 5.2x the algorithmic FLOP on the tensor pipe (ncu of the two-tile kernel: tensor pipe "
                        "active 29.7 %, every MMA at the pipe's full rate). It is bound by the dependent phase chain of a tile "
                        "(store -> barrier -> MMA -> wait -> Softplus epilogue, six MUFU per element) of which two are in flight per "
                        "SM, not by the tensor pipe: see DESIGN.md section 4.2 and profiles/r02_ncu_mlp_duo.md"}
    else:
        achieved = work / (ms / 1e3) / 1e9
        peak = peaks["hbm_gbs"]
        roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peaks["source"], "launches": calls, "avg_launch_ms": ms / max(calls, 1),
                "algorithmic_bytes_per_launch": work / max(calls, 1),
                "note": "algorithmic bytes = BASELINE.md section 3 per point-evaluation (every corner counted as an HBM access) against "
                        "the HBM copy peak; rows that share cells (the six taps of a sample, neighbouring samples of a ray) are served "
                        "by L1 / L2, so this fraction can exceed 1 -- the hard bounds for resident tables are the measured sector-gather "
                        "peaks in hashgrid_microbench.sector_gather_peaks (DESIGN.md section 4.1)"}

    try:        # reporting only: never lose the bench line over it
        step_roof = step_roofline(fg_total / total_rays, (full_total - fg_total) / total_rays, value / world, peaks, args.grad_type,
                                  colour_flop=mlp_flops_of(model.texture),
                                  bg_flop=mlp_flops_of(model.geometry_bg) + mlp_flops_of(model.texture_bg),
                                  tables_l2_resident=int(cfg.model.geometry.xyz_encoding_config.log2_hashmap_size) <= 19)
    except Exception as e:  # pragma: no cover
        step_roof = {"error": repr(e)}
    cpu = cpu_baseline_leg(cfg, model, args) if world == 1 and not args.no_cpu_baseline else None
    cpu1 = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu1 = cpu_config1_protocol()
        except Exception as e:  # pragma: no cover
            cpu1 = {"error": repr(e)}
    hg = hashgrid_microbench(device, peaks) if world == 1 else None

    line = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_description(args, cfg, world),
        "workload_measured": {
            "mean_fg_samples_per_ray": fg_total / total_rays, "mean_samples_per_ray_full": full_total / total_rays,
            "samples_per_s": full_total / (ms_total / 1e3),
            "hash_point_evals_per_s": ((13 if args.grad_type == "finite_difference" else 7) * fg_total + (full_total - fg_total)) / (ms_total / 1e3)},
        "implementation": {
            "mlp": ("tcgen05 tensor cores, 3xf16-split operands + fp32 accumulate (fp32-equivalent, parity-tested at 1e-3); the 65-wide "
                    "centre evaluation returns the last hidden layer from the same kernel and applies the output layer with "
                    "the streaming fp32 linear64 kernels") if args.mlp == "tc" else "fp32 FFMA kernels",
            "optimizer": "fused AdamW inside the timed region", "occupancy_refresh": "every 16th step inside the timed region",
            "parallelism": f"dp{world} (ray-sharded, one NCCL all-reduce of the gradient arena per step)",
            "l2_window": L2_WINDOW,
            "hash_tables": "fp16 shadow (IA_TABLE_FP16=1)" if os.environ.get("IA_TABLE_FP16", "0") not in ("0", "") else "fp32"},
        "step_ms": step_stats,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / K},
        "gpu_launches": launches,
        "roofline": roof,
        "step_roofline": step_roof,
        "cpu_baseline": cpu,
        "cpu_baseline_config1": cpu1,
        "kernels": per_kernel,
        "hashgrid_microbench": hg,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# CPU oracle legs (cpu_baseline of the B200 arm, and the whole `--impl reference` arm)
# ---------------------------------------------------------------------------------------------------------
def build_oracle(cfg, fg_binary=None):
    """CPU oracle model of the same workload.  fg_binary: occupancy grid to reuse; None -> one warm-up refresh of the
    foreground grid with the oracle's own SDF and an analytic background grid (cells inside the unit ball)."""
    from instant_angelo_b200.config import to_primitive
    from oracle import model_ref as mr
    torch.manual_seed(42)
    ref = mr.RefNeuSModel(to_primitive(cfg.model))
    ref.train()
    ref.update_step(0, GLOBAL_STEP0, update_occupancy=False)
    if fg_binary is not None:
        ref.occupancy_grid.binary = fg_binary
    else:
        g = torch.Generator().manual_seed(4242)
        ref.occupancy_grid.every_n_step(0, ref.occ_eval_fn, occ_thre=cfg.model.grid_prune_occ_thre, gen=g)
    c = (torch.arange(256, dtype=torch.float32) + 0.5) / 256 - 0.5
    x, y, z = torch.meshgrid(c, c, c, indexing="ij")
    ref.occupancy_grid_bg.binary = (x * x + y * y + z * z) < 0.25
    return ref


def oracle_step(ref, cfg, rays, rgb, pts, nrm, conf, bg, gs):
    from oracle import model_ref as mr
    ref.zero_grad(set_to_none=True)
    ref.background_color = bg
    out = ref.forward_(rays)
    batch = {"rays": rays, "rgb": rgb, "pts": pts, "pts_normal": nrm, "pts_weights": conf}
    from instant_angelo_b200.config import to_primitive
    terms = mr.training_loss(ref, out, batch, to_primitive(cfg.system.loss), gs)
    terms["loss"].backward()
    return float(terms["loss"]), int(out["num_samples_full"])


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline_leg(cfg, model, args):
    """Oracle ("port") timed on this box's host cores on a bounded sample of the same workload."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = build_oracle(cfg, fg_binary=model.occupancy_grid.binary.cpu())
    ref.occupancy_grid_bg.binary = model.occupancy_grid_bg.binary.cpu()
    def run(n_rays):
        buf, bg = make_batches(1, n_rays, 0, pin=False)[0]
        b, bgc = unpack_batch(buf, bg)
        t = time.perf_counter()
        _, n_s = oracle_step(ref, cfg, b["rays"], b["rgb"], b["pts"], b["pts_normal"], b["pts_weights"], bgc, GLOBAL_STEP0)
        return time.perf_counter() - t, n_s

    t0 = time.perf_counter()
    run(32)                                                     # warm-up (thread pool, allocator)
    probe_s, _ = run(256)                                       # calibration at a batch large enough to amortise fixed costs
    n = args.cpu_rays if args.cpu_rays > 0 else max(256, min(8192, int(256 * 15.0 / max(probe_s, 1e-3))))   # ~15 s of CPU work
    t1 = time.perf_counter()
    dt, ns = run(n)
    t2 = t1 + dt
    return {"value": n / (t2 - t1), "unit": "rays/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(),
            "sample": f"{n} rays of the same workload (same occupancy grids, {ns / n:.1f} samples/ray), forward + losses + backward "
                      f"through the CPU oracle (oracle/model_ref.py, PyTorch fp32, {cores} threads); warm-up run {t1 - t0:.1f} s, timed run {t2 - t1:.1f} s"}


def cpu_config1_protocol():
    """BASELINE.md section 2, CPU row: BASELINE.json configs[0] -- the geometry block of configs/neus-colmap.yaml (16-level
    2^19 hash grid -> 35->64->13 sphere-initialised MLP, grad_type analytic as shipped) on an analytic-sphere scene,
    R = 4096 rays x 128 samples/ray, forward + eikonal/SDF loss + backward through the CPU oracle with all host threads;
    2 warm-ups, median of 5 perf_counter runs."""
    from instant_angelo_b200 import configs
    from instant_angelo_b200.config import to_primitive
    from oracle import model_ref as mr
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    geo = mr.RefVolumeSDF(to_primitive(configs.neus_colmap_geometry("analytic")))
    geo.train()
    geo.update_step(0, 20000)
    g = torch.Generator().manual_seed(1)
    R, S = 4096, 128
    o = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1) * 2.0
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(R, 3, generator=g), dim=-1)
    t = torch.linspace(0.5, 3.5, S)[None, :, None]
    pts = (o[:, None, :] + d[:, None, :] * t).reshape(-1, 3).clamp(-2.5, 2.5)
    target = pts.norm(dim=-1) - 0.5                                  # analytic sphere SDF

    def run():
        geo.zero_grad(set_to_none=True)
        t0 = time.perf_counter()
        sdf, grad = geo(pts, with_grad=True, with_feature=False)
        loss = ((grad.norm(dim=-1) - 1.0) ** 2).mean() + (sdf - target).abs().mean()
        loss.backward()
        return time.perf_counter() - t0

    for _ in range(2):
        run()
    ts = sorted(run() for _ in range(5))
    med = ts[2]
    return {"value": R / med, "unit": "rays/s", "samples_per_s": R * S / med, "hash_point_evals_per_s": R * S / med,
            "seconds_median_of_5": med, "cores": cores, "kind": "port", "cpu_model": cpu_model(),
            "sample": f"BASELINE.json configs[0]: neus-colmap.yaml geometry block, {R} rays x {S} samples/ray, analytic normals, "
                      f"forward + backward, CPU oracle, {cores} threads, 2 warm-ups, median of 5"}


def run_reference(args):
    """Reference arm: the reference's path cannot run here (tinycudann / nerfacc are CUDA-only third-party packages
    that are not installable offline), so this times the CPU oracle restatement of the SAME workload on all host
    cores.  Each step is a bounded sample of `--cpu-rays` rays."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = workload_config(args)
    ref = build_oracle(cfg)
    n = args.cpu_rays if args.cpu_rays > 0 else 512
    K, W = args.steps, args.warmup
    batches = make_batches(K + W, n, 0, pin=False)
    gs = GLOBAL_STEP0
    total_samples = 0
    for i in range(W):
        b, bg = unpack_batch(*batches[i])
        oracle_step(ref, cfg, b["rays"], b["rgb"], b["pts"], b["pts_normal"], b["pts_weights"], bg, gs)
    t0 = time.perf_counter()
    for i in range(W, W + K):
        b, bg = unpack_batch(*batches[i])
        _, ns = oracle_step(ref, cfg, b["rays"], b["rgb"], b["pts"], b["pts_normal"], b["pts_weights"], bg, gs)
        total_samples += ns
    dt = time.perf_counter() - t0
    value = n * K / dt
    sample = (f"each step is a bounded sample of {n} rays of the workload ({WORKLOADS[args.config][1]}, {args.grad_type}), "
              f"{total_samples / max(n * K, 1):.1f} samples/ray, CPU oracle forward + losses + backward, {cores} threads")
    line = {"impl": "reference", "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_description(args, cfg, args.gpus),
            "note": "the reference's GPU path (tinycudann + nerfacc) is not installable in this image and the reference has no CPU "
                    "implementation of its own; this arm times the CPU oracle port (oracle/) of the same workload",
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "cpu_model": cpu_model(), "sample": sample},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Print THE JSON line on the real stdout (everything else a library prints -- e.g. NCCL's version banner --
    has been routed to stderr by main())."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # stray prints of libraries go to stderr; stdout carries exactly one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_GPU, help="rays per GPU per step (BASELINE config 5 sweeps 8192..131072 per GPU)")
    ap.add_argument("--config", default="sparse", choices=sorted(WORKLOADS),
                    help="sparse = BASELINE.json configs[1] (the metric's configuration, default); dense21 = configs[2] (2^21-entry "
                         "tables, 256 samples/ray); wreflection = configs[3] (UniSDF V3 heads)")
    ap.add_argument("--grad-type", dest="grad_type", default="finite_difference", choices=["finite_difference", "analytic"],
                    help="model.geometry.grad_type: finite_difference = the BASELINE north-star workload (default); analytic = the "
                         "reference YAML's shipped value (autograd normals with second-order adjoints + 6 FD curvature taps)")
    ap.add_argument("--mlp", default="tc", choices=["fp32", "tc"],
                    help="MLP arithmetic: tc = tcgen05 tensor cores with 3xf16-split operands (fp32-equivalent, default); fp32 = FFMA kernels")
    ap.add_argument("--cpu-rays", type=int, default=0,
                    help="rays per step of the bounded CPU-oracle sample (0: ~15 s of CPU work for cpu_baseline, 512 for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--l2-persist", dest="l2_persist", type=int, default=int(os.environ.get("IA_L2_PERSIST_MB", "0")),
                    help="MB of hash table to keep in a persisting L2 window (ia_l2_persist); 0 = off")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
