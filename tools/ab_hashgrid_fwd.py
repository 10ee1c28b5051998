"""A/B of the two ia_hashgrid_fwd kernels (one-loop vs split dense/hashed loops) on the point sets of a training step:
bitwise comparison of the outputs and CUDA-event timings.  usage: python tools/ab_hashgrid_fwd.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from instant_angelo_b200 import ops, _lib as L

dev = "cuda"
lib, s = L.load(), L.stream()
g = torch.Generator(device=dev).manual_seed(3)
PLS = 1.3195079107728942


def tap_points(S):
    o = torch.nn.functional.normalize(torch.randn(8192, 3, device=dev, generator=g), dim=-1)
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(8192, 3, device=dev, generator=g), dim=-1)
    t = torch.linspace(0.2, 1.2, S // 8192 + 1, device=dev)[None, :, None]
    pts = (o[:, None] + d[:, None] * t).reshape(-1, 3)[:S]
    eps = 2 * 1.5 / 2048
    off = torch.tensor([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], device=dev, dtype=torch.float32) * eps
    return ((pts[:, None, :] + off).clamp(-1.5, 1.5) / 3.0 + 0.5).reshape(-1, 3).contiguous()


def timed(fn, reps=5):
    ts = []
    for _ in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts[2:])[len(ts[2:]) // 2]


def run(variant, x, table, plan, active, out):
    assert lib.ia_debug_hashgrid_fwd_generic(variant) == 0
    fn = lambda: L.check(lib.ia_hashgrid_fwd(x.data_ptr(), x.shape[0], table.data_ptr(), C.byref(plan), active, out.data_ptr(), s))
    ms = timed(fn)
    lib.ia_debug_hashgrid_fwd_generic(-1)
    return ms


def main():
    x_taps = tap_points(1550000)                        # 9.3 M tap rows, as one FD set of the bench step
    x_rand = torch.rand(1 << 22, 3, device=dev, generator=g)
    x_edge = torch.rand(1000003, 3, device=dev, generator=g)      # ragged tail + exact faces / corners
    x_edge[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [1, 0, 1], [0.5, 0.25, 0.75], [0, 1, 0], [1, 1, 0], [0.999999, 0.5, 0.5], [0.5, 0.5, 1]], device=dev)
    cases = [("taps 9.3M  T=2^19 La=16", x_taps, 19, 16), ("taps 9.3M  T=2^19 La=11", x_taps, 19, 11), ("taps 9.3M  T=2^19 La=4", x_taps, 19, 4),
             ("rand 4.2M  T=2^19 La=16", x_rand, 19, 16), ("rand 4.2M  T=2^21 La=16", x_rand, 21, 16), ("taps 9.3M  T=2^21 La=16", x_taps, 21, 16),
             ("edge 1.0M  T=2^19 La=16", x_edge, 19, 16), ("edge 77    T=2^19 La=3", x_edge[:77].contiguous(), 19, 3)]
    ok = True
    for label, x, log2_t, active in cases:
        plan = ops.make_grid_plan(16, 2, log2_t, 32, PLS)
        table = torch.randn(plan.n_params, device=dev, generator=g) * 0.1
        a = torch.full((x.shape[0], 32), float("nan"), device=dev)
        b = torch.full((x.shape[0], 32), float("nan"), device=dev)
        ms_generic = run(1, x, table, plan, active, a)
        ms_split = run(0, x, table, plan, active, b)
        same = torch.equal(a, b)
        diff = float((a - b).abs().max())
        ok &= diff <= 1e-6 and bool(torch.isfinite(b).all())
        print(f"{label}: generic {ms_generic:.3f} ms  split {ms_split:.3f} ms  ({ms_generic / ms_split:.2f}x)  bit-identical {same}  max|diff| {diff:.2e}", flush=True)
        del table, a, b
    print("AB", "OK" if ok else "MISMATCH")


if __name__ == "__main__":
    main()
