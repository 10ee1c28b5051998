"""CPU: host-side logic around the round-2 kernels that needs no GPU -- the fp16-shadow option of the hash-grid module, the
loss assembly's tensor-expression form (what the fused kernels are held against on the GPU), the refusal of CPU tensors by
every new operator (there is no CPU fallback), and the environment switches."""
import pytest
import torch

from instant_angelo_b200 import ops
from instant_angelo_b200.losses import C, training_loss
from instant_angelo_b200.network_utils import get_encoding

GRID = {"otype": "ProgressiveBandHashGrid", "n_levels": 4, "n_features_per_level": 2, "log2_hashmap_size": 8, "base_resolution": 4,
        "per_level_scale": 1.5, "start_level": 2, "start_step": 0, "update_steps": 10, "include_xyz": True}


def test_table_precision_option_and_state_dict_keys(monkeypatch):
    monkeypatch.delenv("IA_TABLE_FP16", raising=False)
    plain = get_encoding(3, GRID)
    half = get_encoding(3, {**GRID, "table_precision": "fp16"})
    g_plain, g_half = plain.encoding.encoding, half.encoding.encoding
    assert g_plain.table_precision == "fp32" and g_half.table_precision == "fp16"
    assert g_plain.shadow() is None                                    # fp32 tables: the gathers read .params itself
    assert list(plain.state_dict().keys()) == list(half.state_dict().keys()) == ["encoding.encoding.params"]
    with pytest.raises(ValueError):
        get_encoding(3, {**GRID, "table_precision": "bf16"})
    monkeypatch.setenv("IA_TABLE_FP16", "1")
    assert get_encoding(3, GRID).encoding.encoding.table_precision == "fp16"
    assert get_encoding(3, {**GRID, "table_precision": "fp32"}).encoding.encoding.table_precision == "fp32"   # the config wins
    # update_step marks the shadow stale (the fused optimizer writes through raw pointers) without touching the levels' logic
    half.update_step(0, 25)
    assert g_half._shadow_stale and half.encoding.active_levels == 4


def test_new_operators_refuse_cpu_tensors():
    """No CPU fallback: every operator added in round 2 raises on CPU inputs before touching the library."""
    x3, x1 = torch.rand(8, 3), torch.rand(8, 1)
    ri = torch.zeros(8, dtype=torch.int32)
    cases = [
        lambda: ops.ray_samples(x3, x3, ri, x1, x1),
        lambda: ops.normalize3(x3),
        lambda: ops.contract(x3, 1.5, 0),
        lambda: ops.ray_mix(x3, x1, x3, x1, torch.rand(3)),
        lambda: ops.prune_samples(x1, x1, x1, torch.zeros(2, 2, dtype=torch.int32), 1e-4, 0.0),
        lambda: ops.point_losses(x1, x3, x3, x1, 1.0, 1.0),
        lambda: ops.neus_losses(x3, x3, torch.ones(8, dtype=torch.bool), x1, None, x3, x1, None,
                                {"rgb_mse": 1, "rgb_l1": 1, "eikonal": 1, "opaque": 1, "sparsity": 1}, 1.0),
        lambda: ops.fold_head(torch.rand(64 * 88 + 64 + 10), torch.rand(65, 64), torch.rand(65), 88, 65, 88),
        lambda: ops.table_to_half(torch.rand(16)),
        lambda: ops.l2_persist(torch.rand(16)),
    ]
    for fn in cases:
        with pytest.raises(NotImplementedError, match="Only support cuda inputs"):
            fn()


class _Geo(torch.nn.Module):
    def forward(self, pts, with_grad=True, with_feature=False):
        sdf = pts.norm(dim=-1) - 0.5
        return sdf, torch.nn.functional.normalize(pts, dim=-1) * 0.9


class _Model:
    learned_background = False
    geometry = _Geo()


def test_training_loss_tensor_expression_form():
    """The CPU / tensor-expression form of reference systems/neus.py:130-194 -- the definition the fused GPU kernels are tested
    against -- term by term on hand-checkable inputs, including the scalar-times-weights quirk of the sparse-point term."""
    R, S = 4, 6
    out = {"rays_valid_full": torch.tensor([[True], [True], [False], [True]]),
           "comp_rgb_full": torch.full((R, 3), 0.5), "opacity": torch.tensor([[0.0], [0.5], [1.0], [0.25]]),
           "sdf_grad_samples": torch.tensor([[2.0, 0, 0]] * S), "sdf_samples": torch.zeros(S), "sdf_laplace_samples": torch.full((S, 1), -0.2)}
    batch = {"rgb": torch.tensor([[0.5, 0.5, 0.5], [1.0, 0.5, 0.0], [9.0, 9.0, 9.0], [0.5, 0.5, 0.5]]),
             "pts": torch.tensor([[1.0, 0, 0], [0, 0.25, 0]]), "pts_normal": torch.tensor([[2.0, 0, 0], [0, -1.0, 0]]),
             "pts_weights": torch.tensor([1.0, 0.5])}
    cfg = {"lambda_rgb_mse": 10.0, "lambda_rgb_l1": 1.0, "lambda_eikonal": 0.1, "lambda_mask": 0.0, "lambda_opaque": 0.0,
           "lambda_sparsity": 0.5, "sparsity_scale": 1.0, "lambda_curvature": [0, 0.0, 1.0, 100], "lambda_sdf_l1": 2.0,
           "lambda_distortion": 0.0, "lambda_distortion_bg": 0.0}
    t = training_loss(_Model(), out, batch, cfg, global_step=50)
    assert float(t["rgb_mse"]) == pytest.approx((0.25 + 0.25) / 9)               # one of three valid rays is off by (0.5, 0, -0.5)
    assert float(t["rgb_l1"]) == pytest.approx(1.0 / 9)
    assert float(t["eikonal"]) == pytest.approx(1.0) and float(t["sparsity"]) == pytest.approx(1.0)
    assert float(t["curvature"]) == pytest.approx(0.2) and C(cfg["lambda_curvature"], 50) == pytest.approx(0.5)
    assert float(t["sdf_l1"]) == pytest.approx(((0.5 + 0.25) / 2) * 0.75)       # mean|sdf| * mean(weights)   (Appendix C-11)
    assert float(t["normal_cos"]) == pytest.approx((0.0 + 2.0) / 2)
    want = 10 * (0.5 / 9) + 1 / 9 + 0.1 + float(t["opaque"]) * 0 + 0.5 + 0.5 * 0.2 + 2.0 * float(t["sdf_l1"]) + 2.0 * 1.0
    assert float(t["loss"]) == pytest.approx(want, rel=1e-6)
    assert set(t) == {"rgb_mse", "rgb_l1", "eikonal", "opaque", "sparsity", "curvature", "sdf_l1", "normal_cos", "loss"}


def test_flat_sink_protocol_on_cpu_autograd():
    """The accumulate-behind-a-gate protocol of ops._WeightNormFlatFn / _flat_sink, as plain autograd on the CPU: consumers add
    their gradient into a shared buffer and return None, the producer's backward still runs -- once, after every consumer --
    and sees the sum."""
    acc = torch.zeros(3)
    calls = []

    class Gate(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w):
            ctx.set_materialize_grads(False)
            return w.view_as(w)

        @staticmethod
        def backward(ctx, g):
            calls.append(g)
            return acc.clone() if g is None else g + acc

    class Consumer(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w, x):
            ctx.save_for_backward(x)
            return (w * x).sum()

        @staticmethod
        def backward(ctx, g):
            acc.add_(g * ctx.saved_tensors[0])
            return None, None

    w = torch.ones(3, requires_grad=True)
    gate = Gate.apply(w * 2)
    (Consumer.apply(gate, torch.tensor([1.0, 2, 3])) + Consumer.apply(gate, torch.tensor([10.0, 20, 30])) + gate[:1].sum() * 5).backward()
    assert len(calls) == 1 and calls[0] is not None                              # the slice's gradient arrived through autograd ...
    assert torch.equal(w.grad, torch.tensor([2 * (11.0 + 5), 44.0, 66.0]))       # ... and was added to the accumulated part
