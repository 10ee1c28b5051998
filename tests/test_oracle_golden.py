"""CPU: the oracle restatement (oracle/model_ref.py) against fixtures produced by the reference's own Python
(tests/golden/make_golden.py).  Pins VanillaMLP/weight-norm, progressive masking, VolumeSDF FD gradient +
curvature quirks, VolumeDensity, colour heads, get_alpha, forward_/forward_bg_ and the loss terms."""
import numpy as np
import pytest
import torch

from oracle import model_ref as mr
from tests.golden.scenes import golden_loss_config, golden_model_config, sphere_shell_binary
from tests.helpers import GOLDEN_CASES, assert_close, golden_batch, golden_state_dict, load_golden


def build_oracle(fx, case):
    cfg = golden_model_config(**GOLDEN_CASES[case])
    model = mr.RefNeuSModel(cfg)
    missing, unexpected = model.load_state_dict(golden_state_dict(fx), strict=False)
    assert not [k for k in missing if "occupancy" not in k], missing
    assert not [k for k in unexpected if "occupancy" not in k and k != "scene_aabb"], unexpected
    model.train()
    model.occupancy_grid.binary = sphere_shell_binary(128, cfg["radius"])
    if cfg["learned_background"]:
        model.occupancy_grid_bg.binary = torch.ones(256, 256, 256, dtype=torch.bool)
    model.update_step(0, int(fx["global_step"]), update_occupancy=False)
    model.background_color = torch.from_numpy(fx["background_color"])
    return model


@pytest.mark.parametrize("case", list(GOLDEN_CASES))
def test_oracle_matches_reference_python(golden_dir, case):
    fx = load_golden(golden_dir, case)
    model = build_oracle(fx, case)
    batch = golden_batch(fx)
    out = model.forward_(batch["rays"], stratified_u=torch.from_numpy(fx["u_fg"]),
                         rand_directions=torch.from_numpy(fx["rand_directions"]), stratified_u_bg=torch.from_numpy(fx["u_bg"]))
    # integer / index outputs: exact
    assert np.array_equal(out["ray_indices"].numpy(), fx["out.ray_indices"])
    assert int(out["num_samples_full"]) == int(fx["out.num_samples_full"])
    assert np.array_equal(out["rays_valid_full"].numpy(), fx["out.rays_valid_full"])
    if "out.ray_indices_bg" in fx:
        assert np.array_equal(out["ray_indices_bg"].numpy(), fx["out.ray_indices_bg"])
    for k in ["comp_rgb", "comp_normal", "opacity", "depth", "sdf_samples", "sdf_grad_samples", "sdf_laplace_samples",
              "weights", "points", "intervals", "comp_rgb_full", "comp_rgb_bg", "opacity_bg", "depth_bg", "weights_bg"]:
        if "out." + k in fx:
            assert_close(out[k], fx["out." + k], rtol=1e-4, atol=1e-5, name=k)
    terms = mr.training_loss(model, out, batch, golden_loss_config(), int(fx["global_step"]))
    assert_close(terms["loss"], fx["loss"], rtol=1e-4, atol=1e-6, name="loss")
    for k, ref_k in [("rgb_mse", "train/loss_rgb_mse"), ("eikonal", "train/loss_eikonal"), ("curvature", "train/loss_curvature"),
                     ("sdf_l1", "train/loss_sdf_l1"), ("normal_cos", "train/loss_normal_cos"), ("sparsity", "train/loss_sparsity")]:
        if "log." + ref_k in fx:
            assert_close(terms[k], fx["log." + ref_k], rtol=1e-4, atol=1e-6, name=k)
    terms["loss"].backward()
    checked = 0
    for name, p in model.named_parameters():
        key = "grad." + name
        if key not in fx:
            continue
        ref = fx[key]
        scale = np.abs(ref).max()
        assert_close(p.grad, ref, rtol=1e-3, atol=max(1e-7, 1e-4 * scale), name=key)
        checked += 1
    assert checked >= 10
