"""CPU: config loading (YAML + ${...} resolvers + CLI overrides, reference utils/misc.py:7-31), the programmatic
config builders against the reference's YAML files (when /root/reference is present), schedules and level logic."""
import os

import pytest
import torch

from instant_angelo_b200 import configs
from instant_angelo_b200.config import load_config, to_primitive
from instant_angelo_b200.losses import C

REF = "/root/reference/configs"


def test_yaml_loader_resolvers(tmp_path):
    f = tmp_path / "c.yaml"
    f.write_text("""
name: test-${basename:${dataset.root_dir}}
dataset: {root_dir: /data/scene7}
model:
  radius: 1.5
  geometry: {radius: "${model.radius}", feature_dim: 65}
  texture: {input_feature_dim: "${add:${model.geometry.feature_dim}, 6}"}
trainer: {max_steps: 20000}
system:
  warmup_steps: 500
  scheduler: {gamma: "${calc_exp_lr_decay_rate:0.1,${sub:${trainer.max_steps},${system.warmup_steps}}}", eps: 1.e-15}
""")
    cfg = load_config(str(f), cli_args=["model.geometry.grad_type=finite_difference", "model.radius=2.5"])
    assert cfg.name == "test-scene7" and cfg.model.geometry.radius == 2.5 and cfg.model.texture.input_feature_dim == 71
    assert cfg.model.geometry.grad_type == "finite_difference"
    assert abs(cfg.system.scheduler.gamma - 0.1 ** (1 / 19500)) < 1e-12 and cfg.system.scheduler.eps == 1e-15


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("yaml_name,builder", [
    ("neuralangelo-colmap_sparse.yaml", configs.neuralangelo_colmap_sparse),
    ("neuralangelo-colmap_dense.yaml", configs.neuralangelo_colmap_dense),
    ("neuralangelo-colmap_sparse-wreflection.yaml", configs.neuralangelo_colmap_sparse_wreflection)])
def test_builders_match_reference_yaml(yaml_name, builder):
    ref = to_primitive(load_config(os.path.join(REF, yaml_name), cli_args=["dataset.root_dir=/x/y"]))
    got = to_primitive(builder())

    def check(a, b, path):
        if isinstance(a, dict):
            for k, v in a.items():
                assert k in b, f"{path}.{k} missing from the builder"
                check(v, b[k], f"{path}.{k}")
        elif isinstance(a, float) or isinstance(b, float):
            assert abs(float(a) - float(b)) < 1e-12, f"{path}: {a} vs {b}"
        else:
            assert a == b, f"{path}: {a} vs {b}"

    check(ref["model"], got["model"], "model")
    check(ref["system"]["loss"], got["system"]["loss"], "system.loss")
    check(ref["system"]["optimizer"], got["system"]["optimizer"], "system.optimizer")
    check(ref["system"]["scheduler"], got["system"]["scheduler"], "system.scheduler")
    assert ref["trainer"]["max_steps"] == got["trainer"]["max_steps"] and ref["system"]["warmup_steps"] == got["system"]["warmup_steps"]


def test_schedules_and_progressive_levels():
    assert C([0, 0, 0.5, 5000], 1000) == 0.1 and C(10.0, 5) == 10.0 and C([0, 1, 0, 20000], 5000) == 0.75
    from instant_angelo_b200.network_utils import ProgressiveBandHashGrid
    from instant_angelo_b200.geometry import VolumeSDF
    cfg = configs.neuralangelo_colmap_sparse("finite_difference").model.geometry
    cfg.xyz_encoding_config["log2_hashmap_size"] = 10      # keep the CPU test tiny
    geo = VolumeSDF(cfg)
    enc = geo.encoding.encoding
    assert isinstance(enc, ProgressiveBandHashGrid) and enc.active_levels == 4
    for step, level in [(0, 4), (5000, 4), (5999, 4), (6000, 5), (12000, 11), (17000, 16), (20000, 16)]:
        geo.update_step(0, step)
        assert enc.current_level == level
        want_eps = 2 * 1.5 / (32 * 1.3195079107728942 ** (level - 1))
        assert abs(geo._finite_difference_eps - want_eps) < 1e-12
    geo.update_step(0, 0)          # the mask only ever grows (Appendix C-7)
    assert enc.current_level == 4 and enc.active_levels == 16 and enc.mask.sum() == 32
    assert abs(2 * 1.5 / (32 * 1.3195079107728942 ** 3) - 0.0408) < 1e-4
