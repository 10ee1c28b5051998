"""Programmatic builders for the reference's shipped configurations (model + loss sections), so the GPU box
(which has no /root/reference) can instantiate them.  Values restate
configs/neuralangelo-colmap_sparse.yaml, configs/neuralangelo-colmap_dense.yaml,
configs/neuralangelo-colmap_sparse-wreflection.yaml and configs/neus-colmap.yaml;
tests/test_config.py diffs them against the YAML files when the reference tree is present.
`load_config` (config.py) reads the YAMLs directly when a user has them.
"""
from __future__ import annotations

from .config import Config, to_config

PLS = 1.3195079107728942


def _hashgrid(log2_hashmap_size=19, start_step=5000):
    return {"otype": "ProgressiveBandHashGrid", "n_levels": 16, "n_features_per_level": 2,
            "log2_hashmap_size": log2_hashmap_size, "base_resolution": 32, "per_level_scale": PLS,
            "include_xyz": True, "start_level": 4, "start_step": start_step, "update_steps": 1000}


def _mlp(n_hidden_layers=2, **extra):
    d = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64,
         "n_hidden_layers": n_hidden_layers}
    d.update(extra)
    return d


def _sh(degree=4):
    return {"otype": "SphericalHarmonics", "degree": degree}


def neuralangelo_colmap_sparse(grad_type: str = "analytic", log2_hashmap_size: int = 19, mlp_otype: str = "VanillaMLP") -> Config:
    """configs/neuralangelo-colmap_sparse.yaml (model + system.loss)."""
    radius, feature_dim = 1.5, 65
    model = {
        "name": "neus", "radius": radius, "num_samples_per_ray": 512, "train_num_rays": 256, "max_train_num_rays": 8192,
        "grid_prune": True, "grid_prune_occ_thre": 0.001, "dynamic_ray_sampling": True, "batch_image_sampling": True,
        "randomized": True, "ray_chunk": 2048, "cos_anneal_end": 20000, "learned_background": True,
        "background_color": "random", "variance": {"init_val": 0.3, "modulate": False},
        "geometry": {"name": "volume-sdf", "radius": radius, "feature_dim": feature_dim, "grad_type": grad_type,
                     "finite_difference_eps": "progressive",
                     "isosurface": {"method": "mc", "resolution": 512, "chunk": 2097152, "threshold": 0.001},
                     "xyz_encoding_config": _hashgrid(log2_hashmap_size),
                     "mlp_network_config": _mlp(2, otype=mlp_otype, sphere_init=True, sphere_init_radius=0.5, weight_norm=True)},
        "texture": {"name": "volume-dual-color", "input_feature_dim": feature_dim + 6, "diffuse_warmup_steps": 5000,
                    "dir_encoding_config": _sh(4), "mlp_network_config": _mlp(2, otype=mlp_otype), "color_activation": "sigmoid"},
        "num_samples_per_ray_bg": 256,
        "geometry_bg": {"name": "volume-density", "radius": radius, "feature_dim": 8, "density_activation": "trunc_exp",
                        "density_bias": -1, "isosurface": None, "xyz_encoding_config": _hashgrid(log2_hashmap_size),
                        "mlp_network_config": _mlp(1, otype=mlp_otype)},
        "texture_bg": {"name": "volume-radiance", "input_feature_dim": 8, "dir_encoding_config": _sh(4),
                       "mlp_network_config": _mlp(2, otype=mlp_otype), "color_activation": "sigmoid"},
    }
    loss = {"lambda_sdf_l1": [0, 1, 0, 20000], "lambda_normal": 0.0, "lambda_rgb_mse": 10.0, "lambda_rgb_l1": 0.0,
            "lambda_mask": 0.0, "lambda_eikonal": 0.1, "lambda_curvature": [0, 0, 0.5, 5000], "lambda_sparsity": 0.0,
            "lambda_distortion": 0.0, "lambda_distortion_bg": 0.0, "lambda_opaque": 0.0, "sparsity_scale": 1.0}
    optimizer = {"name": "AdamW", "args": {"lr": 0.01, "betas": [0.9, 0.99], "eps": 1e-15},
                 "params": {"geometry": {"lr": 0.01}, "texture": {"lr": 0.01}, "geometry_bg": {"lr": 0.01},
                            "texture_bg": {"lr": 0.01}, "variance": {"lr": 0.001}}}
    warmup_steps, max_steps = 500, 20000
    scheduler = {"name": "SequentialLR", "interval": "step", "milestones": [warmup_steps],
                 "schedulers": [{"name": "LinearLR", "args": {"start_factor": 0.01, "end_factor": 1.0, "total_iters": warmup_steps}},
                                {"name": "ExponentialLR", "args": {"gamma": 0.1 ** (1.0 / (max_steps - warmup_steps))}}]}
    return to_config({"seed": 42, "model": model,
                      "dataset": {"name": "colmap", "apply_mask": False},
                      "system": {"name": "neus-system", "loss": loss, "optimizer": optimizer, "warmup_steps": warmup_steps,
                                 "scheduler": scheduler},
                      "export": {"chunk_size": 2097152, "export_vertex_color": True},
                      "trainer": {"max_steps": max_steps}})


def neuralangelo_colmap_dense(grad_type: str = "analytic", log2_hashmap_size: int = 19) -> Config:
    """configs/neuralangelo-colmap_dense.yaml: as sparse, dual-colour background head and weaker curvature /
    longer-lived point losses.  BASELINE config 3 runs it with log2_hashmap_size=21 and 256 samples/ray."""
    cfg = neuralangelo_colmap_sparse(grad_type, log2_hashmap_size)
    cfg.model.texture_bg["name"] = "volume-dual-color"
    cfg.model.geometry.isosurface["threshold"] = 0.0
    cfg.model.texture.pop("diffuse_warmup_steps", None)
    cfg.system.loss["lambda_curvature"] = [0, 0, 0.05, 5000]
    cfg.system.loss["lambda_sdf_l1"] = [0, 1, 0.1, 20000]
    return cfg


def neuralangelo_colmap_sparse_wreflection(grad_type: str = "analytic") -> Config:
    """configs/neuralangelo-colmap_sparse-wreflection.yaml: UniSDF colour heads (VolumeDualColorV3), SH degree 3,
    1024 samples/ray."""
    cfg = neuralangelo_colmap_sparse(grad_type)
    m = cfg.model
    m["num_samples_per_ray"] = 1024
    m["texture"] = to_config({"name": "volume-dual-colorV3", "input_feature_dim": 65 + 6, "diffuse_warmup_steps": 5000,
                              "dir_encoding_config": _sh(3), "mlp_network_config": _mlp(2),
                              "weitht_network_config": _mlp(1, output_activation="sigmoid"), "color_activation": "sigmoid"})
    cfg.system.loss["lambda_rgb_mse"] = 5.0
    cfg.system.loss["lambda_curvature"] = [0, 0, 1.0e-3, 5000]
    cfg.system.loss["lambda_sdf_l1"] = 0.0
    return cfg


def neus_colmap_geometry(grad_type: str = "analytic") -> Config:
    """Geometry block of configs/neus-colmap.yaml (BASELINE config 1; its texture block is shape-inconsistent
    as shipped)."""
    enc = _hashgrid(19, start_step=0)
    return to_config({"name": "volume-sdf", "radius": 2.5, "feature_dim": 13, "grad_type": grad_type,
                      "isosurface": None, "xyz_encoding_config": enc,
                      "mlp_network_config": _mlp(1, sphere_init=True, sphere_init_radius=0.5, weight_norm=True)})
