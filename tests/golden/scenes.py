"""Small synthetic scene + config shared by the golden generator (tests/golden/make_golden.py), the oracle
tests and the GPU parity tests.  Shapes follow configs/neuralangelo-colmap_sparse(-wreflection).yaml of the
reference with tables shrunk (8 levels, 2^12 entries) so the CPU oracle runs in seconds."""
from __future__ import annotations

import math

import torch


def golden_encoding_config():
    return {"otype": "ProgressiveBandHashGrid", "n_levels": 8, "n_features_per_level": 2, "log2_hashmap_size": 12,
            "base_resolution": 4, "per_level_scale": 1.5, "include_xyz": True, "start_level": 4, "start_step": 0,
            "update_steps": 10}


def golden_model_config(texture: str = "volume-dual-color", learned_background: bool = True, feature_dim: int = 65,
                        grad_type: str = "finite_difference"):
    enc = golden_encoding_config()
    mlp_geo = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64,
               "n_hidden_layers": 2, "sphere_init": True, "sphere_init_radius": 0.5, "weight_norm": True}
    mlp2 = {"otype": "VanillaMLP", "activation": "ReLU", "output_activation": "none", "n_neurons": 64, "n_hidden_layers": 2}
    mlp1 = dict(mlp2, n_hidden_layers=1)
    if texture == "volume-dual-colorV3":
        tex = {"name": texture, "input_feature_dim": feature_dim + 6, "dir_encoding_config": {"otype": "SphericalHarmonics", "degree": 3},
               "mlp_network_config": mlp2, "weitht_network_config": dict(mlp1, output_activation="sigmoid"),
               "color_activation": "sigmoid"}
    else:
        tex = {"name": texture, "input_feature_dim": feature_dim + 6, "dir_encoding_config": {"otype": "SphericalHarmonics", "degree": 4},
               "mlp_network_config": mlp2, "color_activation": "sigmoid"}
    cfg = {
        "name": "neus", "radius": 1.5, "num_samples_per_ray": 32, "train_num_rays": 48, "max_train_num_rays": 8192,
        "grid_prune": True, "grid_prune_occ_thre": 0.001, "dynamic_ray_sampling": False, "batch_image_sampling": True,
        "randomized": True, "ray_chunk": 2048, "cos_anneal_end": 100, "learned_background": learned_background,
        "background_color": "random", "variance": {"init_val": 0.3, "modulate": False},
        "geometry": {"name": "volume-sdf", "radius": 1.5, "feature_dim": feature_dim, "grad_type": grad_type,
                     "finite_difference_eps": "progressive", "isosurface": None, "xyz_encoding_config": enc,
                     "mlp_network_config": mlp_geo},
        "texture": tex,
        "num_samples_per_ray_bg": 16,
        "geometry_bg": {"name": "volume-density", "radius": 1.5, "feature_dim": 8, "density_activation": "trunc_exp",
                        "density_bias": -1, "isosurface": None, "xyz_encoding_config": dict(enc), "mlp_network_config": mlp1},
        "texture_bg": {"name": "volume-radiance", "input_feature_dim": 8,
                       "dir_encoding_config": {"otype": "SphericalHarmonics", "degree": 4}, "mlp_network_config": mlp2,
                       "color_activation": "sigmoid"},
    }
    return cfg


def golden_loss_config():
    return {"lambda_sdf_l1": [0, 1, 0, 200], "lambda_normal": 0.05, "lambda_rgb_mse": 10.0, "lambda_rgb_l1": 0.0,
            "lambda_mask": 0.0, "lambda_eikonal": 0.1, "lambda_curvature": [0, 0, 0.5, 50], "lambda_sparsity": 0.01,
            "lambda_distortion": 0.0, "lambda_distortion_bg": 0.0, "lambda_opaque": 0.0, "sparsity_scale": 1.0}


def sphere_shell_binary(res: int, radius: float, r_in: float = 0.3, r_out: float = 0.75) -> torch.Tensor:
    """Occupancy grid of an analytic shell around the sphere-init surface (cell centres with r_in<|x|<r_out)."""
    c = (torch.arange(res, dtype=torch.float32) + 0.5) / res * (2 * radius) - radius
    x, y, z = torch.meshgrid(c, c, c, indexing="ij")
    d = torch.sqrt(x * x + y * y + z * z)
    return (d > r_in) & (d < r_out)


def make_rays(n: int, gen: torch.Generator, cam_radius: float = 1.0):
    """Pinhole-like rays from camera centres on a sphere of radius `cam_radius` (inside the +-1.5 AABB, the
    reference's normal regime) looking roughly at the origin; a quarter of the rays start outside the box
    (radius 4) and a few miss it entirely.  Returns (rays[n,6], target rgb[n,3])."""
    o = torch.nn.functional.normalize(torch.randn(n, 3, generator=gen), dim=-1) * cam_radius
    o[: n // 4] *= 4.0 / cam_radius
    target = (torch.rand(n, 3, generator=gen) - 0.5) * 0.8
    target[-3:] += 6.0                       # misses
    d = torch.nn.functional.normalize(target - o, dim=-1)
    rays = torch.cat([o, d], dim=-1).contiguous()
    rgb = torch.rand(n, 3, generator=gen)
    return rays, rgb
