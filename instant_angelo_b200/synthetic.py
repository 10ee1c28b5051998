"""Synthetic analytic-sphere scene used by bench.py and the full-size tests: no dataset
or checkpoint is available offline, so rays come from pinhole cameras on a sphere around an analytic sphere of
radius 0.5 (= sphere_init_radius) whose colour is 0.5 + 0.5 * normal.  Ray construction follows the reference's
get_ray_directions / get_rays (models/ray_utils.py:9-43) and per-step pixel sampling
(systems/neus.py:49-55, 95); everything is generated on the CPU with a seeded torch.Generator."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


class SphereScene:
    def __init__(self, n_cameras: int = 32, width: int = 512, height: int = 512, focal: float = 560.0,
                 cam_radius: float = 1.0, sphere_radius: float = 0.5, seed: int = 42):
        self.w, self.h, self.focal, self.sphere_radius = width, height, focal, sphere_radius
        g = torch.Generator().manual_seed(seed)
        # camera centres on a sphere, looking at the origin (OpenGL convention: camera looks along -z)
        centers = F.normalize(torch.randn(n_cameras, 3, generator=g), dim=-1) * cam_radius
        fwd = F.normalize(-centers, dim=-1)
        up = torch.tensor([0.0, 0.0, 1.0]).expand_as(fwd)
        right = F.normalize(torch.cross(fwd, up, dim=-1), dim=-1)
        up2 = torch.cross(right, fwd, dim=-1)
        self.c2w = torch.stack([right, up2, -fwd, centers], dim=-1)      # [N,3,4]
        self.n_cameras = n_cameras

    def sample(self, n_rays: int, gen: torch.Generator):
        """Returns rays[n,6] (origin, unit direction) and target rgb[n,3] on the CPU."""
        idx = torch.randint(0, self.n_cameras, (n_rays,), generator=gen)
        x = torch.randint(0, self.w, (n_rays,), generator=gen)
        y = torch.randint(0, self.h, (n_rays,), generator=gen)
        # get_ray_directions: pixel centres, camera looks along -z, y down
        dirs = torch.stack([(x.float() + 0.5 - self.w / 2) / self.focal, -(y.float() + 0.5 - self.h / 2) / self.focal,
                            -torch.ones(n_rays)], dim=-1)
        c2w = self.c2w[idx]
        rays_d = (dirs[:, None, :] * c2w[:, :, :3]).sum(-1)
        rays_o = c2w[:, :, 3]
        rays_d = F.normalize(rays_d, p=2, dim=-1)
        # analytic target: first hit with the sphere
        b = (rays_o * rays_d).sum(-1)
        c = (rays_o * rays_o).sum(-1) - self.sphere_radius ** 2
        disc = b * b - c
        hit = disc > 0
        t = -b - torch.sqrt(disc.clamp_min(0))
        n = F.normalize(rays_o + t[:, None] * rays_d, dim=-1)
        rgb = torch.where(hit[:, None], 0.5 + 0.5 * n, torch.ones(n_rays, 3))
        return torch.cat([rays_o, rays_d], dim=-1).contiguous(), rgb.contiguous()

    def surface_points(self, n: int, gen: torch.Generator):
        """Sparse 'SfM' points on the surface with normals and confidences (batch['pts*'], systems/neus.py:63-71)."""
        nrm = F.normalize(torch.randn(n, 3, generator=gen), dim=-1)
        return (nrm * self.sphere_radius).contiguous(), nrm.contiguous(), torch.rand(n, generator=gen)
