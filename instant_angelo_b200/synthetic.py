"""Synthetic analytic-sphere scene used by bench.py and the full-size tests: no dataset
or checkpoint is available offline, so rays come from pinhole cameras on a sphere around an analytic sphere of
radius 0.5 (= sphere_init_radius) whose colour is 0.5 + 0.5 * normal.  Ray construction follows the reference's
get_ray_directions / get_rays (models/ray_utils.py:9-43) and per-step pixel sampling
(systems/neus.py:49-55, 95); everything is generated on the CPU with a seeded torch.Generator."""
from __future__ import annotations

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


class SphereDataset:
    """The synthetic scene with the attribute surface of the reference's ColmapDatasetBase (datasets/colmap.py:297-316)
    that NeuSSystem.preprocess_data reads: all_c2w [N,3,4], all_images [N,H,W,3], all_fg_masks [N,H,W], directions [H,W,3],
    all_points / all_points_confidence / pts3d_normal (sparse 'SfM' points on the surface), all_fg_indexs / all_bg_indexs
    (image, y, x) triples, w, h, img_wh, has_mask, apply_mask.  Tensors live on `device`, as the reference keeps its
    dataset on the training GPU (datasets/colmap.py: `.to(self.rank)`)."""

    def __init__(self, n_cameras: int = 32, width: int = 512, height: int = 512, focal: float = 560.0, cam_radius: float = 1.0,
                 sphere_radius: float = 0.5, n_points: int = 65536, seed: int = 42, device="cpu", apply_mask: bool = False):
        from .systems import get_ray_directions, get_rays
        scene = SphereScene(n_cameras, width, height, focal, cam_radius, sphere_radius, seed)
        self.w, self.h, self.img_wh = width, height, (width, height)
        self.has_mask, self.apply_mask = True, apply_mask
        self.directions = get_ray_directions(width, height, focal, focal, width / 2, height / 2)
        self.all_c2w = scene.c2w.float()
        images, masks = [], []
        for i in range(n_cameras):
            rays_o, rays_d = get_rays(self.directions, self.all_c2w[i])
            rays_d = F.normalize(rays_d, p=2, dim=-1)
            b = (rays_o * rays_d).sum(-1)
            disc = b * b - ((rays_o * rays_o).sum(-1) - sphere_radius ** 2)
            hit = disc > 0
            t = -b - torch.sqrt(disc.clamp_min(0))
            n = F.normalize(rays_o + t[:, None] * rays_d, dim=-1)
            images.append(torch.where(hit[:, None], 0.5 + 0.5 * n, torch.ones_like(n)).view(height, width, 3))
            masks.append(hit.float().view(height, width))
        self.all_images, self.all_fg_masks = torch.stack(images), torch.stack(masks)
        g = torch.Generator().manual_seed(seed + 1)
        self.all_points, self.pts3d_normal, self.all_points_confidence = scene.surface_points(n_points, g)
        fg = self.all_fg_masks > 0.5
        self.all_fg_indexs, self.all_bg_indexs = torch.nonzero(fg), torch.nonzero(~fg)
        for k in ("directions", "all_c2w", "all_images", "all_fg_masks", "all_points", "pts3d_normal", "all_points_confidence",
                  "all_fg_indexs", "all_bg_indexs"):
            setattr(self, k, getattr(self, k).to(device))

    def __len__(self):
        return len(self.all_images)
