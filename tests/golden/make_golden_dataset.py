"""Generates tests/golden/colmap_scene/ (a tiny synthetic COLMAP model: sparse/0/*.bin + images) and
tests/golden/colmap_dataset.npz with the REFERENCE's own code:

  * datasets/colmap_utils.py read_cameras_binary / read_images_binary / read_points3d_binary parse the .bin files written
    by instant_angelo_b200.datasets' writers -- pins the binary layouts in both directions;
  * datasets/colmap.py normalize_poses (center camera / lookat / point, up camera), get_center, create_spheric_poses,
    error_to_confidence and the COLMAP -> OpenGL pose conversion of ColmapDatasetBase.setup.

Third-party imports of datasets/colmap.py that do not exist here (open3d, pytorch_lightning, nerfacc) are stubbed; none of
the pinned functions touches them.  Runs only in the build container.

    python tests/golden/make_golden_dataset.py
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from instant_angelo_b200 import datasets as ds  # noqa: E402


def load_reference():
    from tests.golden.make_golden import install_stubs
    install_stubs()
    sys.modules["open3d"] = types.ModuleType("open3d")
    pl = sys.modules["pytorch_lightning"]
    pl.LightningDataModule = object
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    pkg.register = lambda name: (lambda cls: cls)
    sys.modules["datasets"] = pkg
    spec = importlib.util.spec_from_file_location("datasets.colmap_utils", os.path.join(REF, "datasets", "colmap_utils.py"))
    cu = importlib.util.module_from_spec(spec)
    sys.modules["datasets.colmap_utils"] = cu
    spec.loader.exec_module(cu)
    spec = importlib.util.spec_from_file_location("datasets.colmap", os.path.join(REF, "datasets", "colmap.py"))
    cm = importlib.util.module_from_spec(spec)
    sys.modules["datasets.colmap"] = cm
    spec.loader.exec_module(cm)
    return cu, cm


def synthetic_scene(root: str, seed: int = 0):
    """8 cameras on a tilted ring looking at a blob of 400 points; 24x16 images (SIMPLE_RADIAL, as COLMAP's default)."""
    from PIL import Image as PILImage
    g = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "sparse/0"), exist_ok=True)
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    W, H = 24, 16
    cams = {1: ds.Camera(1, "SIMPLE_RADIAL", W, H, np.array([20.0, W / 2, H / 2, 0.01]))}
    pts = g.normal(size=(400, 3)) * np.array([0.6, 0.5, 0.3]) + np.array([0.3, -0.2, 0.1])
    points = {}
    for i in range(400):
        tl = int(g.integers(2, 6))
        points[i + 1] = ds.Point3D(i + 1, pts[i], g.integers(0, 255, 3).astype(np.uint8), float(g.random() * 2),
                                   g.integers(1, 9, tl).astype(np.int32), g.integers(0, 50, tl).astype(np.int32))
    images = {}
    tilt = np.array([[1, 0, 0], [0, np.cos(0.3), -np.sin(0.3)], [0, np.sin(0.3), np.cos(0.3)]])
    for i in range(8):
        a = 2 * np.pi * i / 8
        c = tilt @ np.array([3 * np.cos(a), 3 * np.sin(a), 1.0 + 0.2 * np.sin(3 * a)])
        fwd = (np.array([0.3, -0.2, 0.1]) - c); fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, tilt @ np.array([0, 0, 1.0])); right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], axis=0)                      # world -> camera (COLMAP: x right, y down, z forward)
        t = -R @ c
        # rotation matrix -> quaternion (w, x, y, z)
        w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
        x = (R[2, 1] - R[1, 2]) / (4 * w); y = (R[0, 2] - R[2, 0]) / (4 * w); z = (R[1, 0] - R[0, 1]) / (4 * w)
        n2d = int(g.integers(3, 9))
        name = f"img_{i:03d}.png"
        images[i + 1] = ds.Image(i + 1, np.array([w, x, y, z]), t, 1, name, g.random((n2d, 2)) * [W, H],
                                 g.integers(-1, 400, n2d).astype(np.int64))
        PILImage.fromarray(g.integers(0, 255, (H, W, 3)).astype(np.uint8)).save(os.path.join(root, "images", name))
    ds.write_cameras_binary(os.path.join(root, "sparse/0/cameras.bin"), cams)
    ds.write_images_binary(os.path.join(root, "sparse/0/images.bin"), images)
    ds.write_points3d_binary(os.path.join(root, "sparse/0/points3D.bin"), points)


def main():
    cu, cm = load_reference()
    root = os.path.join(HERE, "colmap_scene")
    synthetic_scene(root)
    fx = {}
    cams = cu.read_cameras_binary(os.path.join(root, "sparse/0/cameras.bin"))
    imgs = cu.read_images_binary(os.path.join(root, "sparse/0/images.bin"))
    pts = cu.read_points3d_binary(os.path.join(root, "sparse/0/points3D.bin"))
    c = cams[1]
    fx["cam"] = np.array([c.id, c.width, c.height], dtype=np.int64)
    fx["cam_model"] = np.array(c.model)
    fx["cam_params"] = np.asarray(c.params)
    fx["img_ids"] = np.array(list(imgs.keys()))
    fx["img_qvec"] = np.stack([imgs[k].qvec for k in imgs])
    fx["img_tvec"] = np.stack([imgs[k].tvec for k in imgs])
    fx["img_names"] = np.array([imgs[k].name for k in imgs])
    fx["img_xys_3"] = imgs[3].xys
    fx["img_p3d_3"] = imgs[3].point3D_ids
    fx["pt_ids"] = np.array(list(pts.keys()))
    fx["pt_xyz"] = np.stack([pts[k].xyz for k in pts])
    fx["pt_rgb"] = np.stack([pts[k].rgb for k in pts])
    fx["pt_err"] = np.array([pts[k].error for k in pts])
    fx["pt_track_img_7"] = pts[7].image_ids
    fx["pt_track_idx_7"] = pts[7].point2D_idxs
    # poses as ColmapDatasetBase.setup builds them (datasets/colmap.py:216-221)
    c2ws = []
    for d in imgs.values():
        R = d.qvec2rotmat()
        t = d.tvec.reshape(3, 1)
        c2w = torch.from_numpy(np.concatenate([R.T, -R.T @ t], axis=1)).float()
        c2w[:, 1:3] *= -1.0
        c2ws.append(c2w)
    c2ws = torch.stack(c2ws)
    fx["c2w"] = c2ws.numpy()
    p3 = torch.from_numpy(fx["pt_xyz"]).float()
    nrm = torch.nn.functional.normalize(torch.randn(p3.shape, generator=torch.Generator().manual_seed(1)), dim=-1)
    fx["normals_in"] = nrm.numpy()
    for center in ("camera", "lookat", "point"):
        poses, pts_n, nrm_n = cm.normalize_poses(c2ws.clone(), p3.clone(), up_est_method="camera", center_est_method=center,
                                                 pts3d_normal=nrm.clone())
        fx[f"norm.{center}.poses"], fx[f"norm.{center}.pts"], fx[f"norm.{center}.normals"] = poses.numpy(), pts_n.numpy(), nrm_n.numpy()
    fx["get_center"] = cm.get_center(p3).numpy()
    fx["spheric"] = cm.create_spheric_poses(torch.from_numpy(fx["norm.camera.poses"])[:, :, 3], n_steps=6).numpy()
    fx["confidence"] = np.array([cm.error_to_confidence(e) for e in fx["pt_err"]])
    np.savez_compressed(os.path.join(HERE, "colmap_dataset.npz"), **fx)
    print({k: v.shape for k, v in fx.items()})


if __name__ == "__main__":
    main()
