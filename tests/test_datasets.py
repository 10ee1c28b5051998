"""CPU: the COLMAP data front-end (instant_angelo_b200/datasets.py) against fixtures produced by the reference's
datasets/colmap_utils.py and datasets/colmap.py (tests/golden/make_golden_dataset.py) on a tiny synthetic COLMAP model
(tests/golden/colmap_scene)."""
import os

import numpy as np
import pytest
import torch

from instant_angelo_b200 import datasets as ds
from instant_angelo_b200.config import to_config

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SCENE = os.path.join(GOLD, "colmap_scene")


@pytest.fixture(scope="module")
def fx():
    z = np.load(os.path.join(GOLD, "colmap_dataset.npz"))
    return {k: z[k] for k in z.files}


def test_binary_readers_match_reference_readers(fx):
    cams = ds.read_cameras_binary(os.path.join(SCENE, "sparse/0/cameras.bin"))
    assert list(cams) == [1]
    c = cams[1]
    assert [c.id, c.width, c.height] == fx["cam"].tolist() and c.model == str(fx["cam_model"])
    assert np.array_equal(c.params, fx["cam_params"])
    imgs = ds.read_images_binary(os.path.join(SCENE, "sparse/0/images.bin"))
    assert list(imgs) == fx["img_ids"].tolist()
    assert np.array_equal(np.stack([i.qvec for i in imgs.values()]), fx["img_qvec"])
    assert np.array_equal(np.stack([i.tvec for i in imgs.values()]), fx["img_tvec"])
    assert [i.name for i in imgs.values()] == fx["img_names"].tolist()
    assert np.array_equal(imgs[3].xys, fx["img_xys_3"]) and np.array_equal(imgs[3].point3D_ids, fx["img_p3d_3"])
    pts = ds.read_points3d_binary(os.path.join(SCENE, "sparse/0/points3D.bin"))
    assert list(pts) == fx["pt_ids"].tolist()
    assert np.array_equal(np.stack([p.xyz for p in pts.values()]), fx["pt_xyz"])
    assert np.array_equal(np.stack([p.rgb for p in pts.values()]), fx["pt_rgb"])
    assert np.array_equal(np.array([p.error for p in pts.values()]), fx["pt_err"])
    assert np.array_equal(pts[7].image_ids, fx["pt_track_img_7"]) and np.array_equal(pts[7].point2D_idxs, fx["pt_track_idx_7"])
    xyz, rgb, err = ds.read_points3d_arrays(os.path.join(SCENE, "sparse/0/points3D.bin"))
    assert np.array_equal(xyz, fx["pt_xyz"]) and np.array_equal(rgb, fx["pt_rgb"]) and np.array_equal(err, fx["pt_err"])


def test_writers_round_trip(tmp_path):
    src = os.path.join(SCENE, "sparse/0")
    for name, rd, wr in (("cameras.bin", ds.read_cameras_binary, ds.write_cameras_binary),
                         ("images.bin", ds.read_images_binary, ds.write_images_binary),
                         ("points3D.bin", ds.read_points3d_binary, ds.write_points3d_binary)):
        out = tmp_path / name
        wr(str(out), rd(os.path.join(src, name)))
        assert out.read_bytes() == open(os.path.join(src, name), "rb").read(), name


def test_pose_conversion_and_normalisation_match_reference(fx):
    imgs = ds.read_images_binary(os.path.join(SCENE, "sparse/0/images.bin"))
    c2w = torch.stack([ds.colmap_to_c2w(i.qvec, i.tvec) for i in imgs.values()])
    assert np.array_equal(c2w.numpy(), fx["c2w"])
    # camera looks along -z of its own frame towards the scene, OpenGL convention
    look = -c2w[:, :, 2]
    to_scene = torch.nn.functional.normalize(torch.tensor([0.3, -0.2, 0.1]) - c2w[:, :, 3], dim=-1)
    assert float((look * to_scene).sum(-1).min()) > 0.999
    pts = torch.from_numpy(fx["pt_xyz"]).float()
    nrm = torch.from_numpy(fx["normals_in"])
    for center in ("camera", "lookat", "point"):
        poses, p, n = ds.normalize_poses(c2w.clone(), pts.clone(), "camera", center, nrm.clone())
        for got, key in ((poses, "poses"), (p, "pts"), (n, "normals")):
            want = fx[f"norm.{center}.{key}"]
            assert np.allclose(got.numpy(), want, rtol=1e-5, atol=1e-6), (center, key, np.abs(got.numpy() - want).max())
        assert abs(float(poses[..., 3].norm(dim=-1).min()) - 1.0) < 1e-5          # closest camera at distance 1
    assert np.allclose(ds.get_center(pts).numpy(), fx["get_center"], atol=1e-6)
    sph = ds.create_spheric_poses(torch.from_numpy(fx["norm.camera.poses"])[:, :, 3], n_steps=6)
    assert np.allclose(sph.numpy(), fx["spheric"], atol=1e-6)
    assert np.allclose(ds.error_to_confidence(fx["pt_err"]), fx["confidence"])
    with pytest.raises(NotImplementedError):
        ds.normalize_poses(c2w, pts, "camera", "nope")


def test_ground_plane_up_and_normals():
    g = torch.Generator().manual_seed(0)
    plane = torch.cat([torch.rand(600, 2, generator=g) * 4 - 2, torch.randn(600, 1, generator=g) * 0.002], dim=1)
    blob = torch.randn(150, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 0.8])
    pts = torch.cat([plane, blob])
    eq = ds._ransac_plane(pts, thresh=0.01)
    assert abs(float(eq[2].abs()) - 1.0) < 1e-3 and abs(float(eq[3])) < 0.01
    cams = torch.eye(3, 4)[None].repeat(4, 1, 1)
    cams[:, :, 3] = torch.tensor([[2.0, 0, 1.5], [-2.0, 0, 1.5], [0, 2.0, 1.5], [0, -2.0, 1.5]])
    poses, p, _ = ds.normalize_poses(cams, pts, "ground", "camera")
    assert float(p[:600, 2].std()) < 0.01 and float(p[600:, 2].mean()) < float(p[:600, 2].mean()) + 1.0   # ground stays flat in z
    # k-NN PCA normals: a sampled plane has normals +-z, a sphere has radial normals (unoriented)
    n_plane = ds.estimate_normals(plane, radius=0.5, max_nn=30)
    assert float(n_plane[:, 2].abs().min()) > 0.99
    sph = torch.nn.functional.normalize(torch.randn(2000, 3, generator=g), dim=-1)
    n_sph = ds.estimate_normals(sph, radius=0.3, max_nn=30)
    assert float((n_sph * sph).sum(-1).abs().mean()) > 0.98
    lonely = ds.estimate_normals(torch.tensor([[0.0, 0, 0], [5.0, 5, 5]]), radius=0.1)
    assert torch.equal(lonely, torch.tensor([[0.0, 0, 1], [0.0, 0, 1]]))


def test_colmap_dataset_feeds_preprocess_data():
    import types
    from instant_angelo_b200.systems import NeuSSystem
    from instant_angelo_b200 import configs
    dcfg = to_config({"name": "colmap", "root_dir": SCENE, "img_downscale": 2, "up_est_method": "camera", "center_est_method": "lookat",
                      "n_test_traj_steps": 5, "apply_mask": False})
    d = ds.ColmapDataset(dcfg, "train")
    assert (d.w, d.h) == (12, 8) and d.factor == 0.5 and not d.has_mask and not d.apply_mask
    assert d.all_images.shape == (8, 8, 12, 3) and d.all_fg_masks.shape == (8, 8, 12) and d.directions.shape == (8, 12, 3)
    assert d.all_c2w.shape == (8, 3, 4) and d.all_points.shape == (400, 3) and d.pts3d_normal.shape == (400, 3)
    assert d.all_fg_indexs.shape == (8 * 8 * 12, 3) and d.all_bg_indexs.shape == (0, 3)
    assert 0.0 <= float(d.all_images.min()) and float(d.all_images.max()) <= 1.0 and len(d) == 8
    assert abs(float(d.all_c2w[..., 3].norm(dim=-1).min()) - 1.0) < 1e-5
    assert torch.all((d.all_points_confidence > 0) & (d.all_points_confidence <= 0.5))
    # intrinsics scaled by the down-scale factor: the principal ray of the centre pixel is (almost) -z
    assert torch.allclose(d.directions[4, 6], torch.tensor([0.05, -0.05, -1.0]), atol=1e-6)
    cfg = configs.neuralangelo_colmap_sparse()
    system = NeuSSystem(cfg, dataset=d, model=types.SimpleNamespace(background_color=None))
    b = {}
    torch.manual_seed(0)
    system.preprocess_data(b, "train")
    assert b["rays"].shape == (256, 6) and b["rgb"].shape == (256, 3) and b["pts"].shape == (256, 3) and b["pts_normal"].shape == (256, 3)
    assert torch.allclose(b["rays"][:, 3:].norm(dim=-1), torch.ones(256), atol=1e-5)
    t = ds.ColmapDataset(dcfg, "test")
    assert t.all_c2w.shape == (5, 3, 4) and t.all_images.shape == (5, 8, 12, 3) and t.all_points.numel() == 0
    v = {"index": torch.tensor([2])}
    system.dataset = t
    system.preprocess_data(v, "test")
    assert v["rays"].shape == (96, 6)
