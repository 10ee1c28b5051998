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


def test_ply_reader_formats(tmp_path):
    g = np.random.default_rng(0)
    props = {"x": g.normal(size=50).astype(np.float32), "y": g.normal(size=50).astype(np.float32),
             "z": g.normal(size=50).astype(np.float32), "nx": g.normal(size=50).astype(np.float32),
             "ny": g.normal(size=50).astype(np.float32), "nz": g.normal(size=50).astype(np.float32),
             "red": g.integers(0, 255, 50).astype(np.uint8), "confidence": g.random(50).astype(np.float32)}
    p = str(tmp_path / "le.ply")
    ds.write_ply_vertices(p, props)
    got = ds.read_ply_vertices(p)
    assert list(got) == list(props)
    for k in props:
        assert np.array_equal(got[k], props[k]) and got[k].dtype == props[k].dtype, k
    # ascii, with a comment and a trailing face element that must be ignored
    a = tmp_path / "a.ply"
    lines = ["ply", "format ascii 1.0", "comment made by hand", "element vertex 3", "property float x", "property float y",
             "property double z", "property uchar red", "element face 1", "property list uchar int vertex_indices", "end_header",
             "0 1 2.5 7", "1 0 -2 255", "0.5 0.5 0.25 0", "3 0 1 2"]
    a.write_text("\n".join(lines) + "\n")
    got = ds.read_ply_vertices(str(a))
    assert got["z"].dtype == np.float64 and got["red"].tolist() == [7, 255, 0] and got["x"].tolist() == [0.0, 1.0, 0.5]
    # big endian
    be = tmp_path / "be.ply"
    rec = np.zeros(4, dtype=np.dtype([("x", ">f4"), ("y", ">f4"), ("z", ">f4")]))
    rec["x"], rec["y"], rec["z"] = [1, 2, 3, 4], [5, 6, 7, 8], [-1, -2, -3, -4]
    be.write_bytes(b"ply\nformat binary_big_endian 1.0\nelement vertex 4\nproperty float x\nproperty float y\nproperty float z\nend_header\n" + rec.tobytes())
    got = ds.read_ply_vertices(str(be))
    assert got["z"].tolist() == [-1.0, -2.0, -3.0, -4.0]
    (tmp_path / "bad.ply").write_bytes(b"plx\n")
    with pytest.raises(ValueError):
        ds.read_ply_vertices(str(tmp_path / "bad.ply"))


def test_dense_prior_branch(tmp_path):
    """config.dense_pcd_path (configs/neuralangelo-colmap_dense.yaml: dense/fused.ply): points, the PLY's own normals
    and confidences replace the sparse points + estimated normals (reference datasets/colmap.py:246-256)."""
    root = tmp_path / "scene"
    root.mkdir()
    for sub in ("sparse", "images"):
        os.symlink(os.path.join(SCENE, sub), root / sub)
    (root / "dense").mkdir()
    g = np.random.default_rng(1)
    n = 37
    xyz = g.normal(size=(n, 3)).astype(np.float32) * 0.3 + np.array([0.3, -0.2, 0.1], np.float32)
    nrm = g.normal(size=(n, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    conf = g.random(n).astype(np.float32)
    ds.write_ply_vertices(str(root / "dense" / "fused.ply"),
                          {"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2], "nx": nrm[:, 0], "ny": nrm[:, 1], "nz": nrm[:, 2],
                           "confidence": conf})
    base = {"name": "colmap", "root_dir": str(root), "img_downscale": 2, "up_est_method": "camera", "center_est_method": "lookat",
            "n_test_traj_steps": 5, "apply_mask": False}
    d = ds.ColmapDataset(to_config({**base, "dense_pcd_path": "dense/fused.ply"}), "train")
    assert d.all_points.shape == (n, 3) and d.pts3d_normal.shape == (n, 3)
    assert np.allclose(d.all_points_confidence.numpy(), conf)
    # the same rigid normalisation the sparse branch applies: normals are the PLY's, rotated, still unit length
    imgs = ds.read_images_binary(os.path.join(SCENE, "sparse/0/images.bin"))
    c2w = torch.stack([ds.colmap_to_c2w(i.qvec, i.tvec) for i in imgs.values()])
    _, want_p, want_n = ds.normalize_poses(c2w, torch.from_numpy(xyz), "camera", "lookat", torch.from_numpy(nrm))
    assert torch.allclose(d.all_points, want_p.float(), atol=1e-6) and torch.allclose(d.pts3d_normal, want_n.float(), atol=1e-6)
    assert torch.allclose(d.pts3d_normal.norm(dim=-1), torch.ones(n), atol=1e-5)
    # without a confidence column: ones
    ds.write_ply_vertices(str(root / "dense" / "noconf.ply"),
                          {"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2], "nx": nrm[:, 0], "ny": nrm[:, 1], "nz": nrm[:, 2]})
    d2 = ds.ColmapDataset(to_config({**base, "dense_pcd_path": "dense/noconf.ply"}), "train")
    assert torch.equal(d2.all_points_confidence, torch.ones(n))
    with pytest.raises(AssertionError, match="exists"):
        ds.ColmapDataset(to_config({**base, "dense_pcd_path": "dense/missing.ply"}), "train")
    # sparse branch unchanged when the key is null (the sparse YAMLs)
    d3 = ds.ColmapDataset(to_config({**base, "dense_pcd_path": None}), "train")
    assert d3.all_points.shape == (400, 3)
