"""Data front-end of the path: the reference's COLMAP dataset (datasets/colmap.py, datasets/colmap_utils.py) as the
object NeuSSystem.preprocess_data reads.

  * COLMAP binary models (`sparse/0/{cameras,images,points3D}.bin`): each file is read once and decoded through numpy
    record dtypes at running offsets (tracks and 2-D point lists are variable-length, so records are walked in order),
    plus writers so that synthetic scenes and tests can produce the same files.  Field layouts are COLMAP's published
    binary format; the reference's own readers parse what the writers emit (tests/golden/make_golden_dataset.py).
  * pose conventions and normalisation: world->camera (qvec, tvec) -> camera->world [R^T | -R^T t] with the y/z axes
    flipped to OpenGL (datasets/colmap.py:218-221); `normalize_poses` (datasets/colmap.py:29-108) for the deterministic
    estimators (`center_est_method` camera / lookat / point, `up_est_method` camera); `create_spheric_poses` (:110-129).
  * `ColmapDataset`: the attribute surface of ColmapDatasetBase (all_c2w, all_images, all_fg_masks, directions, all_points,
    all_points_confidence, pts3d_normal, all_fg_indexs, all_bg_indexs, w, h, img_wh, has_mask, apply_mask), tensors placed
    on `device` so that the per-step sampling of NeuSSystem.preprocess_data indexes them where the kernels run.

Deviations, all outside what can be pinned here: `up_est_method: ground` uses the reference's own RANSAC dependency
(pyransac3d, random and not installed) -- a seeded 3-point RANSAC stands in; point normals come from open3d's k-NN PCA in
the reference (not installed) -- the same estimator is written in tensor operations (unoriented, like open3d's).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# model_id -> (name, number of parameters), COLMAP src/base/camera_models.h
CAMERA_MODELS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
                 5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
                 9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}
_MODEL_IDS = {name: (mid, n) for mid, (name, n) in CAMERA_MODELS.items()}


@dataclass
class Camera:
    id: int
    model: str
    width: int
    height: int
    params: np.ndarray


@dataclass
class Image:
    id: int
    qvec: np.ndarray          # (w, x, y, z), world -> camera
    tvec: np.ndarray
    camera_id: int
    name: str
    xys: np.ndarray           # [P, 2]
    point3D_ids: np.ndarray   # [P] int64, -1 = no 3-D point

    def qvec2rotmat(self) -> np.ndarray:
        return qvec2rotmat(self.qvec)


@dataclass
class Point3D:
    id: int
    xyz: np.ndarray
    rgb: np.ndarray
    error: float
    image_ids: np.ndarray
    point2D_idxs: np.ndarray


def qvec2rotmat(q) -> np.ndarray:
    w, x, y, z = [float(v) for v in q]
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


_CAM_HEAD = np.dtype([("id", "<i4"), ("model", "<i4"), ("w", "<u8"), ("h", "<u8")])
_IMG_HEAD = np.dtype([("id", "<i4"), ("q", "<f8", 4), ("t", "<f8", 3), ("cam", "<i4")])
_PT2D = np.dtype([("xy", "<f8", 2), ("p3d", "<i8")])
_PT3D = np.dtype([("id", "<u8"), ("xyz", "<f8", 3), ("rgb", "u1", 3), ("err", "<f8"), ("track", "<u8")])
_TRACK = np.dtype([("img", "<i4"), ("idx", "<i4")])


def read_cameras_binary(path: str) -> Dict[int, Camera]:
    buf = open(path, "rb").read()
    n, off, out = int(np.frombuffer(buf, "<u8", 1, 0)[0]), 8, {}
    for _ in range(n):
        h = np.frombuffer(buf, _CAM_HEAD, 1, off)[0]
        off += _CAM_HEAD.itemsize
        name, n_par = CAMERA_MODELS[int(h["model"])]
        params = np.frombuffer(buf, "<f8", n_par, off).copy()
        off += 8 * n_par
        out[int(h["id"])] = Camera(int(h["id"]), name, int(h["w"]), int(h["h"]), params)
    return out


def read_images_binary(path: str) -> Dict[int, Image]:
    buf = open(path, "rb").read()
    n, off, out = int(np.frombuffer(buf, "<u8", 1, 0)[0]), 8, {}
    for _ in range(n):
        h = np.frombuffer(buf, _IMG_HEAD, 1, off)[0]
        off += _IMG_HEAD.itemsize
        end = buf.index(b"\x00", off)
        name = buf[off:end].decode("utf-8")
        off = end + 1
        n2d = int(np.frombuffer(buf, "<u8", 1, off)[0])
        off += 8
        p = np.frombuffer(buf, _PT2D, n2d, off)
        off += _PT2D.itemsize * n2d
        out[int(h["id"])] = Image(int(h["id"]), h["q"].copy(), h["t"].copy(), int(h["cam"]), name, p["xy"].copy(), p["p3d"].copy())
    return out


def read_points3d_binary(path: str) -> Dict[int, Point3D]:
    buf = open(path, "rb").read()
    n, off, out = int(np.frombuffer(buf, "<u8", 1, 0)[0]), 8, {}
    for _ in range(n):
        h = np.frombuffer(buf, _PT3D, 1, off)[0]
        off += _PT3D.itemsize
        tl = int(h["track"])
        tr = np.frombuffer(buf, _TRACK, tl, off)
        off += _TRACK.itemsize * tl
        out[int(h["id"])] = Point3D(int(h["id"]), h["xyz"].copy(), h["rgb"].copy(), float(h["err"]), tr["img"].copy(), tr["idx"].copy())
    return out


def read_points3d_arrays(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(xyz [N,3], rgb [N,3], error [N]) without building per-point records: all the dataset needs from points3D.bin."""
    buf = open(path, "rb").read()
    n, off = int(np.frombuffer(buf, "<u8", 1, 0)[0]), 8
    xyz, rgb, err = np.empty((n, 3)), np.empty((n, 3), np.uint8), np.empty(n)
    for i in range(n):
        h = np.frombuffer(buf, _PT3D, 1, off)[0]
        xyz[i], rgb[i], err[i] = h["xyz"], h["rgb"], h["err"]
        off += _PT3D.itemsize + _TRACK.itemsize * int(h["track"])
    return xyz, rgb, err


def write_cameras_binary(path: str, cameras: Dict[int, Camera]) -> None:
    with open(path, "wb") as f:
        f.write(np.uint64(len(cameras)).tobytes())
        for c in cameras.values():
            mid, n_par = _MODEL_IDS[c.model]
            assert len(c.params) == n_par
            f.write(np.array([(c.id, mid, c.width, c.height)], _CAM_HEAD).tobytes())
            f.write(np.asarray(c.params, "<f8").tobytes())


def write_images_binary(path: str, images: Dict[int, Image]) -> None:
    with open(path, "wb") as f:
        f.write(np.uint64(len(images)).tobytes())
        for im in images.values():
            f.write(np.array([(im.id, im.qvec, im.tvec, im.camera_id)], _IMG_HEAD).tobytes())
            f.write(im.name.encode("utf-8") + b"\x00")
            f.write(np.uint64(len(im.point3D_ids)).tobytes())
            rec = np.empty(len(im.point3D_ids), _PT2D)
            rec["xy"], rec["p3d"] = np.asarray(im.xys).reshape(-1, 2), im.point3D_ids
            f.write(rec.tobytes())


def write_points3d_binary(path: str, points: Dict[int, Point3D]) -> None:
    with open(path, "wb") as f:
        f.write(np.uint64(len(points)).tobytes())
        for p in points.values():
            f.write(np.array([(p.id, p.xyz, p.rgb, p.error, len(p.image_ids))], _PT3D).tobytes())
            rec = np.empty(len(p.image_ids), _TRACK)
            rec["img"], rec["idx"] = p.image_ids, p.point2D_idxs
            f.write(rec.tobytes())


# ---------------------------------------------------------------------------------------------------------------------
# datasets/colmap.py:20-129
# ---------------------------------------------------------------------------------------------------------------------
def get_center(pts: torch.Tensor) -> torch.Tensor:
    """Mean of the points whose distance to the plain mean is within 1.5 sigma and 1.5 IQR (datasets/colmap.py:20-27)."""
    center = pts.mean(0)
    dis = (pts - center[None, :]).norm(p=2, dim=-1)
    mean, std = dis.mean(), dis.std()
    q25, q75 = torch.quantile(dis, 0.25), torch.quantile(dis, 0.75)
    valid = (dis > mean - 1.5 * std) & (dis < mean + 1.5 * std) & (dis > mean - (q75 - q25) * 1.5) & (dis < mean + (q75 - q25) * 1.5)
    return pts[valid].mean(0)


def _ransac_plane(pts: torch.Tensor, thresh: float = 0.01, iters: int = 1000, seed: int = 0) -> torch.Tensor:
    """Plane (A, B, C, D) with the most points within `thresh` among `iters` seeded 3-point hypotheses."""
    g = torch.Generator().manual_seed(seed)
    n = pts.shape[0]
    idx = torch.randint(0, n, (iters, 3), generator=g)
    p0, p1, p2 = pts[idx[:, 0]], pts[idx[:, 1]], pts[idx[:, 2]]
    raw = torch.cross(p1 - p0, p2 - p0, dim=-1)
    degenerate = raw.norm(dim=-1) < 1e-12          # repeated or collinear sample points span no plane
    nrm = F.normalize(raw, dim=-1)
    d = -(nrm * p0).sum(-1)
    best, best_count = None, -1
    for lo in range(0, iters, 64):
        dist = (pts @ nrm[lo:lo + 64].T + d[lo:lo + 64][None]).abs()
        cnt = (dist < thresh).sum(0).masked_fill(degenerate[lo:lo + 64], -1)
        k = int(cnt.argmax())
        if int(cnt[k]) > best_count:
            best_count, best = int(cnt[k]), torch.cat([nrm[lo + k], d[lo + k][None]])
    return best


def normalize_poses(poses: torch.Tensor, pts: torch.Tensor, up_est_method: str, center_est_method: str,
                    pts3d_normal: Optional[torch.Tensor] = None):
    """datasets/colmap.py:29-108: scene centre, up direction, rigid alignment, scale so that the closest camera is at
    distance 1.  poses [N,3,4] camera->world, pts [P,3].  Returns (poses, pts, pts3d_normal)."""
    if center_est_method in ("camera", "point"):
        center = poses[..., 3].mean(0)
    elif center_est_method == "lookat":
        cams_ori = poses[..., 3]
        cams_dir = F.normalize(poses[:, :3, :3] @ torch.as_tensor([0.0, 0.0, -1.0]), dim=-1)
        A = torch.stack([cams_dir, -cams_dir.roll(1, 0)], dim=-1)
        b = -cams_ori + cams_ori.roll(1, 0)
        t = torch.linalg.lstsq(A, b).solution
        center = (torch.stack([cams_dir, cams_dir.roll(1, 0)], dim=-1) * t[:, None, :] +
                  torch.stack([cams_ori, cams_ori.roll(1, 0)], dim=-1)).mean((0, 2))
    else:
        raise NotImplementedError(f"Unknown center estimation method: {center_est_method}")

    if up_est_method == "ground":
        plane_eq = _ransac_plane(pts, thresh=0.01)
        z = F.normalize(plane_eq[:3], dim=-1)
        signed_distance = (torch.cat([pts, torch.ones_like(pts[..., 0:1])], dim=-1) * plane_eq).sum(-1)
        if signed_distance.mean() < 0:
            z = -z
    elif up_est_method == "camera":
        z = F.normalize((poses[..., 3] - center).mean(0), dim=0)
    else:
        raise NotImplementedError(f"Unknown up estimation method: {up_est_method}")

    y_ = torch.as_tensor([z[1], -z[0], 0.0])
    x = F.normalize(torch.linalg.cross(y_, z), dim=0)
    y = torch.linalg.cross(z, x)
    row = torch.as_tensor([[0.0, 0.0, 0.0, 1.0]])
    homo = lambda p: torch.cat([p, row[None].expand(p.shape[0], -1, -1)], dim=1)
    pt_h = lambda p: torch.cat([p, torch.ones_like(p[:, 0:1])], dim=-1)[..., None]
    Rc = torch.stack([x, y, z], dim=1)
    R = Rc.T
    if center_est_method == "point":
        inv_trans = torch.cat([torch.cat([R, torch.zeros(3, 1)], dim=1), row], dim=0)
        poses_norm = (inv_trans @ homo(poses))[:, :3]
        pts = (inv_trans @ pt_h(pts))[:, :3, 0]
        poses_min, poses_max = poses_norm[..., 3].min(0)[0], poses_norm[..., 3].max(0)[0]
        pts_fg = pts[(poses_min[0] < pts[:, 0]) & (pts[:, 0] < poses_max[0]) & (poses_min[1] < pts[:, 1]) & (pts[:, 1] < poses_max[1])]
        t = -get_center(pts_fg).reshape(3, 1)
        inv_trans = torch.cat([torch.cat([torch.eye(3), t], dim=1), row], dim=0)
        poses_norm = (inv_trans @ homo(poses_norm))[:, :3]
        scale = poses_norm[..., 3].norm(p=2, dim=-1).min()
        poses_norm[..., 3] /= scale
        pts = (inv_trans @ pt_h(pts))[:, :3, 0]
    else:
        t = -R @ center.reshape(3, 1)
        inv_trans = torch.cat([torch.cat([R, t], dim=1), row], dim=0)
        poses_norm = (inv_trans @ homo(poses))[:, :3]
        scale = poses_norm[..., 3].norm(p=2, dim=-1).min()
        poses_norm[..., 3] /= scale
        pts = (inv_trans @ pt_h(pts))[:, :3, 0]
    if pts3d_normal is not None:
        pts3d_normal = (R @ pts3d_normal.T).T
    pts = pts / scale
    return poses_norm, pts, pts3d_normal


def create_spheric_poses(cameras: torch.Tensor, n_steps: int = 120) -> torch.Tensor:
    """Test trajectory: a circle at the cameras' mean height and mean distance, looking at the origin (datasets/colmap.py:110-129)."""
    center = torch.zeros(3, dtype=cameras.dtype)
    mean_d = (cameras - center[None, :]).norm(p=2, dim=-1).mean()
    mean_h = cameras[:, 2].mean()
    r = (mean_d ** 2 - mean_h ** 2).sqrt()
    up = torch.as_tensor([0.0, 0.0, 1.0], dtype=cameras.dtype)
    out = []
    for theta in torch.linspace(0, 2 * math.pi, n_steps):
        cam_pos = torch.stack([r * theta.cos(), r * theta.sin(), mean_h])
        l = F.normalize(center - cam_pos, p=2, dim=0)
        s = F.normalize(torch.linalg.cross(l, up), p=2, dim=0)
        u = F.normalize(torch.linalg.cross(s, l), p=2, dim=0)
        out.append(torch.cat([torch.stack([s, u, -l], dim=1), cam_pos[:, None]], dim=1))
    return torch.stack(out, dim=0)


def error_to_confidence(error):
    """datasets/colmap.py:149-155."""
    return 1 / (1 + np.exp(1 * np.asarray(error)))


def estimate_normals(pts: torch.Tensor, radius: float = 0.1, max_nn: int = 30, chunk: int = 4096) -> torch.Tensor:
    """Unoriented point normals: smallest-eigenvalue direction of the covariance of the up to `max_nn` nearest neighbours
    within `radius` (the estimator open3d's estimate_normals(KDTreeSearchParamHybrid(radius, max_nn)) applies at
    datasets/colmap.py:262-264), by chunked brute-force distances on the points' device."""
    n = pts.shape[0]
    out = torch.zeros_like(pts)
    k = min(max_nn, n)
    for lo in range(0, n, chunk):
        q = pts[lo:lo + chunk]
        d = torch.cdist(q, pts)
        dist, idx = d.topk(k, dim=1, largest=False)
        nb = pts[idx]                                        # [c, k, 3]
        w = (dist <= radius).to(pts.dtype)[..., None]
        cnt = w.sum(1).clamp_min(1.0)
        mean = (nb * w).sum(1) / cnt
        x = (nb - mean[:, None]) * w
        cov = x.transpose(1, 2) @ x / cnt[..., None]
        _, vec = torch.linalg.eigh(cov)
        nrm = vec[..., 0]
        few = (w.sum(1)[:, 0] < 3)
        nrm[few] = torch.as_tensor([0.0, 0.0, 1.0], dtype=pts.dtype, device=pts.device)      # open3d's default for < 3 neighbours
        out[lo:lo + chunk] = nrm
    return out


# ---------------------------------------------------------------------------------------------------------------------
# dense prior: COLMAP's dense/fused.ply (datasets/colmap.py:246-256 reads it with `plyfile`, which is not installable here)
# ---------------------------------------------------------------------------------------------------------------------
_PLY_TYPES = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
              "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
              "double": "f8", "float64": "f8"}


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """The `vertex` element of a PLY file as {property: array} (ascii, binary_little_endian and binary_big_endian; scalar
    properties only -- what COLMAP's stereo fusion and the reference's MVS preprocessing write)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements = None, []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: unterminated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[-1], None))
                else:
                    if tok[1] not in _PLY_TYPES:
                        raise ValueError(f"{path}: unknown PLY property type {tok[1]!r}")
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0][0] != "vertex":
            raise ValueError(f"{path}: the first PLY element must be `vertex`")
        _, count, props = elements[0]
        if any(t is None for _, t in props):
            raise ValueError(f"{path}: list properties on the vertex element are not supported")
        if fmt == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, max_rows=count, ndmin=2) if count else np.zeros((0, len(props)))
            return {name: rows[:, i].astype(t) for i, (name, t) in enumerate(props)}
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt!r}")
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, order + t) for name, t in props])
        rec = np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count)
        return {name: np.ascontiguousarray(rec[name]).astype(rec[name].dtype.newbyteorder("=")) for name, _ in props}


def write_ply_vertices(path: str, props: Dict[str, np.ndarray]) -> None:
    """binary_little_endian PLY with one `vertex` element (fixtures; float32 / uint8 properties)."""
    names = list(props)
    n = len(props[names[0]])
    dt = np.dtype([(k, "<u1" if props[k].dtype == np.uint8 else "<f4") for k in names])
    rec = np.zeros(n, dtype=dt)
    for k in names:
        rec[k] = props[k]
    with open(path, "wb") as f:
        f.write(b"ply\nformat binary_little_endian 1.0\n")
        f.write(f"element vertex {n}\n".encode())
        for k in names:
            f.write(f"property {'uchar' if props[k].dtype == np.uint8 else 'float'} {k}\n".encode())
        f.write(b"end_header\n")
        f.write(rec.tobytes())


def load_dense_prior(path: str) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """datasets/colmap.py:246-256: points, the PLY's own normals and its `confidence` column (ones when absent)."""
    assert os.path.exists(path), f"Please check whether {path} exists"
    v = read_ply_vertices(path)
    for k in ("x", "y", "z", "nx", "ny", "nz"):
        if k not in v:
            raise KeyError(f"{path}: vertex property {k!r} missing (the dense prior needs positions and normals)")
    pts = torch.from_numpy(np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float32))
    nrm = torch.from_numpy(np.stack([v["nx"], v["ny"], v["nz"]], axis=1).astype(np.float32))
    conf = torch.from_numpy(v["confidence"].astype(np.float32)) if "confidence" in v else torch.ones(pts.shape[0])
    return pts, nrm, conf


def colmap_to_c2w(qvec, tvec) -> torch.Tensor:
    """world->camera (q, t) of images.bin -> camera->world [3,4] in the OpenGL convention (datasets/colmap.py:217-221)."""
    R = qvec2rotmat(qvec)
    t = np.asarray(tvec, dtype=np.float64).reshape(3, 1)
    c2w = torch.from_numpy(np.concatenate([R.T, -R.T @ t], axis=1)).float()
    c2w[:, 1:3] *= -1.0
    return c2w


def intrinsics(cam: Camera, factor: float):
    """(fx, fy, cx, cy) scaled by `factor` for the models the reference parses (datasets/colmap.py:181-198)."""
    p = cam.params
    if cam.model in ("SIMPLE_RADIAL", "SIMPLE_PINHOLE"):
        return p[0] * factor, p[0] * factor, p[1] * factor, p[2] * factor
    if cam.model in ("PINHOLE", "OPENCV"):
        return p[0] * factor, p[1] * factor, p[2] * factor, p[3] * factor
    raise ValueError(f"Please parse the intrinsics for camera model {cam.model}!")


class ColmapDataset:
    """ColmapDatasetBase.setup (datasets/colmap.py:157-316) for split 'train' / 'val' / 'test'."""

    def __init__(self, config, split: str = "train", device="cpu"):
        from PIL import Image as PILImage
        from .systems import get_ray_directions
        self.config, self.split = config, split
        root = config["root_dir"]
        cams = read_cameras_binary(os.path.join(root, "sparse/0/cameras.bin"))
        cam = cams[1]
        H, W = int(cam.height), int(cam.width)
        if "img_wh" in config:
            w, h = config["img_wh"]
            assert round(W / w * h) == H
        elif "img_downscale" in config:
            w, h = int(W / config["img_downscale"] + 0.5), int(H / config["img_downscale"] + 0.5)
        else:
            raise KeyError("Either img_wh or img_downscale should be specified.")
        self.w, self.h, self.img_wh, self.factor = w, h, (w, h), w / W
        fx, fy, cx, cy = intrinsics(cam, self.factor)
        self.directions = get_ray_directions(w, h, fx, fy, cx, cy)
        imdata = read_images_binary(os.path.join(root, "sparse/0/images.bin"))
        mask_dir = os.path.join(root, "mask")
        self.has_mask = os.path.exists(mask_dir)
        self.apply_mask = self.has_mask and bool(config.get("apply_mask", False))
        c2ws, images, masks, fg_idx, bg_idx = [], [], [], [], []
        for i, d in enumerate(imdata.values()):
            c2ws.append(colmap_to_c2w(d.qvec, d.tvec))
            if split in ("train", "val"):
                img = PILImage.open(os.path.join(root, "images", d.name)).resize((w, h), PILImage.BICUBIC)
                img = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0)
                img = img[..., None].expand(-1, -1, 3) if img.ndim == 2 else img[..., :3]
                if self.has_mask:
                    cands = [p for p in (os.path.join(mask_dir, d.name), os.path.join(mask_dir, d.name[3:])) if os.path.exists(p)]
                    assert len(cands) == 1
                    m = PILImage.open(cands[0]).convert("L").resize((w, h), PILImage.BICUBIC)
                    mask = torch.from_numpy(np.asarray(m, dtype=np.float32) / 255.0)
                else:
                    mask = torch.ones(h, w)
                nz, z = torch.nonzero(mask.bool()), torch.nonzero(~mask.bool())
                fg_idx.append(torch.cat([torch.full((nz.shape[0], 1), i), nz], dim=1))
                bg_idx.append(torch.cat([torch.full((z.shape[0], 1), i), z], dim=1))
                images.append(img.contiguous())
                masks.append(mask)
        all_c2w = torch.stack(c2ws, dim=0)
        if config.get("dense_pcd_path", None) is not None:
            # dense MVS prior with its own oriented normals and confidences (datasets/colmap.py:246-256)
            pts3d, normals, conf = load_dense_prior(os.path.join(root, config["dense_pcd_path"]))
        else:
            xyz, _, err = read_points3d_arrays(os.path.join(root, "sparse/0/points3D.bin"))
            pts3d = torch.from_numpy(xyz).float()
            conf = torch.from_numpy(error_to_confidence(err)).float()
            normals = estimate_normals(pts3d, radius=0.1, max_nn=30)
        all_c2w, pts3d, normals = normalize_poses(all_c2w, pts3d, config["up_est_method"], config["center_est_method"], normals)
        if split == "test":
            n = int(config["n_test_traj_steps"])
            self.all_c2w = create_spheric_poses(all_c2w[:, :, 3], n_steps=n)
            self.all_images, self.all_fg_masks = torch.zeros(n, h, w, 3), torch.zeros(n, h, w)
            self.all_points, self.all_points_confidence = torch.tensor([]), torch.tensor([])
            self.all_fg_indexs, self.all_bg_indexs = torch.tensor([]), torch.tensor([])
        else:
            self.all_c2w = all_c2w
            self.all_images, self.all_fg_masks = torch.stack(images).float(), torch.stack(masks).float()
            self.all_points, self.all_points_confidence = pts3d, conf
            self.all_fg_indexs, self.all_bg_indexs = torch.cat(fg_idx, dim=0), torch.cat(bg_idx, dim=0)
        self.pts3d_normal = normals.float()
        for k in ("directions", "all_c2w", "all_images", "all_fg_masks", "all_points", "all_points_confidence", "pts3d_normal",
                  "all_fg_indexs", "all_bg_indexs"):
            setattr(self, k, getattr(self, k).float().to(device) if getattr(self, k).is_floating_point() else getattr(self, k).to(device))

    def __len__(self):
        return len(self.all_images)

    def __getitem__(self, index):
        return {"index": index}
