"""CPU: marching cubes of the export path (instant_angelo_b200/isosurface.py; reference models/geometry.py:33-113).
PyMCubes is not installable here, so the implementation is held to what defines a correct iso-surface: closed,
consistently oriented 2-manifolds with the right topology, vertices on the iso-level, normals towards increasing
values, and blocks that tile without gaps."""
import math

import numpy as np
import pytest
import torch

from instant_angelo_b200 import isosurface as iso


def _grid(n, lo=-1.0, hi=1.0):
    c = torch.linspace(lo, hi, n)
    return torch.meshgrid(c, c, c, indexing="ij")


def _manifold_stats(verts, faces):
    """Returns (closed & oriented, Euler characteristic).  Closed + consistently oriented: every directed edge occurs
    exactly once and its reverse exactly once."""
    f = faces.numpy()
    d = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    assert (d[:, 0] != d[:, 1]).all(), "degenerate triangle (repeated vertex index)"
    key = d[:, 0].astype(np.int64) * (verts.shape[0] + 1) + d[:, 1]
    rkey = d[:, 1].astype(np.int64) * (verts.shape[0] + 1) + d[:, 0]
    uk, cnt = np.unique(key, return_counts=True)
    closed = (cnt == 1).all() and np.array_equal(np.sort(rkey), uk)
    n_edges = len(uk) // 2
    used = np.unique(f).size
    return bool(closed), used - n_edges + f.shape[0]


def _signed_volume(verts, faces):
    v = verts.double()[faces]
    return float((v[:, 0] * torch.cross(v[:, 1], v[:, 2], dim=-1)).sum() / 6.0)


def test_generated_table_is_consistent():
    assert len(iso._TABLE) == 256 and iso._TABLE[0] == [] and iso._TABLE[255] == []
    for code in range(256):
        inside = [(code >> c) & 1 for c in range(8)]
        crossing = {e for e, (a, b) in enumerate(iso._EDGES) if inside[a] != inside[b]}
        used = {e for t in iso._TABLE[code] for e in t}
        assert used == crossing, code                    # every sign change carries a vertex, nothing else does
    # a single inside corner is one triangle; the complement has the same triangle with the opposite orientation
    assert len(iso._TABLE[1]) == 1 and len(iso._TABLE[254]) == 1
    assert sorted(iso._TABLE[1][0]) == sorted(iso._TABLE[254][0]) and iso._TABLE[1][0] != iso._TABLE[254][0]
    assert iso._MAX_TRIS <= 12


def test_sphere_is_a_closed_outward_oriented_genus0_surface():
    x, y, z = _grid(41)
    r = 0.6
    vol = torch.sqrt(x * x + y * y + z * z) - r
    verts, faces = iso.marching_cubes(vol, 0.0)
    closed, chi = _manifold_stats(verts, faces)
    assert closed and chi == 2
    world = verts / 40.0 * 2.0 - 1.0
    assert float((world.norm(dim=-1) - r).abs().max()) < 2e-3          # linear interpolation of a distance field
    vol_mesh = _signed_volume(world, faces)
    assert vol_mesh > 0 and abs(vol_mesh - 4.0 / 3.0 * math.pi * r ** 3) / (4.0 / 3.0 * math.pi * r ** 3) < 0.01
    # non-zero iso value moves the surface outwards
    v2, f2 = iso.marching_cubes(vol, 0.1)
    assert abs(float((v2 / 40.0 * 2.0 - 1.0).norm(dim=-1).mean()) - 0.7) < 2e-3


def test_topology_torus_and_two_components():
    x, y, z = _grid(49)
    torus = torch.sqrt((torch.sqrt(x * x + y * y) - 0.6) ** 2 + z * z) - 0.22
    verts, faces = iso.marching_cubes(torus)
    closed, chi = _manifold_stats(verts, faces)
    assert closed and chi == 0
    two = torch.minimum(torch.sqrt((x - 0.45) ** 2 + y * y + z * z), torch.sqrt((x + 0.45) ** 2 + y * y + z * z)) - 0.3
    verts, faces = iso.marching_cubes(two)
    closed, chi = _manifold_stats(verts, faces)
    assert closed and chi == 4


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ambiguous_faces_stay_watertight(seed):
    """White noise makes every cube configuration and ambiguous face occur; the pairing rule is per face, so the surface
    must still be closed and consistently oriented (padding keeps it away from the volume boundary)."""
    g = torch.Generator().manual_seed(seed)
    vol = torch.ones(22, 22, 22)
    vol[2:-2, 2:-2, 2:-2] = torch.randn(18, 18, 18, generator=g)
    verts, faces = iso.marching_cubes(vol, 0.0)
    codes = set()
    inside = vol < 0
    code = torch.zeros(21, 21, 21, dtype=torch.int64)
    for c, (cx, cy, cz) in enumerate(iso._CORNERS):
        code |= inside[cx:21 + cx, cy:21 + cy, cz:21 + cz].to(torch.int64) << c
    assert len(torch.unique(code)) > 250                      # practically all 256 configurations are exercised
    closed, _ = _manifold_stats(verts, faces)
    assert closed
    # every vertex sits on a grid edge at the interpolated crossing
    frac = verts - verts.floor()
    assert ((frac > 0).sum(dim=-1) <= 1).all()


def test_empty_and_full_volumes():
    v, f = iso.marching_cubes(torch.ones(5, 5, 5))
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = iso.marching_cubes(-torch.ones(5, 5, 5))
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = iso.marching_cubes(torch.ones(1, 4, 4))
    assert f.shape == (0, 3)


def test_helper_blocks_tile_the_lattice_like_the_reference():
    """MarchingCubeHelper (models/geometry.py:36-113): sdf_func returns the SDF with its sign flipped; blocks of
    block_res + 1 samples share their boundary layer, so block-wise extraction covers the same surface."""
    r = 0.55
    sdf_func = lambda p: -(p.norm(dim=-1) - r)
    bounds = [(-1.0, 1.0)] * 3
    one = iso.MarchingCubeHelper(sdf_func, bounds, resolution=32, block_res=256)(threshold=0.001)
    many = iso.MarchingCubeHelper(sdf_func, bounds, resolution=32, block_res=8)
    assert (many.num_blocks_x, many.num_blocks_y, many.num_blocks_z) == (4, 4, 4)
    tiled = many(threshold=0.001)
    assert one["t_pos_idx"].dtype == torch.int64 and one["v_pos"].shape[1] == 3
    assert tiled["t_pos_idx"].shape[0] == one["t_pos_idx"].shape[0]            # same triangles, vertices duplicated at block seams
    assert tiled["v_pos"].shape[0] > one["v_pos"].shape[0]
    for mesh in (one, tiled):
        assert float((mesh["v_pos"].norm(dim=-1) - (r + 0.001)).abs().max()) < 3e-3
        vol = _signed_volume(mesh["v_pos"], mesh["t_pos_idx"])
        assert abs(vol - 4.0 / 3.0 * math.pi * (r + 0.001) ** 3) / vol < 0.02
    closed, chi = _manifold_stats(one["v_pos"], one["t_pos_idx"])
    assert closed and chi == 2


def test_save_obj(tmp_path):
    x, y, z = _grid(9)
    verts, faces = iso.marching_cubes(torch.sqrt(x * x + y * y + z * z) - 0.5)
    path = tmp_path / "m.obj"
    iso.save_obj(str(path), verts, faces, v_rgb=torch.rand(verts.shape[0], 3))
    lines = path.read_text().splitlines()
    assert sum(l.startswith("v ") for l in lines) == verts.shape[0] and sum(l.startswith("f ") for l in lines) == faces.shape[0]
    assert len(lines[0].split()) == 7 and min(int(t) for l in lines if l.startswith("f ") for t in l.split()[1:]) == 1


def test_geometry_isosurface_and_model_export_wiring(monkeypatch):
    """BaseImplicitGeometry.isosurface / NeuSModel.export (reference models/geometry.py:80-113, models/neus.py:308-318) with
    the network evaluations replaced by an analytic sphere (the kernels need a GPU): lattice, sign convention,
    threshold 0.001, chunked vertex attributes."""
    from instant_angelo_b200 import configs, make
    from instant_angelo_b200.config import to_config
    cfg = configs.neuralangelo_colmap_sparse("finite_difference")
    for blk in (cfg.model.geometry, cfg.model.geometry_bg):
        blk.xyz_encoding_config["log2_hashmap_size"] = 8
    cfg.model.geometry.isosurface["resolution"] = 24          # 2 / 24 spacing over [-1.5, 1.5): 36 samples per axis
    cfg.model.geometry.isosurface["block_res"] = 16
    model = make("neus", cfg.model)
    geo = model.geometry
    r0 = 0.8
    monkeypatch.setattr(geo, "forward_level", lambda p: p.norm(dim=-1) - r0)
    mesh = model.isosurface()
    assert set(mesh) == {"v_pos", "t_pos_idx"} and mesh["t_pos_idx"].dtype == torch.int64
    assert float((mesh["v_pos"].norm(dim=-1) - (r0 + 0.001)).abs().max()) < 5e-3          # iso level 0.001, as the reference
    assert _signed_volume(mesh["v_pos"], mesh["t_pos_idx"]) > 0                           # normals towards increasing SDF
    calls = []

    def fake_forward(pts, with_grad=True, with_feature=True, **kw):
        calls.append(pts.shape[0])
        n = torch.nn.functional.normalize(pts, dim=-1)
        return pts.norm(dim=-1) - r0, 2.0 * n, torch.cat([pts.new_zeros(pts.shape[0], 1), pts, pts.new_zeros(pts.shape[0], 64)], dim=-1)

    monkeypatch.setattr(geo, "forward", fake_forward)
    out = model.export(to_config({"export_vertex_color": True, "chunk_size": 1000}))
    nv = out["v_pos"].shape[0]
    assert calls == [1000] * (nv // 1000) + ([nv % 1000] if nv % 1000 else [])
    assert out["v_rgb"].shape == (nv, 3) and out["v_norm"].shape == (nv, 3)
    assert torch.allclose(out["v_rgb"], torch.sigmoid(out["v_pos"]), atol=1e-6)           # features[..., 1:4]
    assert torch.allclose(out["v_norm"].norm(dim=-1), torch.ones(nv), atol=1e-5)
    no_col = model.export(to_config({"export_vertex_color": False}))
    assert "v_rgb" not in no_col
    cfg.model.geometry_bg["isosurface"] = None
    with pytest.raises(NotImplementedError):
        model.geometry_bg.isosurface()
