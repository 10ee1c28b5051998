"""Iso-surface extraction for the export path: reference models/geometry.py:33-113 (MarchingCubeHelper,
BaseImplicitGeometry.isosurface) and models/neus.py:308-318 (NeuSModel.export).

The reference hands the sampled SDF volume to PyMCubes / CuMCubes (neither is installable here).  This module is a
marching-cubes implementation written from the definition, in tensor operations that run on the device that holds the
volume (the SDF blocks come straight from VolumeSDF.forward_level, so nothing is copied to the host before the mesh exists):

  * the per-configuration triangle table is GENERATED at import, not transcribed: on every cube face the sign changes
    are paired into segments (two crossings: one segment; four crossings, the ambiguous face: each inside corner is cut
    off on its own), the segments close into loops on the cube surface, each loop is oriented towards increasing values
    and fan-triangulated.  The pairing depends on the signs of the face's four corners only, so the two cubes that share a
    face agree on it and the mesh is watertight by construction (tests/test_isosurface.py checks closed, consistently
    oriented 2-manifolds on spheres, a torus and ambiguous-face volumes);
  * vertices are shared inside a block (one vertex per crossed grid edge), blocks are concatenated as the reference
    concatenates its per-block meshes.

Outputs follow PyMCubes: vertices in index units of the sampled volume (the helper rescales to world coordinates as
models/geometry.py:99 does), int64 triangles, surface at `volume == isovalue`, normals towards increasing values
(outwards for an SDF).  Triangulation inside a cube differs from PyMCubes' table where the table is a matter of choice.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Tuple

import torch

# ---- cube conventions --------------------------------------------------------------------------------------------
# corner c = (c & 1, (c >> 1) & 1, (c >> 2) & 1) = (x, y, z); edge e = 4 * axis + rank of its lower corner among the four
# corners with bit `axis` clear.
_CORNERS = [((c >> 0) & 1, (c >> 1) & 1, (c >> 2) & 1) for c in range(8)]
_EDGES: List[Tuple[int, int]] = []
for _axis in range(3):
    for _c in range(8):
        if not (_c >> _axis) & 1:
            _EDGES.append((_c, _c | (1 << _axis)))
_EDGE_ID = {e: i for i, e in enumerate(_EDGES)}


def _face_cycles():
    faces = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in (0, 1):
            cyc = []
            for (du, dv) in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = (side << axis) | (du << u) | (dv << v)
                cyc.append(c)
            faces.append(cyc)
    return faces


_FACES = _face_cycles()


def _edge_between(a: int, b: int) -> int:
    return _EDGE_ID[(min(a, b), max(a, b))]


def _build_table():
    table: List[List[Tuple[int, int, int]]] = []
    for code in range(256):
        inside = [(code >> c) & 1 == 1 for c in range(8)]
        link: Dict[int, List[int]] = {}

        def connect(e0, e1):
            link.setdefault(e0, []).append(e1)
            link.setdefault(e1, []).append(e0)

        for cyc in _FACES:
            cross = [k for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]      # face edge k: corner k -> k+1
            fe = lambda k: _edge_between(cyc[k % 4], cyc[(k + 1) % 4])
            if len(cross) == 2:
                connect(fe(cross[0]), fe(cross[1]))
            elif len(cross) == 4:
                for k in range(4):                      # ambiguous face: cut off each inside corner on its own
                    if inside[cyc[k]]:
                        connect(fe(k - 1), fe(k))
        tris: List[Tuple[int, int, int]] = []
        seen = set()
        for start in sorted(link):
            if start in seen:
                continue
            assert len(link[start]) == 2
            loop, prev, cur = [start], None, start
            while True:
                seen.add(cur)
                a, b = link[cur]
                nxt = a if a != prev else b
                if prev is None:
                    nxt = a
                if nxt == start:
                    break
                prev, cur = cur, nxt
                loop.append(cur)
            # orientation: normal towards the outside (larger values) end of the crossed edges
            mid = [tuple((_CORNERS[_EDGES[e][0]][d] + _CORNERS[_EDGES[e][1]][d]) / 2.0 for d in range(3)) for e in loop]
            n = [0.0, 0.0, 0.0]
            for i in range(len(mid)):                   # Newell
                p, q = mid[i], mid[(i + 1) % len(mid)]
                n[0] += (p[1] - q[1]) * (p[2] + q[2])
                n[1] += (p[2] - q[2]) * (p[0] + q[0])
                n[2] += (p[0] - q[0]) * (p[1] + q[1])
            score = 0.0
            for e in loop:
                a, b = _EDGES[e]
                i_c, o_c = (a, b) if inside[a] else (b, a)
                score += sum(n[d] * (_CORNERS[o_c][d] - _CORNERS[i_c][d]) for d in range(3))
            assert abs(score) > 1e-9, (code, loop)
            if score < 0:
                loop = loop[::-1]
            for i in range(1, len(loop) - 1):
                tris.append((loop[0], loop[i], loop[i + 1]))
        table.append(tris)
    return table


_TABLE = _build_table()
_MAX_TRIS = max(len(t) for t in _TABLE)
_TRI_TABLE = torch.full((256, _MAX_TRIS, 3), -1, dtype=torch.int64)
for _code, _tris in enumerate(_TABLE):
    for _i, _t in enumerate(_tris):
        _TRI_TABLE[_code, _i] = torch.tensor(_t)
_N_TRIS = torch.tensor([len(t) for t in _TABLE], dtype=torch.int64)
# per edge: offset of its lower corner inside the cube and its axis
_EDGE_BASE = torch.tensor([_CORNERS[a] for a, _ in _EDGES], dtype=torch.int64)
_EDGE_AXIS = torch.tensor([e // 4 for e in range(12)], dtype=torch.int64)


@torch.no_grad()
def marching_cubes(volume: torch.Tensor, isovalue: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """volume [X, Y, Z] -> (vertices [V, 3] float32 in index units, triangles [F, 3] int64), on volume.device."""
    assert volume.ndim == 3
    dev = volume.device
    vol = volume.float()
    X, Y, Z = vol.shape
    if min(X, Y, Z) < 2:
        return torch.zeros(0, 3, device=dev), torch.zeros(0, 3, dtype=torch.int64, device=dev)
    inside = vol < isovalue
    code = torch.zeros(X - 1, Y - 1, Z - 1, dtype=torch.int64, device=dev)
    for c, (cx, cy, cz) in enumerate(_CORNERS):
        code |= inside[cx:X - 1 + cx, cy:Y - 1 + cy, cz:Z - 1 + cz].to(torch.int64) << c
    active = torch.nonzero((code != 0) & (code != 255))                       # [Na, 3] cube origins
    if active.shape[0] == 0:
        return torch.zeros(0, 3, device=dev), torch.zeros(0, 3, dtype=torch.int64, device=dev)
    acode = code[active[:, 0], active[:, 1], active[:, 2]]
    tri_edges = _TRI_TABLE.to(dev)[acode]                                      # [Na, T, 3] local edge ids, -1 = none
    valid = tri_edges[..., 0] >= 0
    cube_of_tri = torch.nonzero(valid)[:, 0]
    tri_edges = tri_edges[valid]                                               # [F, 3]
    origin = active[cube_of_tri]                                               # [F, 3]
    base = origin[:, None, :] + _EDGE_BASE.to(dev)[tri_edges]                  # [F, 3, 3] lower grid corner of each edge
    axis = _EDGE_AXIS.to(dev)[tri_edges]                                       # [F, 3]
    # one vertex per crossed grid edge: key = (lower corner, axis)
    key = ((base[..., 0] * Y + base[..., 1]) * Z + base[..., 2]) * 3 + axis
    uniq, inverse = torch.unique(key.reshape(-1), return_inverse=True)
    faces = inverse.reshape(-1, 3)
    u_axis = uniq % 3
    lin = torch.div(uniq, 3, rounding_mode="floor")
    bz = lin % Z
    by = torch.div(lin, Z, rounding_mode="floor") % Y
    bx = torch.div(lin, Y * Z, rounding_mode="floor")
    lo = torch.stack([bx, by, bz], dim=-1)
    step = torch.nn.functional.one_hot(u_axis, 3)
    hi = lo + step
    v0 = vol[lo[:, 0], lo[:, 1], lo[:, 2]]
    v1 = vol[hi[:, 0], hi[:, 1], hi[:, 2]]
    t = ((isovalue - v0) / (v1 - v0)).clamp(0.0, 1.0)
    verts = lo.float() + step.float() * t[:, None]
    return verts, faces


class MarchingCubeHelper(torch.nn.Module):
    """reference models/geometry.py:36-113: lattice of `resolution` steps over [-1, 1) scaled by the bounds' own spacing
    (intv = 2 / resolution, as written there), evaluated block by block (block_res + 1 samples per axis so that
    neighbouring blocks share their boundary layer), marching cubes per block, blocks concatenated."""

    def __init__(self, sdf_func: Callable[[torch.Tensor], torch.Tensor], bounds, resolution: int, block_res: int = 256, method: str = "mc",
                 device=None):
        super().__init__()
        self.sdf_func, self.bounds, self.resolution = sdf_func, bounds, int(resolution)
        self.intv = 2.0 / self.resolution
        self.block_res, self.method = block_res, method
        self.device = device
        ((x_min, x_max), (y_min, y_max), (z_min, z_max)) = [(float(a), float(b)) for a, b in bounds]
        self.x_grid = torch.arange(x_min, x_max, self.intv)
        self.y_grid = torch.arange(y_min, y_max, self.intv)
        self.z_grid = torch.arange(z_min, z_max, self.intv)
        self.num_blocks_x = int(math.ceil(len(self.x_grid) / block_res))
        self.num_blocks_y = int(math.ceil(len(self.y_grid) / block_res))
        self.num_blocks_z = int(math.ceil(len(self.z_grid) / block_res))

    @torch.no_grad()
    def forward(self, threshold: float = 0.0) -> Dict[str, torch.Tensor]:
        verts_all, faces_all, n_verts = [], [], 0
        b = self.block_res
        for idx in range(self.num_blocks_x * self.num_blocks_y * self.num_blocks_z):
            bx = idx // (self.num_blocks_y * self.num_blocks_z)
            by = (idx // self.num_blocks_z) % self.num_blocks_y
            bz = idx % self.num_blocks_z
            xs, ys, zs = self.x_grid[bx * b:bx * b + b + 1], self.y_grid[by * b:by * b + b + 1], self.z_grid[bz * b:bz * b + b + 1]
            x, y, z = torch.meshgrid(xs, ys, zs, indexing="ij")
            xyz = torch.stack([x, y, z], dim=-1)
            if self.device is not None:
                xyz = xyz.to(self.device)
            level = self.sdf_func(xyz)                 # = -forward_level (models/geometry.py:85): the SDF with its sign flipped
            verts, faces = marching_cubes(-level, threshold)      # mc_func(-level, threshold), models/geometry.py:84
            if verts.shape[0] > 0:
                verts_all.append((verts * self.intv + xyz[0, 0, 0].to(verts)).cpu())
                faces_all.append((faces + n_verts).cpu())
                n_verts += verts.shape[0]
        if not verts_all:
            return {"v_pos": torch.zeros(0, 3), "t_pos_idx": torch.zeros(0, 3, dtype=torch.int64)}
        return {"v_pos": torch.cat(verts_all, dim=0), "t_pos_idx": torch.cat(faces_all, dim=0)}


def save_obj(path: str, v_pos: torch.Tensor, t_pos_idx: torch.Tensor, v_rgb: torch.Tensor = None, **_unused) -> None:
    """Wavefront OBJ with optional per-vertex colours (what trimesh writes for utils/mixins.py save_mesh)."""
    v = v_pos.detach().cpu().double()
    f = t_pos_idx.detach().cpu() + 1
    c = v_rgb.detach().cpu().double().clamp(0, 1) if v_rgb is not None else None
    with open(path, "w") as fp:
        for i in range(v.shape[0]):
            if c is None:
                fp.write("v %.8f %.8f %.8f\n" % tuple(v[i].tolist()))
            else:
                fp.write("v %.8f %.8f %.8f %.6f %.6f %.6f\n" % (*v[i].tolist(), *c[i].tolist()))
        for i in range(f.shape[0]):
            fp.write("f %d %d %d\n" % tuple(f[i].tolist()))
