// Device-side geometry of the multiresolution hash grid (tcnn GridEncoding semantics), shared by
// the stand-alone encoder kernels (hashgrid.cu) and the fused encoder + MLP kernels (mlp_tc.cu).
#pragma once
#include "ia_common.cuh"

namespace {

struct GridParams {
    int32_t n_levels;
    int32_t active;
    int32_t n_dense;   // leading levels with a dense (un-hashed) index when every later level is hashed; -1 otherwise
    float scale[IA_MAX_LEVELS];
    uint32_t res[IA_MAX_LEVELS];
    uint32_t size[IA_MAX_LEVELS];
    uint32_t offset[IA_MAX_LEVELS];
    uint32_t hashed[IA_MAX_LEVELS];
};

__device__ __forceinline__ uint32_t entry_index(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t res, uint32_t size,
                                                bool hashed)
{
    // tcnn grid_index(): dense stride index, or coherent prime hash when the dense grid would overflow
    // the level.  Hashed levels have a power-of-two size (2^log2_hashmap_size) => mask; dense levels can
    // exceed `size` only on the upper boundary (corner == res), by less than one `size`.
    if (hashed) return (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) & (size - 1u);
    // For cells inside the grid (all callers check cell_outside() first): idx <= res + res^2 + res^3 < 2 * size (size = res^3
    // rounded up to a multiple of 8), so tcnn's `idx % size` is one conditional subtraction -- no integer division on the
    // gather path
    uint32_t idx = cx + cy * res + cz * res * res;
    return idx >= size ? idx - size : idx;
}

// tcnn's index for ANY cell: points outside the unit cube (e.g. un-clamped COLMAP points handed to VolumeSDF by the
// sparse-point losses) give negative / huge cell coordinates whose uint32 stride sum wraps anywhere; `% size` keeps the
// index inside the level exactly as tcnn's does.  Never on the hot path.
__device__ __noinline__ uint32_t entry_index_total(uint32_t cx, uint32_t cy, uint32_t cz, uint32_t res, uint32_t size, bool hashed)
{
    if (hashed) return (cx ^ (cy * 2654435761u) ^ (cz * 805459861u)) & (size - 1u);
    return (cx + cy * res + cz * res * res) % size;
}

// A point inside the unit cube has cells inside the grid at every level (0 <= ix <= res - 1), which is what the conditional
// subtraction of entry_index() needs.  Everything else takes the total index.  One test per POINT, outside the level loops:
// the hot loops carry no extra instruction and the total path is a separate, cold instance of the loop body.
__device__ __forceinline__ bool point_outside(float px, float py, float pz)
{
    return !(px >= 0.f && px <= 1.f && py >= 0.f && py <= 1.f && pz >= 0.f && pz <= 1.f);
}

// the four (y, z) corner indices of one x side of a cell; TOTAL: the cell may lie outside the grid
template <bool TOTAL>
__device__ __forceinline__ void corner4(uint32_t cx, uint32_t iy, uint32_t iz, uint32_t res, uint32_t size, bool hashed,
                                        uint32_t &i00, uint32_t &i10, uint32_t &i01, uint32_t &i11)
{
    if (!TOTAL) {
        i00 = entry_index(cx, iy, iz, res, size, hashed);
        i10 = entry_index(cx, iy + 1, iz, res, size, hashed);
        i01 = entry_index(cx, iy, iz + 1, res, size, hashed);
        i11 = entry_index(cx, iy + 1, iz + 1, res, size, hashed);
    } else {
        i00 = entry_index_total(cx, iy, iz, res, size, hashed);
        i10 = entry_index_total(cx, iy + 1, iz, res, size, hashed);
        i01 = entry_index_total(cx, iy, iz + 1, res, size, hashed);
        i11 = entry_index_total(cx, iy + 1, iz + 1, res, size, hashed);
    }
}

struct CellCoords {
    uint32_t ix, iy, iz;
    float wx, wy, wz;
};

__device__ __forceinline__ CellCoords locate(float px, float py, float pz, float scale)
{
    CellCoords c;
    float fx = fmaf(scale, px, 0.5f), fy = fmaf(scale, py, 0.5f), fz = fmaf(scale, pz, 0.5f);
    float gx = floorf(fx), gy = floorf(fy), gz = floorf(fz);
    c.wx = fx - gx;
    c.wy = fy - gy;
    c.wz = fz - gz;
    c.ix = (uint32_t)(int)gx;
    c.iy = (uint32_t)(int)gy;
    c.iz = (uint32_t)(int)gz;
    return c;
}

// ---- one thread, one (point, level): all 8 corners.  Used by the fused encoder + MLP kernels (mlp_tc.cu), where thread
// (row, column group) fills its own 16-byte slice (4 levels x 2 features) of the row's tensor-core operand.  Compared with the
// lane-pair mapping of the stand-alone encoder (hashgrid.cu) the cell is located once instead of twice and the eight values
// are combined by seven lerps per feature.  `l` is warp-uniform at every call site (the level set depends on the warp's
// column group only), so the dense / hashed / outside-the-grid branches do not diverge inside the unit cube.
template <bool TOTAL>
__device__ __forceinline__ void hg_corner8(const GridParams &P, int l, const CellCoords &c, uint32_t (&i)[8])
{
    const uint32_t res = P.res[l], size = P.size[l];
    if (P.hashed[l]) {
        const uint32_t m = size - 1u;
        const uint32_t hy0 = c.iy * 2654435761u, hy1 = hy0 + 2654435761u;
        const uint32_t hz0 = c.iz * 805459861u, hz1 = hz0 + 805459861u;
        const uint32_t x0 = c.ix, x1 = c.ix + 1u;
        i[0] = (x0 ^ hy0 ^ hz0) & m; i[1] = (x1 ^ hy0 ^ hz0) & m;
        i[2] = (x0 ^ hy1 ^ hz0) & m; i[3] = (x1 ^ hy1 ^ hz0) & m;
        i[4] = (x0 ^ hy0 ^ hz1) & m; i[5] = (x1 ^ hy0 ^ hz1) & m;
        i[6] = (x0 ^ hy1 ^ hz1) & m; i[7] = (x1 ^ hy1 ^ hz1) & m;
    } else if (!TOTAL) {
        const uint32_t r2 = res * res;
        const uint32_t b = c.ix + c.iy * res + c.iz * r2;
        i[0] = b; i[1] = b + 1u; i[2] = b + res; i[3] = b + res + 1u;
        i[4] = b + r2; i[5] = b + r2 + 1u; i[6] = b + r2 + res; i[7] = b + r2 + res + 1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) i[k] = i[k] >= size ? i[k] - size : i[k];
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            i[k] = entry_index_total(c.ix + (k & 1), c.iy + ((k >> 1) & 1), c.iz + (k >> 2), res, size, false);
    }
}

// features (f0, f1) of level l at p = (px, py, pz); TOTAL = false requires p in [0,1]^3 (point_outside())
template <bool TOTAL>
__device__ __forceinline__ void hg_gather_level(const GridParams &P, const float2 *__restrict__ table, int l, float px, float py,
                                                float pz, float &f0, float &f1)
{
    const CellCoords c = locate(px, py, pz, P.scale[l]);
    uint32_t i[8];
    hg_corner8<TOTAL>(P, l, c, i);
    const float2 *__restrict__ tl = table + P.offset[l];
    float2 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __ldg(tl + i[k]);
    const float x00a = fmaf(c.wx, v[1].x - v[0].x, v[0].x), x00b = fmaf(c.wx, v[1].y - v[0].y, v[0].y);
    const float x10a = fmaf(c.wx, v[3].x - v[2].x, v[2].x), x10b = fmaf(c.wx, v[3].y - v[2].y, v[2].y);
    const float x01a = fmaf(c.wx, v[5].x - v[4].x, v[4].x), x01b = fmaf(c.wx, v[5].y - v[4].y, v[4].y);
    const float x11a = fmaf(c.wx, v[7].x - v[6].x, v[6].x), x11b = fmaf(c.wx, v[7].y - v[6].y, v[6].y);
    const float y0a = fmaf(c.wy, x10a - x00a, x00a), y0b = fmaf(c.wy, x10b - x00b, x00b);
    const float y1a = fmaf(c.wy, x11a - x01a, x01a), y1b = fmaf(c.wy, x11b - x01b, x01b);
    f0 = fmaf(c.wz, y1a - y0a, y0a);
    f1 = fmaf(c.wz, y1b - y0b, y0b);
}

inline int fill_params(const ia_grid_plan *plan, int32_t active_levels, GridParams *P)
{
    IA_REQUIRE(plan != nullptr, "hashgrid: plan is NULL");
    IA_REQUIRE(plan->n_features == 2, "hashgrid: only n_features_per_level == 2 is supported (got %d)", plan->n_features);
    IA_REQUIRE(plan->n_levels >= 1 && plan->n_levels <= 16, "hashgrid: n_levels must be in [1,16] (got %d)", plan->n_levels);
    IA_REQUIRE(active_levels >= 0 && active_levels <= plan->n_levels, "hashgrid: active_levels %d out of range", active_levels);
    P->n_levels = plan->n_levels;
    P->active = active_levels;
    for (int l = 0; l < plan->n_levels; ++l) {
        P->scale[l] = plan->scale[l];
        P->res[l] = plan->res[l];
        P->size[l] = plan->size[l];
        P->offset[l] = plan->offset[l];
        P->hashed[l] = plan->hashed[l];
        IA_REQUIRE(!plan->hashed[l] || (plan->size[l] & (plan->size[l] - 1)) == 0,
                   "hashgrid: hashed level %d has a non power-of-two size %u", l, plan->size[l]);
    }
    // tcnn's levels grow monotonically, so the dense ones come first; any other plan keeps the generic kernels
    int nd = 0;
    while (nd < plan->n_levels && !plan->hashed[nd]) ++nd;
    P->n_dense = nd;
    for (int l = nd; l < plan->n_levels; ++l)
        if (!plan->hashed[l]) P->n_dense = -1;
    return IA_OK;
}


}  // namespace
