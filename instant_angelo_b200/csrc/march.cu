// Occupancy-grid bitfield maintenance, ray/AABB slab test and two-pass occupancy-grid ray marching.
// Replaces nerfacc==0.3.3 OccupancyGrid._update / ray_aabb_intersect / ray_marching / render_visibility as
// called from reference models/neus.py:64-74, 108-111, 153, 159-169, 209-220.
//
// THIS FILE IS COMPILED WITH -fmad=false: outputs (t_min/t_max, packed_info, ray_indices, t_starts,
// t_ends, occupancy bits) are integer/bit-exact contracts, so every float operation is a single IEEE
// binary32 op in the order written; the only fused operations are the explicit __fmaf_rn calls (sample
// position, squared norm), matching oracle/march_ref.c.
//
// The marching kernels read a PACKED bitfield (1 bit per cell: 256 KB for 128^3, L1-resident) instead of
// nerfacc's 1-byte bools (2 MB); ia_occ_update emits both so the bool grid stays the reference-visible
// state.
#include <math.h>

#include "ia_common.cuh"

namespace {

struct GridDev {
    float lo[3], hi[3];
    int res[3];
    int type;
};

__device__ __forceinline__ float calc_dt(float t, float cone_angle, float dt_min, float dt_max)
{
    return fminf(fmaxf(t * cone_angle, dt_min), dt_max);
}

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ bool occupied_at(const float xyz[3], const GridDev &G, const uint32_t *__restrict__ bits)
{
    if (G.type == IA_AABB) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (xyz[k] < G.lo[k] || xyz[k] > G.hi[k]) return false;
    }
    float u[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) u[k] = (xyz[k] - G.lo[k]) / (G.hi[k] - G.lo[k]);
    if (G.type == IA_UN_BOUNDED_SPHERE) {
        float v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = u[k] * 2.0f - 1.0f;
        const float nsq = __fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], v[0] * v[0]));
        const float nrm = sqrtf(nsq);
        if (nrm > 1.0f) {
            const float s = 2.0f - 1.0f / nrm;
#pragma unroll
            for (int k = 0; k < 3; ++k) v[k] = s * (v[k] / nrm);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) u[k] = v[k] * 0.25f + 0.5f;
    }
    if (bits == nullptr) return true;
    const int ix = min(max((int)(u[0] * (float)G.res[0]), 0), G.res[0] - 1);
    const int iy = min(max((int)(u[1] * (float)G.res[1]), 0), G.res[1] - 1);
    const int iz = min(max((int)(u[2] * (float)G.res[2]), 0), G.res[2] - 1);
    const int64_t idx = (int64_t)ix * G.res[1] * G.res[2] + (int64_t)iy * G.res[2] + iz;
    return (__ldg(bits + (idx >> 5)) >> (idx & 31)) & 1u;
}

__device__ __forceinline__ float distance_to_next_voxel(const float xyz[3], const float d[3], const float inv_d[3],
                                                        const GridDev &G)
{
    float best[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float r = (float)G.res[k];
        const float p = ((xyz[k] - G.lo[k]) / (G.hi[k] - G.lo[k])) * r;
        const float target = floorf(p + 0.5f + 0.5f * sgnf(d[k]));
        best[k] = ((target - p) * inv_d[k]) / r * (G.hi[k] - G.lo[k]);
    }
    const float t = fminf(fminf(best[0], best[1]), best[2]);
    return fmaxf(t, 0.0f);
}

// One thread per ray.  WRITE=false: count pass.  WRITE=true: emit samples at packed_info[r].offset.
template <bool WRITE>
__device__ __forceinline__ void march_ray(int64_t i, const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                          const float *__restrict__ t_min, const float *__restrict__ t_max, const GridDev &G,
                                          const uint32_t *__restrict__ bits, float step_size, float cone_angle,
                                          const int32_t *__restrict__ packed_info, int32_t *__restrict__ num_steps,
                                          int32_t *__restrict__ ray_indices, float *__restrict__ t_starts, float *__restrict__ t_ends)
{
    const float o[3] = {rays_o[3 * i], rays_o[3 * i + 1], rays_o[3 * i + 2]};
    const float d[3] = {rays_d[3 * i], rays_d[3 * i + 1], rays_d[3 * i + 2]};
    const float inv_d[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
    const float near = t_min[i], far = t_max[i];
    const float dt_min = step_size, dt_max = 1e10f;
    int64_t base = 0;
    if (WRITE) base = packed_info[2 * i];

    int j = 0;
    float t0 = near;
    float dt = calc_dt(t0, cone_angle, dt_min, dt_max);
    float t1 = t0 + dt;
    float t_mid = (t0 + t1) * 0.5f;
    while (t_mid < far) {
        float xyz[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) xyz[k] = __fmaf_rn(t_mid, d[k], o[k]);
        if (occupied_at(xyz, G, bits)) {
            if (WRITE) {
                t_starts[base + j] = t0;
                t_ends[base + j] = t1;
                ray_indices[base + j] = (int32_t)i;
            }
            ++j;
            t0 = t1;
            t1 = t0 + calc_dt(t0, cone_angle, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        } else if (G.type == IA_AABB) {
            float t_target = t_mid + distance_to_next_voxel(xyz, d, inv_d, G);
            t_target = fminf(t_target, far);
            do {
                t_mid += dt_min;
            } while (t_mid < t_target);
            dt = calc_dt(t_mid, cone_angle, dt_min, dt_max);
            t0 = t_mid - dt * 0.5f;
            t1 = t_mid + dt * 0.5f;
        } else {
            t0 = t1;
            t1 = t0 + calc_dt(t0, cone_angle, dt_min, dt_max);
            t_mid = (t0 + t1) * 0.5f;
        }
    }
    if (!WRITE) num_steps[i] = j;
}

template <bool WRITE>
__global__ void __launch_bounds__(64)
march_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, const float *__restrict__ t_min,
             const float *__restrict__ t_max, int64_t n_rays, const GridDev G, const uint32_t *__restrict__ bits,
             float step_size, float cone_angle, const int32_t *__restrict__ packed_info,
             int32_t *__restrict__ num_steps, int32_t *__restrict__ ray_indices, float *__restrict__ t_starts,
             float *__restrict__ t_ends)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    march_ray<WRITE>(i, rays_o, rays_d, t_min, t_max, G, bits, step_size, cone_angle, packed_info, num_steps, ray_indices, t_starts, t_ends);
}

// Two marches of the same rays in one launch (blockIdx.y selects the set): the foreground march through the AABB grid and the
// background march through the contracted grid (reference models/neus.py:209-220 and :159-169) are independent, and one
// thread per ray leaves either of them with 128 CTAs of 64 threads on 148 SMs -- latency bound, half the machine idle.
struct MarchSet {
    const float *t_min, *t_max;
    GridDev G;
    const uint32_t *bits;
    float step_size, cone_angle;
    const int32_t *packed_info;
    int32_t *num_steps, *ray_indices;
    float *t_starts, *t_ends;
};

template <bool WRITE>
__global__ void __launch_bounds__(64)
march_pair_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t n_rays, const MarchSet A, const MarchSet B)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    if (blockIdx.y == 0)
        march_ray<WRITE>(i, rays_o, rays_d, A.t_min, A.t_max, A.G, A.bits, A.step_size, A.cone_angle, A.packed_info, A.num_steps,
                         A.ray_indices, A.t_starts, A.t_ends);
    else
        march_ray<WRITE>(i, rays_o, rays_d, B.t_min, B.t_max, B.G, B.bits, B.step_size, B.cone_angle, B.packed_info, B.num_steps,
                         B.ray_indices, B.t_starts, B.t_ends);
}

__global__ void __launch_bounds__(256)
aabb_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d, int64_t n_rays, float x0, float y0,
            float z0, float x1, float y1, float z1, int clamp_zero, float *__restrict__ t_min,
            float *__restrict__ t_max)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const float ox = rays_o[3 * i], oy = rays_o[3 * i + 1], oz = rays_o[3 * i + 2];
    const float dx = rays_d[3 * i], dy = rays_d[3 * i + 1], dz = rays_d[3 * i + 2];
    float tmin = (x0 - ox) / dx, tmax = (x1 - ox) / dx;
    if (tmin > tmax) { float t = tmin; tmin = tmax; tmax = t; }
    float tymin = (y0 - oy) / dy, tymax = (y1 - oy) / dy;
    if (tymin > tymax) { float t = tymin; tymin = tymax; tymax = t; }
    if (tmin > tymax || tymin > tmax) { t_min[i] = 1e10f; t_max[i] = 1e10f; return; }
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    float tzmin = (z0 - oz) / dz, tzmax = (z1 - oz) / dz;
    if (tzmin > tzmax) { float t = tzmin; tzmin = tzmax; tzmax = t; }
    if (tmin > tzmax || tzmin > tmax) { t_min[i] = 1e10f; t_max[i] = 1e10f; return; }
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    if (clamp_zero && !(tmin > 0.0f)) tmin = 0.0f;
    t_min[i] = tmin;
    t_max[i] = tmax;
}

// Single-CTA exclusive scan with a carried prefix (n_rays is 10^3..10^6; latency ~ n/4096 block scans).
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const int32_t *__restrict__ num_steps, int64_t n, int32_t *__restrict__ packed_info,
            int64_t *__restrict__ total)
{
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t chunk = 0; chunk < n; chunk += (int64_t)SCAN_THREADS * SCAN_ITEMS) {
        const int64_t first = chunk + (int64_t)tid * SCAN_ITEMS;
        int32_t v[SCAN_ITEMS];
        int64_t local = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (first + k < n) ? num_steps[first + k] : 0;
            local += v[k];
        }
        int64_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sums[lane];
            int64_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;  // exclusive prefix of the warp totals
        }
        __syncthreads();
        const int64_t carry = carry_s;
        int64_t run = carry + warp_sums[warp] + (incl - local);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (first + k < n) {
                packed_info[2 * (first + k)] = (int32_t)run;
                packed_info[2 * (first + k) + 1] = v[k];
            }
            run += v[k];
        }
        __syncthreads();
        if (tid == SCAN_THREADS - 1) carry_s = run;
        __syncthreads();
    }
    if (tid == 0) *total = carry_s;
}

__global__ void __launch_bounds__(128)
visibility_kernel(const float *__restrict__ alphas, const int32_t *__restrict__ packed_info, int64_t n_rays,
                  float early_stop_eps, float alpha_thre, uint8_t *__restrict__ vis)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const int64_t base = packed_info[2 * i];
    const int cnt = packed_info[2 * i + 1];
    float T = 1.0f;
    for (int j = 0; j < cnt; ++j) {
        const float a = alphas[base + j];
        bool v = T >= early_stop_eps;
        if (alpha_thre > 0.0f) v = v && (a >= alpha_thre);
        vis[base + j] = v ? 1 : 0;
        T = T * (1.0f - a);
    }
}

// ---- visibility pruning of marched samples with their compaction (nerfacc.ray_marching with sigma_fn: alphas from the
// densities, render_visibility, boolean indexing of the three sample arrays and a new packed_info) as count -> scan -> write,
// the marcher's own structure, instead of ~22 tensor-operator launches.  alpha = 1 - exp(-sigma (t1 - t0)) is rounded
// operation by operation like the tensor expression; T and the test are those of visibility_kernel.
__global__ void __launch_bounds__(128)
prune_count_kernel(const float *__restrict__ sigmas, const float *__restrict__ t0, const float *__restrict__ t1,
                   const int32_t *__restrict__ packed_info, int64_t n_rays, float early_stop_eps, float alpha_thre,
                   uint8_t *__restrict__ vis, int32_t *__restrict__ num_kept)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const int64_t base = packed_info[2 * i];
    const int cnt = packed_info[2 * i + 1];
    float T = 1.0f;
    int kept = 0;
    for (int j = 0; j < cnt; ++j) {
        const float a = __fsub_rn(1.0f, expf(__fmul_rn(-sigmas[base + j], __fsub_rn(t1[base + j], t0[base + j]))));
        bool v = T >= early_stop_eps;
        if (alpha_thre > 0.0f) v = v && (a >= alpha_thre);
        vis[base + j] = v ? 1 : 0;
        kept += v ? 1 : 0;
        T = T * (1.0f - a);
    }
    num_kept[i] = kept;
}

__global__ void __launch_bounds__(128)
prune_write_kernel(const uint8_t *__restrict__ vis, const int32_t *__restrict__ packed_old, const int32_t *__restrict__ packed_new,
                   const float *__restrict__ t0, const float *__restrict__ t1, int64_t n_rays, int32_t *__restrict__ out_ri,
                   float *__restrict__ out_t0, float *__restrict__ out_t1)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const int64_t base = packed_old[2 * i];
    const int cnt = packed_old[2 * i + 1];
    int64_t dst = packed_new[2 * i];
    for (int j = 0; j < cnt; ++j) {
        if (vis[base + j]) {
            out_ri[dst] = (int32_t)i;
            out_t0[dst] = t0[base + j];
            out_t1[dst] = t1[base + j];
            ++dst;
        }
    }
}

// ---- occupancy grid ---------------------------------------------------------------------------------

__global__ void fill_kernel(float *p, int64_t n, float v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void occ_scatter_max_kernel(const int64_t *__restrict__ idx, const float *__restrict__ occ, int64_t n,
                                       float *__restrict__ tmp)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = fmaxf(occ[i], 0.0f);  // int ordering == float ordering for v >= 0; sentinel is -1
    atomicMax(reinterpret_cast<int *>(tmp) + idx[i], __float_as_int(v));
}

// occs <- max(occs*decay, new) on touched cells; accumulates sum(occs) in double.
__global__ void __launch_bounds__(256)
occ_ema_kernel(float *__restrict__ occs, int64_t num_cells, const float *__restrict__ tmp,
               const float *__restrict__ occ_dense, float decay, double *__restrict__ sum)
{
    __shared__ double part[8];
    double acc = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < num_cells; c += (int64_t)gridDim.x * blockDim.x) {
        float o = occs[c];
        if (occ_dense != nullptr) {
            o = fmaxf(o * decay, fmaxf(occ_dense[c], 0.0f));
            occs[c] = o;
        } else {
            const float t = tmp[c];
            if (t >= 0.0f) {
                o = fmaxf(o * decay, t);
                occs[c] = o;
            }
        }
        acc += (double)o;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += part[w];
        atomicAdd(sum, s);
    }
}

__global__ void __launch_bounds__(256)
occ_threshold_kernel(const float *__restrict__ occs, int64_t num_cells, const double *__restrict__ sum, float occ_thre,
                     uint8_t *__restrict__ binary, uint32_t *__restrict__ bitfield)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float mean = (float)(*sum / (double)num_cells);
    const float thr = fminf(mean, occ_thre);
    const bool on = c < num_cells && occs[c] > thr;
    if (c < num_cells && binary != nullptr) binary[c] = on ? 1 : 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && c < num_cells && bitfield != nullptr) bitfield[c >> 5] = word;
}

__global__ void __launch_bounds__(256)
occ_pack_kernel(const uint8_t *__restrict__ binary, int64_t num_cells, uint32_t *__restrict__ bitfield)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = c < num_cells && binary[c] != 0;
    const uint32_t word = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31) == 0 && c < num_cells) bitfield[c >> 5] = word;
}

int to_grid(const ia_grid_desc *g, GridDev *G)
{
    IA_REQUIRE(g != nullptr, "march: grid descriptor is NULL");
    IA_REQUIRE(g->contraction == IA_AABB || g->contraction == IA_UN_BOUNDED_SPHERE,
               "march: contraction type %d unsupported", g->contraction);
    for (int k = 0; k < 3; ++k) {
        G->lo[k] = g->roi[k];
        G->hi[k] = g->roi[3 + k];
        G->res[k] = g->res[k];
        IA_REQUIRE(g->res[k] >= 1, "march: grid resolution must be >= 1");
    }
    G->type = g->contraction;
    return IA_OK;
}

}  // namespace

extern "C" int32_t ia_aabb(const float *rays_o, const float *rays_d, int64_t n_rays, const float *aabb, int32_t clamp_zero,
                           float *t_min, float *t_max, void *stream)
{
    IA_REQUIRE(aabb != nullptr, "aabb: aabb_host6 is NULL");
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (rays_o && rays_d && t_min && t_max)), "aabb: NULL pointer");
    if (n_rays == 0) return IA_OK;
    aabb_kernel<<<(unsigned)ia_ceil_div(n_rays, 256), 256, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, n_rays, aabb[0], aabb[1], aabb[2], aabb[3], aabb[4], aabb[5], clamp_zero, t_min, t_max);
    IA_LAUNCH_OK("aabb_kernel");
    return IA_OK;
}

extern "C" int32_t ia_march_count(const float *rays_o, const float *rays_d, const float *t_min, const float *t_max,
                                  int64_t n_rays, const ia_grid_desc *grid, const uint32_t *bitfield, float step_size,
                                  float cone_angle, int32_t *num_steps, void *stream)
{
    GridDev G;
    int rc = to_grid(grid, &G);
    if (rc) return rc;
    IA_REQUIRE(step_size > 0.f, "march: render_step_size must be > 0");
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (rays_o && rays_d && t_min && t_max && num_steps)), "march_count: NULL pointer");
    if (n_rays == 0) return IA_OK;
    march_kernel<false><<<(unsigned)ia_ceil_div(n_rays, 64), 64, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, t_min, t_max, n_rays, G, bitfield, step_size, cone_angle, nullptr, num_steps, nullptr, nullptr,
        nullptr);
    IA_LAUNCH_OK("march_kernel<count>");
    return IA_OK;
}

extern "C" int64_t ia_march_scan_workspace_bytes(int64_t n_rays) { (void)n_rays; return 16; }

extern "C" int32_t ia_march_scan(const int32_t *num_steps, int64_t n_rays, int32_t *packed_info, int64_t *total_dev,
                                 void *workspace, void *stream)
{
    (void)workspace;
    IA_REQUIRE(total_dev != nullptr, "march_scan: total_dev is NULL");
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (num_steps && packed_info)), "march_scan: NULL pointer");
    scan_kernel<<<1, SCAN_THREADS, 0, (cudaStream_t)stream>>>(num_steps, n_rays, packed_info, total_dev);
    IA_LAUNCH_OK("scan_kernel");
    return IA_OK;
}

extern "C" int32_t ia_march_total(const int64_t *total_dev, int64_t *total_host, void *stream)
{
    IA_REQUIRE(total_dev && total_host, "march_total: NULL pointer");
    IA_CUDA_OK(cudaMemcpyAsync(total_host, total_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    IA_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return IA_OK;
}

extern "C" int32_t ia_march_write(const float *rays_o, const float *rays_d, const float *t_min, const float *t_max,
                                  int64_t n_rays, const ia_grid_desc *grid, const uint32_t *bitfield, float step_size,
                                  float cone_angle, const int32_t *packed_info, int32_t *ray_indices, float *t_starts,
                                  float *t_ends, void *stream)
{
    GridDev G;
    int rc = to_grid(grid, &G);
    if (rc) return rc;
    IA_REQUIRE(step_size > 0.f, "march: render_step_size must be > 0");
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (rays_o && rays_d && t_min && t_max && packed_info)), "march_write: NULL pointer");
    if (n_rays == 0) return IA_OK;
    march_kernel<true><<<(unsigned)ia_ceil_div(n_rays, 64), 64, 0, (cudaStream_t)stream>>>(
        rays_o, rays_d, t_min, t_max, n_rays, G, bitfield, step_size, cone_angle, packed_info, nullptr, ray_indices,
        t_starts, t_ends);
    IA_LAUNCH_OK("march_kernel<write>");
    return IA_OK;
}

namespace {
int to_set(const ia_march_set *h, bool write, MarchSet *S, const char *which)
{
    IA_REQUIRE(h != nullptr, "march_pair: set %s is NULL", which);
    int rc = to_grid(h->grid, &S->G);
    if (rc) return rc;
    IA_REQUIRE(h->step_size > 0.f, "march_pair: render_step_size of set %s must be > 0", which);
    IA_REQUIRE(h->t_min && h->t_max, "march_pair: t_min / t_max of set %s is NULL", which);
    IA_REQUIRE(write ? (h->packed_info != nullptr) : (h->num_steps != nullptr), "march_pair: set %s lacks the %s-pass buffers", which,
               write ? "write" : "count");
    S->t_min = h->t_min; S->t_max = h->t_max; S->bits = h->bitfield;
    S->step_size = h->step_size; S->cone_angle = h->cone_angle;
    S->packed_info = h->packed_info; S->num_steps = h->num_steps;
    S->ray_indices = h->ray_indices; S->t_starts = h->t_starts; S->t_ends = h->t_ends;
    return IA_OK;
}
}  // namespace

extern "C" int32_t ia_march_pair(const float *rays_o, const float *rays_d, int64_t n_rays, const ia_march_set *a, const ia_march_set *b,
                                 int32_t write, void *stream)
{
    MarchSet A, B;
    int rc = to_set(a, write != 0, &A, "a");
    if (rc) return rc;
    rc = to_set(b, write != 0, &B, "b");
    if (rc) return rc;
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (rays_o && rays_d)), "march_pair: NULL rays");
    if (n_rays == 0) return IA_OK;
    const dim3 grid((unsigned)ia_ceil_div(n_rays, 64), 2);
    if (write) march_pair_kernel<true><<<grid, 64, 0, (cudaStream_t)stream>>>(rays_o, rays_d, n_rays, A, B);
    else march_pair_kernel<false><<<grid, 64, 0, (cudaStream_t)stream>>>(rays_o, rays_d, n_rays, A, B);
    IA_LAUNCH_OK("march_pair_kernel");
    return IA_OK;
}

extern "C" int32_t ia_march_totals(const int64_t *totals_dev, int32_t count, int64_t *totals_host, void *stream)
{
    IA_REQUIRE(totals_dev && totals_host && count >= 1 && count <= 16, "march_totals: bad arguments");
    IA_CUDA_OK(cudaMemcpyAsync(totals_host, totals_dev, sizeof(int64_t) * (size_t)count, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    IA_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    return IA_OK;
}

extern "C" int32_t ia_visibility(const float *alphas, const int32_t *packed_info, int64_t n_rays, float early_stop_eps,
                                 float alpha_thre, uint8_t *visible, void *stream)
{
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || packed_info), "visibility: NULL pointer");
    if (n_rays == 0) return IA_OK;
    visibility_kernel<<<(unsigned)ia_ceil_div(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        alphas, packed_info, n_rays, early_stop_eps, alpha_thre, visible);
    IA_LAUNCH_OK("visibility_kernel");
    return IA_OK;
}

extern "C" int32_t ia_prune_count(const float *sigmas, const float *t_starts, const float *t_ends, const int32_t *packed_info,
                                  int64_t n_rays, float early_stop_eps, float alpha_thre, uint8_t *visible, int32_t *num_kept,
                                  void *stream)
{
    // visible / sigmas / t_* are NULL when no ray has a sample (empty tensors): every count in packed_info is then zero
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (packed_info && num_kept)), "prune_count: NULL pointer");
    if (n_rays == 0) return IA_OK;
    prune_count_kernel<<<(unsigned)ia_ceil_div(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        sigmas, t_starts, t_ends, packed_info, n_rays, early_stop_eps, alpha_thre, visible, num_kept);
    IA_LAUNCH_OK("prune_count_kernel");
    return IA_OK;
}

extern "C" int32_t ia_prune_write(const uint8_t *visible, const int32_t *packed_info, const int32_t *packed_info_kept,
                                  const float *t_starts, const float *t_ends, int64_t n_rays, int32_t *ray_indices_kept,
                                  float *t_starts_kept, float *t_ends_kept, void *stream)
{
    IA_REQUIRE(n_rays >= 0 && (n_rays == 0 || (visible && packed_info && packed_info_kept)), "prune_write: NULL pointer");
    if (n_rays == 0) return IA_OK;
    prune_write_kernel<<<(unsigned)ia_ceil_div(n_rays, 128), 128, 0, (cudaStream_t)stream>>>(
        visible, packed_info, packed_info_kept, t_starts, t_ends, n_rays, ray_indices_kept, t_starts_kept, t_ends_kept);
    IA_LAUNCH_OK("prune_write_kernel");
    return IA_OK;
}

extern "C" int64_t ia_occ_workspace_bytes(int64_t num_cells) { return 256 + num_cells * (int64_t)sizeof(float); }

extern "C" int32_t ia_occ_update(const int64_t *idx, const float *occ, int64_t n, float *occs, int64_t num_cells,
                                 float ema_decay, float occ_thre, uint8_t *binary, uint32_t *bitfield, void *workspace,
                                 void *stream)
{
    IA_REQUIRE(occs && workspace && num_cells > 0, "occ_update: NULL pointer");
    IA_REQUIRE(num_cells % 32 == 0, "occ_update: num_cells must be a multiple of 32 (got %lld)", (long long)num_cells);
    IA_REQUIRE(idx != nullptr || n == num_cells, "occ_update: idx == NULL requires n == num_cells");
    cudaStream_t s = (cudaStream_t)stream;
    double *sum = reinterpret_cast<double *>(workspace);
    float *tmp = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + 256);
    IA_CUDA_OK(cudaMemsetAsync(sum, 0, sizeof(double), s));
    const unsigned cell_blocks = (unsigned)ia_ceil_div(num_cells, 256);
    const unsigned ema_blocks = (unsigned)min((int64_t)cell_blocks, (int64_t)ia_sm_count() * 8);
    if (idx != nullptr) {
        fill_kernel<<<cell_blocks, 256, 0, s>>>(tmp, num_cells, -1.0f);
        IA_LAUNCH_OK("fill_kernel");
        if (n > 0) {
            occ_scatter_max_kernel<<<(unsigned)ia_ceil_div(n, 256), 256, 0, s>>>(idx, occ, n, tmp);
            IA_LAUNCH_OK("occ_scatter_max_kernel");
        }
        occ_ema_kernel<<<ema_blocks, 256, 0, s>>>(occs, num_cells, tmp, nullptr, ema_decay, sum);
    } else {
        occ_ema_kernel<<<ema_blocks, 256, 0, s>>>(occs, num_cells, nullptr, occ, ema_decay, sum);
    }
    IA_LAUNCH_OK("occ_ema_kernel");
    occ_threshold_kernel<<<cell_blocks, 256, 0, s>>>(occs, num_cells, sum, occ_thre, binary, bitfield);
    IA_LAUNCH_OK("occ_threshold_kernel");
    return IA_OK;
}

extern "C" int32_t ia_occ_pack(const uint8_t *binary, int64_t num_cells, uint32_t *bitfield, void *stream)
{
    IA_REQUIRE(binary && bitfield && num_cells > 0 && num_cells % 32 == 0, "occ_pack: bad arguments");
    occ_pack_kernel<<<(unsigned)ia_ceil_div(num_cells, 256), 256, 0, (cudaStream_t)stream>>>(binary, num_cells, bitfield);
    IA_LAUNCH_OK("occ_pack_kernel");
    return IA_OK;
}
