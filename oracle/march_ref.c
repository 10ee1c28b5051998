/* ORACLE (test infrastructure only) -- plain C restatement of nerfacc==0.3.3's ray/AABB slab test and
 * occupancy-grid ray marching, the third-party kernels Instant-angelo calls at
 *   models/neus.py:153          (ray_aabb_intersect)
 *   models/neus.py:159-169      (background ray_marching, UN_BOUNDED_SPHERE grid, cone stepping)
 *   models/neus.py:209-220      (foreground ray_marching, AABB grid, fixed step, DDA-style skipping)
 * nerfacc 0.3.3 (requirements.txt:3) is not vendored under /root/reference and not installable here, so
 * this follows its published algorithm.  PARITY UNPINNED: the reference
 * holds no tests/golden vectors for this boundary; the known-answer tests live in tests/.
 *
 * Floating-point contract (what "bit-exact" means for the CUDA path): every operation is an IEEE
 * binary32 operation in the order written here; the only fused multiply-add is the sample position
 * xyz = fmaf(t_mid, d, o) (what nvcc emits for nerfacc's `origin + t_mid * dir`) and the squared norm
 * of the sphere contraction.  Build with -ffp-contract=off so the compiler adds no others.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 */
#include <math.h>
#include <stdint.h>

#define IA_AABB 0
#define IA_UN_BOUNDED_SPHERE 2 /* nerfacc ContractionType enum: AABB=0, UN_BOUNDED_TANH=1, UN_BOUNDED_SPHERE=2 */

static inline void swapf(float *a, float *b) { float t = *a; *a = *b; *b = t; }

/* nerfacc/cuda/csrc/intersection.cu: slab test; miss => (1e10, 1e10).  clamp_zero!=0 applies the
 * kernel wrapper's `t_min = max(t_min, 0)` (see DESIGN.md "t_min clamp"). */
void ia_ref_aabb(const float *o, const float *d, int64_t n_rays, const float *aabb, int clamp_zero,
                 float *t_min, float *t_max)
{
    for (int64_t i = 0; i < n_rays; ++i) {
        const float *ro = o + 3 * i, *rd = d + 3 * i;
        float tmin = (aabb[0] - ro[0]) / rd[0];
        float tmax = (aabb[3] - ro[0]) / rd[0];
        if (tmin > tmax) swapf(&tmin, &tmax);
        float tymin = (aabb[1] - ro[1]) / rd[1];
        float tymax = (aabb[4] - ro[1]) / rd[1];
        if (tymin > tymax) swapf(&tymin, &tymax);
        if (tmin > tymax || tymin > tmax) { t_min[i] = 1e10f; t_max[i] = 1e10f; continue; }
        if (tymin > tmin) tmin = tymin;
        if (tymax < tmax) tmax = tymax;
        float tzmin = (aabb[2] - ro[2]) / rd[2];
        float tzmax = (aabb[5] - ro[2]) / rd[2];
        if (tzmin > tzmax) swapf(&tzmin, &tzmax);
        if (tmin > tzmax || tzmin > tmax) { t_min[i] = 1e10f; t_max[i] = 1e10f; continue; }
        if (tzmin > tmin) tmin = tzmin;
        if (tzmax < tmax) tmax = tzmax;
        if (clamp_zero && !(tmin > 0.0f)) tmin = 0.0f;
        t_min[i] = tmin;
        t_max[i] = tmax;
    }
}

static inline float calc_dt(float t, float cone_angle, float dt_min, float dt_max)
{
    float v = t * cone_angle;
    return fminf(fmaxf(v, dt_min), dt_max);
}

static inline float sgn(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* world -> unit cube of the grid (nerfacc contraction.cuh) */
static inline void to_unit(const float xyz[3], const float *roi, int type, float u[3])
{
    for (int k = 0; k < 3; ++k) u[k] = (xyz[k] - roi[k]) / (roi[3 + k] - roi[k]);
    if (type == IA_UN_BOUNDED_SPHERE) {
        float v[3];
        for (int k = 0; k < 3; ++k) v[k] = u[k] * 2.0f - 1.0f;
        float nsq = fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0]));
        float n = sqrtf(nsq);
        if (n > 1.0f) {
            float s = 2.0f - 1.0f / n;
            for (int k = 0; k < 3; ++k) v[k] = s * (v[k] / n);
        }
        for (int k = 0; k < 3; ++k) u[k] = v[k] * 0.25f + 0.5f;
    }
}

static inline int occupied_at(const float xyz[3], const float *roi, int type, const int *res,
                              const uint8_t *binary)
{
    if (type == IA_AABB) {
        for (int k = 0; k < 3; ++k)
            if (xyz[k] < roi[k] || xyz[k] > roi[3 + k]) return 0;
    }
    float u[3];
    to_unit(xyz, roi, type, u);
    int ix = clampi((int)(u[0] * (float)res[0]), 0, res[0] - 1);
    int iy = clampi((int)(u[1] * (float)res[1]), 0, res[1] - 1);
    int iz = clampi((int)(u[2] * (float)res[2]), 0, res[2] - 1);
    int64_t idx = (int64_t)ix * res[1] * res[2] + (int64_t)iy * res[2] + iz;
    return binary[idx] != 0;
}

static inline float distance_to_next_voxel(const float xyz[3], const float d[3], const float inv_d[3],
                                           const float *roi, const int *res)
{
    float t = INFINITY;
    float best[3];
    for (int k = 0; k < 3; ++k) {
        float r = (float)res[k];
        float p = ((xyz[k] - roi[k]) / (roi[3 + k] - roi[k])) * r;
        float target = floorf(p + 0.5f + 0.5f * sgn(d[k]));
        best[k] = ((target - p) * inv_d[k]) / r * (roi[3 + k] - roi[k]);
    }
    t = fminf(fminf(best[0], best[1]), best[2]);
    return fmaxf(t, 0.0f);
}

/* One pass of nerfacc's ray_marching_kernel.  packed_info == NULL: count pass (num_steps[r] written).
 * Otherwise packed_info[r] = (offset, count) and the three outputs are written. */
void ia_ref_march(const float *o, const float *d, const float *t_min, const float *t_max, int64_t n_rays,
                  const float *roi, const uint8_t *binary, const int *res, int type, float step_size,
                  float cone_angle, const int32_t *packed_info, int32_t *num_steps, int32_t *ray_indices,
                  float *t_starts, float *t_ends)
{
    const float dt_min = step_size, dt_max = 1e10f;
    for (int64_t i = 0; i < n_rays; ++i) {
        const float *ro = o + 3 * i, *rd = d + 3 * i;
        const float inv_d[3] = {1.0f / rd[0], 1.0f / rd[1], 1.0f / rd[2]};
        const float near = t_min[i], far = t_max[i];
        int64_t base = packed_info ? packed_info[2 * i] : 0;
        int j = 0;
        float t0 = near;
        float dt = calc_dt(t0, cone_angle, dt_min, dt_max);
        float t1 = t0 + dt;
        float t_mid = (t0 + t1) * 0.5f;
        while (t_mid < far) {
            float xyz[3];
            for (int k = 0; k < 3; ++k) xyz[k] = fmaf(t_mid, rd[k], ro[k]);
            if (occupied_at(xyz, roi, type, res, binary)) {
                if (packed_info) {
                    t_starts[base + j] = t0;
                    t_ends[base + j] = t1;
                    ray_indices[base + j] = (int32_t)i;
                }
                ++j;
                t0 = t1;
                t1 = t0 + calc_dt(t0, cone_angle, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            } else if (type == IA_AABB) {
                float t_target = t_mid + distance_to_next_voxel(xyz, rd, inv_d, roi, res);
                t_target = fminf(t_target, far);
                do { t_mid += dt_min; } while (t_mid < t_target);
                dt = calc_dt(t_mid, cone_angle, dt_min, dt_max);
                t0 = t_mid - dt * 0.5f;
                t1 = t_mid + dt * 0.5f;
            } else {
                t0 = t1;
                t1 = t0 + calc_dt(t0, cone_angle, dt_min, dt_max);
                t_mid = (t0 + t1) * 0.5f;
            }
        }
        if (!packed_info) num_steps[i] = j;
    }
}

/* nerfacc render_transmittance_from_alpha (sequential per-ray product) + render_visibility:
 * vis[j] = (T_j >= early_stop_eps) && (alpha_thre <= 0 || alpha_j >= alpha_thre),  T_j = prod_{k<j}(1-a_k). */
void ia_ref_visibility(const float *alphas, const int32_t *packed_info, int64_t n_rays, float early_stop_eps,
                       float alpha_thre, uint8_t *vis)
{
    for (int64_t i = 0; i < n_rays; ++i) {
        int64_t base = packed_info[2 * i];
        int n = packed_info[2 * i + 1];
        float T = 1.0f;
        for (int j = 0; j < n; ++j) {
            float a = alphas[base + j];
            int v = T >= early_stop_eps;
            if (alpha_thre > 0.0f) v = v && (a >= alpha_thre);
            vis[base + j] = (uint8_t)v;
            T = T * (1.0f - a);
        }
    }
}
