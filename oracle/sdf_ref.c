/* TEST INFRASTRUCTURE -- plain-C (float) restatement of the reference's brute-force
 * voxel SDF kernel, sdf/sdf/csrc/sdf_cuda_kernel.cu.  One call fills phi[B][G][G][G].
 * The reference kernel is CUDA + ATen and does not build against torch 2.11
 * (AT_CHECK / .type() / .data<T>()), so it is restated here; this file
 * is compiled by oracle/sdf_oracle.py / __graft_entry__.build() with plain gcc.
 *
 * Cited lines are of sdf_cuda_kernel.cu:
 *   point_segment_distance :73-92, intersect_triangle :95-138 (EPSILON 1e-6 :8),
 *   point_triangle_distance :155-237 (max(det,1e-30f) clamp :172),
 *   voxel loop :242-304 (centre formula :260-263, hit counted iff t >= 0 :285,
 *   parity rule :291-293, min_distance init 1000 :267),
 *   host shape logic :312-317 (num_faces = faces.size(0); blocks = total/512,
 *   integer division, so a tail of < 512 voxels is left at its initial 0).
 */
#include <math.h>
#include <stdint.h>

#define EPS_RAY 0.000001

static float mag2f(const float* x) { float l = 0; for (int i = 0; i < 3; ++i) l += x[i] * x[i]; return l; }
static float dot3f(const float* x, const float* y) { float l = 0; for (int i = 0; i < 3; ++i) l += x[i] * y[i]; return l; }
static float distf(const float* x, const float* y) {
    float l = 0, d;
    for (int i = 0; i < 3; ++i) { d = x[i] - y[i]; l += d * d; }
    return sqrtf(l);
}

static float seg_dist(const float* x0, const float* x1, const float* x2, float* r) {
    float dx[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
    float m2 = mag2f(dx);
    float s12 = (float)(dot3f(x2, dx) - dot3f(x0, dx)) / m2;
    if (s12 < 0) s12 = 0; else if (s12 > 1) s12 = 1;
    for (int i = 0; i < 3; ++i) r[i] = s12 * x1[i] + (1 - s12) * x2[i];
    return distf(x0, r);
}

/* Moller-Trumbore as written in the reference (double constants, float data). */
static int ray_tri(const float* orig, const float* dir, const float* v0, const float* v1,
                   const float* v2, float* t) {
    float e1[3], e2[3], tv[3], pv[3], qv[3], det, inv_det, u, v;
    for (int i = 0; i < 3; ++i) { e1[i] = v1[i] - v0[i]; e2[i] = v2[i] - v0[i]; }
    pv[0] = dir[1] * e2[2] - dir[2] * e2[1];
    pv[1] = dir[2] * e2[0] - dir[0] * e2[2];
    pv[2] = dir[0] * e2[1] - dir[1] * e2[0];
    det = e1[0] * pv[0] + e1[1] * pv[1] + e1[2] * pv[2];
    if (det > -EPS_RAY && det < EPS_RAY) return 0;
    inv_det = (float)(1.0 / det);
    for (int i = 0; i < 3; ++i) tv[i] = orig[i] - v0[i];
    u = (tv[0] * pv[0] + tv[1] * pv[1] + tv[2] * pv[2]) * inv_det;
    if (u < 0.0 || u > 1.0) return 0;
    qv[0] = tv[1] * e1[2] - tv[2] * e1[1];
    qv[1] = tv[2] * e1[0] - tv[0] * e1[2];
    qv[2] = tv[0] * e1[1] - tv[1] * e1[0];
    v = (dir[0] * qv[0] + dir[1] * qv[1] + dir[2] * qv[2]) * inv_det;
    if (v < 0.0 || (u + v) > 1.0) return 0;
    *t = (e2[0] * qv[0] + e2[1] * qv[1] + e2[2] * qv[2]) * inv_det;
    return 1;
}

static void tri_closest(const float* x0, const float* x1, const float* x2, const float* x3, float* r) {
    float x13[3], x23[3], x03[3];
    for (int i = 0; i < 3; ++i) { x13[i] = x1[i] - x3[i]; x23[i] = x2[i] - x3[i]; x03[i] = x0[i] - x3[i]; }
    float m13 = mag2f(x13), m23 = mag2f(x23), d = dot3f(x13, x23);
    float invdet = 1.f / fmaxf(m13 * m23 - d * d, 1e-30f);
    float a = dot3f(x13, x03), b = dot3f(x23, x03);
    float w23 = invdet * (m23 * a - d * b);
    float w31 = invdet * (m13 * b - d * a);
    float w12 = 1 - w23 - w31;
    if (w23 >= 0 && w31 >= 0 && w12 >= 0) {
        for (int i = 0; i < 3; ++i) r[i] = w23 * x1[i] + w31 * x2[i] + w12 * x3[i];
        return;
    }
    float r1[3], r2[3], d1, d2;
    if (w23 > 0)      { d1 = seg_dist(x0, x1, x2, r1); d2 = seg_dist(x0, x1, x3, r2); }
    else if (w31 > 0) { d1 = seg_dist(x0, x1, x2, r1); d2 = seg_dist(x0, x2, x3, r2); }
    else              { d1 = seg_dist(x0, x1, x3, r1); d2 = seg_dist(x0, x2, x3, r2); }
    const float* best = (d1 < d2) ? r1 : r2;
    for (int i = 0; i < 3; ++i) r[i] = best[i];
}

/* phi value of one voxel (bn, k, j, i) */
static float voxel_phi(const int32_t* faces, const float* verts, int num_faces, int G, int i, int j, int k) {
    const float dx = 2. / (G - 1);
    const float c[3] = {(float)(-1 + (i + 0.5) * dx), (float)(-1 + (j + 0.5) * dx), (float)(-1 + (k + 0.5) * dx)};
    int hits = 0;
    float min_d = 1000;
    for (int f = 0; f < num_faces; ++f) {
        const float* v1 = verts + 3 * faces[3 * f + 0];
        const float* v2 = verts + 3 * faces[3 * f + 1];
        const float* v3 = verts + 3 * faces[3 * f + 2];
        float cp[3], t;
        tri_closest(c, v1, v2, v3, cp);
        float dist = distf(c, cp);
        if (dist < min_d) min_d = dist;
        float dir[3] = {-1.0f - c[0], -1.0f - c[1], -1.0f - c[2]};
        if (ray_tri(c, dir, v1, v2, v3, &t) && t >= 0) hits++;
    }
    if (hits % 2 == 0) min_d = 0.f;
    return min_d;
}

/* phi: [B][G][G][G] zero-initialised by the caller; faces: int32 [*,3]; verts: [B][num_verts][3]
 * (already normalised to [-1,1]).  `num_faces` is what the reference host code passes
 * (faces.size(0)): 1 for the call made by fitting.py:367, F for the intended semantics. */
void sdf_ref_grid(float* phi, const int32_t* faces, const float* verts, int batch, int num_faces,
                  int num_verts, int G) {
    const long total = (long)batch * G * G * G;
    const long covered = (total / 512) * 512;            /* blocks = total / threads (:316-317) */
#pragma omp parallel for schedule(static)
    for (long tid = 0; tid < covered; ++tid) {
        int i = tid % G, j = (tid / G) % G, k = (tid / ((long)G * G)) % G;
        int bn = tid / ((long)G * G * G);
        phi[tid] = voxel_phi(faces, verts + (long)bn * num_verts * 3, num_faces, G, i, j, k);
    }
}

/* phi at an explicit list of voxel ids (for sparse checks of the all-faces semantics) */
void sdf_ref_voxels(float* out, const int64_t* voxel_ids, long n, const int32_t* faces,
                    const float* verts, int num_faces, int G) {
#pragma omp parallel for schedule(static)
    for (long q = 0; q < n; ++q) {
        long tid = voxel_ids[q];
        int i = tid % G, j = (tid / G) % G, k = (tid / ((long)G * G)) % G;
        out[q] = voxel_phi(faces, verts, num_faces, G, i, j, k);
    }
}

/* ---- candidate-list evaluation (prototype of the accelerated all-faces mode) ----
 * Same arithmetic as voxel_phi, but the distance minimum runs over dist_idx[dist_ptr[q] .. dist_ptr[q+1]) and the
 * ray-parity count over ray_idx[ray_ptr[q] .. ray_ptr[q+1]) instead of all faces.  Returns min_d BEFORE the parity
 * rule in out_d and the hit count in out_hits, so that the caller can decide whether the distance list was large
 * enough (ring search) before combining them.  oracle/sdf_binned.py builds the lists and proves they reproduce the
 * brute force bit for bit. */
void sdf_ref_voxels_lists(float* out_d, int32_t* out_hits, const int64_t* voxel_ids, long n, const int32_t* faces,
                          const float* verts, int G, const int64_t* dist_ptr, const int32_t* dist_idx,
                          const int64_t* ray_ptr, const int32_t* ray_idx) {
    const float dx = 2. / (G - 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (long q = 0; q < n; ++q) {
        long tid = voxel_ids[q];
        int i = tid % G, j = (tid / G) % G, k = (tid / ((long)G * G)) % G;
        const float c[3] = {(float)(-1 + (i + 0.5) * dx), (float)(-1 + (j + 0.5) * dx), (float)(-1 + (k + 0.5) * dx)};
        float min_d = 1000;
        for (int64_t e = dist_ptr[q]; e < dist_ptr[q + 1]; ++e) {
            const int f = dist_idx[e];
            float cp[3];
            tri_closest(c, verts + 3 * faces[3 * f], verts + 3 * faces[3 * f + 1], verts + 3 * faces[3 * f + 2], cp);
            float dist = distf(c, cp);
            if (dist < min_d) min_d = dist;
        }
        int hits = 0;
        float dir[3] = {-1.0f - c[0], -1.0f - c[1], -1.0f - c[2]};
        for (int64_t e = ray_ptr[q]; e < ray_ptr[q + 1]; ++e) {
            const int f = ray_idx[e];
            float t;
            if (ray_tri(c, dir, verts + 3 * faces[3 * f], verts + 3 * faces[3 * f + 1], verts + 3 * faces[3 * f + 2], &t) && t >= 0)
                hits++;
        }
        out_d[q] = min_d;
        out_hits[q] = hits;
    }
}
