// Geometric primitives of the reference's SDF kernel (sdf/sdf/csrc/sdf_cuda_kernel.cu), float arithmetic as there.
// __host__ __device__: tests/hostsim runs the same functions on the CPU (all-faces binning, mvs_sdf_bins.cuh).
#pragma once
#include "mvs_math.cuh"

namespace mvs {

// ---------------------------------------------------------------------------------- geometry (float, as the reference)
MVS_HD float dist3(const float* a, const float* b) {
    const float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
    return sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
}
MVS_HD float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// sdf_cuda_kernel.cu:73-92
MVS_HD float segment_distance(const float* x0, const float* x1, const float* x2, float* r) {
    const float dx[3] = {x2[0] - x1[0], x2[1] - x1[1], x2[2] - x1[2]};
    const float m2 = dot3(dx, dx);
    float s12 = (dot3(x2, dx) - dot3(x0, dx)) / m2;
    s12 = s12 < 0.f ? 0.f : (s12 > 1.f ? 1.f : s12);
    r[0] = s12 * x1[0] + (1.f - s12) * x2[0];
    r[1] = s12 * x1[1] + (1.f - s12) * x2[1];
    r[2] = s12 * x1[2] + (1.f - s12) * x2[2];
    return dist3(x0, r);
}
// sdf_cuda_kernel.cu:155-237 (closest point), returns the distance
MVS_HD float triangle_distance(const float* x0, const float* x1, const float* x2, const float* x3) {
    float x13[3], x23[3], x03[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { x13[i] = x1[i] - x3[i]; x23[i] = x2[i] - x3[i]; x03[i] = x0[i] - x3[i]; }
    const float m13 = dot3(x13, x13), m23 = dot3(x23, x23), d = dot3(x13, x23);
    const float invdet = 1.f / fmaxf(m13 * m23 - d * d, 1e-30f);
    const float a = dot3(x13, x03), b = dot3(x23, x03);
    const float w23 = invdet * (m23 * a - d * b);
    const float w31 = invdet * (m13 * b - d * a);
    const float w12 = 1.f - w23 - w31;
    float r[3];
    if (w23 >= 0.f && w31 >= 0.f && w12 >= 0.f) {
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = w23 * x1[i] + w31 * x2[i] + w12 * x3[i];
        return dist3(x0, r);
    }
    float r2[3], d1, d2;
    if (w23 > 0.f) { d1 = segment_distance(x0, x1, x2, r); d2 = segment_distance(x0, x1, x3, r2); }
    else if (w31 > 0.f) { d1 = segment_distance(x0, x1, x2, r); d2 = segment_distance(x0, x2, x3, r2); }
    else { d1 = segment_distance(x0, x1, x3, r); d2 = segment_distance(x0, x2, x3, r2); }
    // the reference returns the closest POINT and the caller re-measures the distance to it (:281-282)
    return (d1 < d2) ? dist3(x0, r) : dist3(x0, r2);
}
// sdf_cuda_kernel.cu:95-150: ray from the voxel centre towards (-1,-1,-1); hit counted iff t >= 0
MVS_HD bool ray_hits(const float* c, const float* v0, const float* v1, const float* v2) {
    const float dir[3] = {-1.f - c[0], -1.f - c[1], -1.f - c[2]};
    float e1[3], e2[3], tv[3], pv[3], qv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { e1[i] = v1[i] - v0[i]; e2[i] = v2[i] - v0[i]; }
    pv[0] = dir[1] * e2[2] - dir[2] * e2[1];
    pv[1] = dir[2] * e2[0] - dir[0] * e2[2];
    pv[2] = dir[0] * e2[1] - dir[1] * e2[0];
    const float det = dot3(e1, pv);
    if (det > -1e-6 && det < 1e-6) return false;
    const float inv_det = (float)(1.0 / (double)det);
#pragma unroll
    for (int i = 0; i < 3; ++i) tv[i] = c[i] - v0[i];
    const float u = dot3(tv, pv) * inv_det;
    if (u < 0.f || u > 1.f) return false;
    qv[0] = tv[1] * e1[2] - tv[2] * e1[1];
    qv[1] = tv[2] * e1[0] - tv[0] * e1[2];
    qv[2] = tv[0] * e1[1] - tv[1] * e1[0];
    const float v = dot3(dir, qv) * inv_det;
    if (v < 0.f || (u + v) > 1.f) return false;
    const float t = dot3(e2, qv) * inv_det;
    return t >= 0.f;
}
MVS_HD void voxel_centre(int i, int j, int k, int G, float* c) {     // sdf_cuda_kernel.cu:260-263
    const float dx = (float)(2. / (G - 1));
    c[0] = (float)(-1 + (i + 0.5) * dx);
    c[1] = (float)(-1 + (j + 0.5) * dx);
    c[2] = (float)(-1 + (k + 0.5) * dx);
}

}  // namespace mvs
