// Accelerated all-faces SDF: phi at a voxel centre c in the reference's INTENDED semantics
// (sdf_cuda_kernel.cu:242-304 looping over ALL F triangles):
//     phi(c) = 0                                  if an even number of triangles is hit by the ray c -> (-1,-1,-1)
//              min over triangles of dist(c, t)   otherwise
// evaluated over CANDIDATE LISTS instead of all 13 776 triangles, with the SAME primitives (mvs_sdf_geom.cuh) on the same
// float operands, so that the result is bit-identical to the brute force:
//   (1) parity: every ray ends in the corner O = (-1,-1,-1), so the central projection from O
//           pi(p) = ((p_y + 1), (p_z + 1)) / ((p_x + 1) + (p_y + 1) + (p_z + 1))
//       maps a ray to a POINT and a triangle to a triangle (all coordinates > -1: the mesh lives in +-0.84 of its box).  Triangles
//       are binned by the eps-padded bounding box of their projection into kBinR x kBinR bins over the projected mesh; only the
//       triangles in the bin of pi(c) can be hit.
//   (2) distance: triangles are binned by their eps-padded bounding boxes into kBinC^3 cells over [-1,1]^3.  A triangle that touches
//       no cell within Chebyshev ring r of c's cell is at least r * h away; rings are added until the running minimum is below that.
// Both structures depend on the posed vertices and are rebuilt per frame and closure evaluation (sdf_bins_kernel, mvs_sdf.cu: a
// counting sort in shared memory).  The functions here are __host__ __device__: tests/hostsim builds the same structures serially
// on the CPU and checks phi against the brute force, voxel by voxel, bit for bit (tests/test_hostsim_sdf_bins.py).
// Algorithm prototyped in oracle/sdf_binned.py (float64 binning); this is the float version with wider safety margins.
#pragma once
#include "mvs_sdf_geom.cuh"

namespace mvs {

constexpr int kBinC = 32;                          // distance cells per axis, h = 1/16
constexpr int kBinR = 128;                         // ray bins per axis
constexpr int kBinCells = kBinC * kBinC * kBinC;
constexpr int kBinRays = kBinR * kBinR;
constexpr int kBinCapD = 1 << 17;                  // entries of a frame's cell lists (measured: ~60 k on the synthetic mesh)
constexpr int kBinCapR = 1 << 18;                  // entries of a frame's ray-bin lists (~110 k)
constexpr float kBinEpsCell = 1e-5f;               // padding of triangle boxes (box coordinates)
constexpr float kBinEpsRay = 1e-4f;                // padding of projected boxes (projection coordinates, range (0, 1))

struct SdfBinsView {                               // one frame's structures
    const float* tri;                              // [F][9] triangle corners in box coordinates (the brute force's operands)
    const int* cell_ptr;                           // [kBinCells + 1]
    const unsigned short* cell_idx;
    const int* ray_ptr;                            // [kBinRays + 1]
    const unsigned short* ray_idx;
    float s_lo[2], s_scale[2];                     // bin = floor((pi - s_lo) * s_scale)
};

MVS_HD void sdf_project(const float* p, float* s) {
    const float q0 = p[0] + 1.f, q1 = p[1] + 1.f, q2 = p[2] + 1.f;
    const float sum = (q0 + q1) + q2;
    s[0] = q1 / sum; s[1] = q2 / sum;
}

MVS_HD int bin_clamp(float v, int n) {
    const float f = floorf(v);
    return f < 0.f ? 0 : (f > (float)(n - 1) ? n - 1 : (int)f);
}

// cells touched by the padded bounding box of a triangle (p9: three corners), inclusive ranges
MVS_HD void tri_cell_range(const float* p9, int* lo, int* hi) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float mn = fminf(p9[a], fminf(p9[3 + a], p9[6 + a])) - kBinEpsCell;
        const float mx = fmaxf(p9[a], fmaxf(p9[3 + a], p9[6 + a])) + kBinEpsCell;
        lo[a] = bin_clamp((mn + 1.f) * (0.5f * kBinC), kBinC);
        hi[a] = bin_clamp((mx + 1.f) * (0.5f * kBinC), kBinC);
    }
}

// bounding box of the projected triangle (projection coordinates)
MVS_HD void tri_proj_box(const float* p9, float* smin, float* smax) {
    float s0[2], s1[2], s2[2];
    sdf_project(p9, s0); sdf_project(p9 + 3, s1); sdf_project(p9 + 6, s2);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        smin[a] = fminf(s0[a], fminf(s1[a], s2[a]));
        smax[a] = fmaxf(s0[a], fmaxf(s1[a], s2[a]));
    }
}

// s_lo / s_scale of the frame from the bounding box of the projected mesh
MVS_HD void ray_bin_frame(const float* mesh_smin, const float* mesh_smax, float* s_lo, float* s_scale) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        s_lo[a] = mesh_smin[a] - 2.f * kBinEpsRay;
        s_scale[a] = (float)kBinR / ((mesh_smax[a] + 2.f * kBinEpsRay) - s_lo[a]);
    }
}

MVS_HD void tri_ray_range(const float* smin, const float* smax, const float* s_lo, const float* s_scale, int* lo, int* hi) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        lo[a] = bin_clamp(((smin[a] - kBinEpsRay) - s_lo[a]) * s_scale[a], kBinR);
        hi[a] = bin_clamp(((smax[a] + kBinEpsRay) - s_lo[a]) * s_scale[a], kBinR);
    }
}

// phi at voxel centre c over the frame's candidate lists.  evals (may be null): [0] += ray candidates, [1] += distance candidates.
MVS_HD float phi_binned(const float* c, const SdfBinsView& v, int* evals) {
    // (1) parity of ray crossings
    float s[2];
    sdf_project(c, s);
    const float sf0 = (s[0] - v.s_lo[0]) * v.s_scale[0], sf1 = (s[1] - v.s_lo[1]) * v.s_scale[1];
    if (!(sf0 >= 0.f && sf0 < (float)kBinR && sf1 >= 0.f && sf1 < (float)kBinR)) return 0.f;     // projects outside the mesh
    const int rb = (int)floorf(sf0) * kBinR + (int)floorf(sf1);
    int hits = 0;
    for (int e = v.ray_ptr[rb]; e < v.ray_ptr[rb + 1]; ++e) {
        const float* p = v.tri + 9 * (int)v.ray_idx[e];
        if (ray_hits(c, p, p + 3, p + 6)) ++hits;
    }
    if (evals) evals[0] += v.ray_ptr[rb + 1] - v.ray_ptr[rb];
    if (hits % 2 == 0) return 0.f;
    // (2) minimum distance: the 27 cells around c's cell, then shell after shell
    int cc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) cc[a] = bin_clamp((c[a] + 1.f) * (0.5f * kBinC), kBinC);
    const float h = 2.f / (float)kBinC;
    float min_d = 1000.f;
    for (int r = 1; r <= kBinC; ++r) {
        const int x0 = cc[0] - r < 0 ? 0 : cc[0] - r, x1 = cc[0] + r > kBinC - 1 ? kBinC - 1 : cc[0] + r;
        const int y0 = cc[1] - r < 0 ? 0 : cc[1] - r, y1 = cc[1] + r > kBinC - 1 ? kBinC - 1 : cc[1] + r;
        const int z0 = cc[2] - r < 0 ? 0 : cc[2] - r, z1 = cc[2] + r > kBinC - 1 ? kBinC - 1 : cc[2] + r;
        for (int x = x0; x <= x1; ++x)
            for (int y = y0; y <= y1; ++y)
                for (int z = z0; z <= z1; ++z) {
                    // rings beyond the first only add their shell (the inner cube was searched by the previous rings)
                    if (r > 1) {
                        const int dx = x > cc[0] ? x - cc[0] : cc[0] - x, dy = y > cc[1] ? y - cc[1] : cc[1] - y;
                        const int dz = z > cc[2] ? z - cc[2] : cc[2] - z;
                        const int dm = dx > dy ? (dx > dz ? dx : dz) : (dy > dz ? dy : dz);
                        if (dm < r) continue;
                    }
                    const int cell = (x * kBinC + y) * kBinC + z;
                    const int e0 = v.cell_ptr[cell], e1 = v.cell_ptr[cell + 1];
                    for (int e = e0; e < e1; ++e) {
                        const float* p = v.tri + 9 * (int)v.cell_idx[e];
                        const float dd = triangle_distance(c, p, p + 3, p + 6);
                        if (dd < min_d) min_d = dd;
                    }
                    if (evals) evals[1] += e1 - e0;
                }
        // a triangle touching no cell within ring r is at least r * h away (c lies inside its own cell); margin for the float
        // cell assignment of c and of the padded boxes
        if (min_d < (float)r * h * (1.f - 1e-4f)) break;
    }
    return min_d;
}

#if defined(__CUDACC__)
// The same evaluation by a whole warp for ONE voxel centre (all 32 lanes pass the same c): the lanes share out the candidates of
// every list, the parity is a sum and the distance a minimum over the lanes -- both exact, so the result is phi_binned's bit for bit.
__device__ __noinline__ float phi_binned_warp(const float* c, const SdfBinsView* vp) {
    const SdfBinsView& v = *vp;
    const int lane = threadIdx.x & 31;
    float s[2];
    sdf_project(c, s);
    const float sf0 = (s[0] - v.s_lo[0]) * v.s_scale[0], sf1 = (s[1] - v.s_lo[1]) * v.s_scale[1];
    if (!(sf0 >= 0.f && sf0 < (float)kBinR && sf1 >= 0.f && sf1 < (float)kBinR)) return 0.f;
    const int rb = (int)floorf(sf0) * kBinR + (int)floorf(sf1);
    int hits = 0;
    for (int e = v.ray_ptr[rb] + lane; e < v.ray_ptr[rb + 1]; e += 32) {
        const float* p = v.tri + 9 * (int)v.ray_idx[e];
        if (ray_hits(c, p, p + 3, p + 6)) ++hits;
    }
    hits = __reduce_add_sync(0xffffffffu, hits);
    if (hits % 2 == 0) return 0.f;
    int cc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) cc[a] = bin_clamp((c[a] + 1.f) * (0.5f * kBinC), kBinC);
    const float h = 2.f / (float)kBinC;
    float min_d = 1000.f;
    for (int r = 1; r <= kBinC; ++r) {
        const int x0 = cc[0] - r < 0 ? 0 : cc[0] - r, x1 = cc[0] + r > kBinC - 1 ? kBinC - 1 : cc[0] + r;
        const int y0 = cc[1] - r < 0 ? 0 : cc[1] - r, y1 = cc[1] + r > kBinC - 1 ? kBinC - 1 : cc[1] + r;
        const int z0 = cc[2] - r < 0 ? 0 : cc[2] - r, z1 = cc[2] + r > kBinC - 1 ? kBinC - 1 : cc[2] + r;
        for (int x = x0; x <= x1; ++x)
            for (int y = y0; y <= y1; ++y) {
                // the cells of one (x, y) column are consecutive in memory: their lists form ONE contiguous range
                int za = z0, zb = z1;
                if (r > 1) {
                    const int dx = x > cc[0] ? x - cc[0] : cc[0] - x, dy = y > cc[1] ? y - cc[1] : cc[1] - y;
                    if ((dx > dy ? dx : dy) < r) {          // inner column: only its two end cells belong to the shell
                        for (int pass = 0; pass < 2; ++pass) {
                            const int z = pass ? cc[2] + r : cc[2] - r;
                            if (z < 0 || z > kBinC - 1) continue;
                            const int cell = (x * kBinC + y) * kBinC + z;
                            for (int e = v.cell_ptr[cell] + lane; e < v.cell_ptr[cell + 1]; e += 32) {
                                const float* p = v.tri + 9 * (int)v.cell_idx[e];
                                const float dd = triangle_distance(c, p, p + 3, p + 6);
                                if (dd < min_d) min_d = dd;
                            }
                        }
                        continue;
                    }
                }
                const int base = (x * kBinC + y) * kBinC;
                for (int e = v.cell_ptr[base + za] + lane; e < v.cell_ptr[base + zb + 1]; e += 32) {
                    const float* p = v.tri + 9 * (int)v.cell_idx[e];
                    const float dd = triangle_distance(c, p, p + 3, p + 6);
                    if (dd < min_d) min_d = dd;
                }
            }
        min_d = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(min_d)));      // distances are >= 0: same order as uint
        if (min_d < (float)r * h * (1.f - 1e-4f)) break;
    }
    return min_d;
}
#endif

}  // namespace mvs
