"""TEST INFRASTRUCTURE / PROTOTYPE (never imported by the product) -- the algorithm of the accelerated all-faces SDF
mode, validated on the CPU against the brute-force restatement oracle/sdf_ref.c before any kernel
is written for it.

The reference kernel (sdf_cuda_kernel.cu:242-304), in its INTENDED all-faces form, computes for a voxel centre c
  (1) the minimum over ALL F triangles of the point-triangle distance, and
  (2) the parity of the number of triangles hit by the ray from c towards the box corner O = (-1,-1,-1),
F = 13 776 tests of each kind per voxel.  The product never builds the grid: phi is only needed at the <= 8 voxels
around every mesh vertex (`sdf_fused_kernel`), which all lie next to the surface.  Two exact culls then remove almost
every test; "exact" = the surviving candidates are evaluated with the SAME arithmetic and provably contain every
triangle that can change the result, so phi is bit-identical to the brute force:

  (1) distance: triangles are binned by their bounding boxes into a uniform C^3 cell grid over [-1,1]^3.  A triangle
      that touches no cell within Chebyshev ring r of the voxel's cell is farther than r * h (h = cell size).  Rings are
      added until the running minimum is below that bound -- next to the surface ring 1 is enough.
  (2) parity: every ray ends in the same point O, so a central projection from O maps each ray to a POINT and each
      triangle to a triangle: s = (p - O) / sum(p - O), 2-D coordinates (s_y, s_z) in (0,1)^2.  Triangles are
      rasterised conservatively (epsilon-padded) into an R x R grid over the projected mesh; only the triangles in the
      bin of the voxel's own projection can be hit.
Both bin structures depend on the posed vertices and are rebuilt per closure evaluation (13 776 triangles: a counting
sort in shared memory on the device)."""
from __future__ import annotations

import ctypes

import numpy as np

from oracle import sdf_oracle


def _csr(lists_owner, lists_item, n_owner):
    order = np.argsort(lists_owner, kind="stable")
    ptr = np.zeros(n_owner + 1, dtype=np.int64)
    np.add.at(ptr, lists_owner + 1, 1)
    return np.cumsum(ptr), lists_item[order].astype(np.int32)


def _bin_boxes(lo, hi, n_bins):
    """(owner bin id, triangle id) pairs for integer boxes lo..hi (inclusive) in a d-dimensional bin grid"""
    d = lo.shape[1]
    owners, items = [], []
    ext = hi - lo + 1
    for f in range(lo.shape[0]):
        rngs = [np.arange(lo[f, a], hi[f, a] + 1) for a in range(d)]
        mesh = np.stack(np.meshgrid(*rngs, indexing="ij"), -1).reshape(-1, d)
        flat = np.zeros(mesh.shape[0], dtype=np.int64)
        for a in range(d):
            flat = flat * n_bins + mesh[:, a]
        owners.append(flat)
        items.append(np.full(mesh.shape[0], f, dtype=np.int32))
    del ext
    return np.concatenate(owners), np.concatenate(items)


class BinnedSdf:
    def __init__(self, faces, verts_norm, grid_size, cells=128, ray_bins=256, eps=1e-5, raster=True):
        self.faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
        self.verts = np.ascontiguousarray(verts_norm, dtype=np.float32).reshape(-1, 3)
        self.G, self.C, self.R = int(grid_size), int(cells), int(ray_bins)
        tri = self.verts[self.faces].astype(np.float64)                       # [F,3,3]
        # (1) distance cells over [-1,1]^3 (coordinates outside are clamped into the border cells)
        self.h = 2.0 / self.C
        lo = np.clip(np.floor((tri.min(1) + 1.0) / self.h).astype(np.int64), 0, self.C - 1)
        hi = np.clip(np.floor((tri.max(1) + 1.0) / self.h).astype(np.int64), 0, self.C - 1)
        own, it = _bin_boxes(lo, hi, self.C)
        self.cell_ptr, self.cell_idx = _csr(own, it, self.C ** 3)
        # (2) central projection from O = (-1,-1,-1)
        s = self._project(tri.reshape(-1, 3)).reshape(-1, 3, 2)
        # the R x R bins cover the bounding box of the PROJECTED MESH only (a body seen from the corner fills a few per
        # cent of the unit square); a query point that projects outside of it cannot be hit by anything
        self.s_lo = s.reshape(-1, 2).min(0) - 2 * eps
        self.s_scale = self.R / (s.reshape(-1, 2).max(0) + 2 * eps - self.s_lo)
        lo = np.clip(np.floor((s.min(1) - eps - self.s_lo) * self.s_scale).astype(np.int64), 0, self.R - 1)
        hi = np.clip(np.floor((s.max(1) + eps - self.s_lo) * self.s_scale).astype(np.int64), 0, self.R - 1)
        own, it = _bin_boxes(lo, hi, self.R)
        if raster:
            # conservative rasterisation: keep (bin, triangle) only if the eps-padded bin rectangle is not separated from
            # the projected triangle by one of the triangle's edge lines (the bounding-box axes are already tested)
            bx, by = own // self.R, own % self.R
            pad = eps * self.s_scale
            x0, x1 = bx - pad[0], bx + 1 + pad[0]
            y0, y1 = by - pad[1], by + 1 + pad[1]
            t = (s[it] - self.s_lo) * self.s_scale                                       # [n,3,2] in bin units
            keep = np.ones(len(own), dtype=bool)
            area = (t[:, 1, 0] - t[:, 0, 0]) * (t[:, 2, 1] - t[:, 0, 1]) - (t[:, 1, 1] - t[:, 0, 1]) * (t[:, 2, 0] - t[:, 0, 0])
            sgn = np.where(area >= 0, 1.0, -1.0)
            for a in range(3):
                p, q = t[:, a], t[:, (a + 1) % 3]
                ex, ey = q[:, 0] - p[:, 0], q[:, 1] - p[:, 1]
                # signed distance (times edge length) of the rectangle corner farthest INSIDE; all corners outside <=> max < 0
                best = np.full(len(own), -np.inf)
                for cx in (x0, x1):
                    for cy in (y0, y1):
                        best = np.maximum(best, sgn * (ex * (cy - p[:, 1]) - ey * (cx - p[:, 0])))
                tol = 1e-9 + 4 * np.abs(pad).max() * (np.abs(ex) + np.abs(ey))
                keep &= ~(best < -tol) | (np.abs(area) < 1e-12)                          # degenerate projections: keep
            own, it = own[keep], it[keep]
        self.ray_ptr, self.ray_idx = _csr(own, it, self.R ** 2)

    @staticmethod
    def _project(p):
        q = np.asarray(p, np.float64) + 1.0
        return q[:, 1:3] / q.sum(1, keepdims=True)

    def centres(self, voxel_ids):
        G = self.G
        dx = np.float32(2.0 / (G - 1))
        ijk = np.stack([voxel_ids % G, (voxel_ids // G) % G, (voxel_ids // (G * G)) % G], 1)
        return (-1 + (ijk + 0.5) * dx).astype(np.float32)

    def _ring_lists(self, cell, r):
        """CSR of the triangles in the cells within Chebyshev distance r of each query cell (duplicates removed)"""
        C = self.C
        ptr, idx = [0], []
        offs = np.arange(-r, r + 1)
        for a, b, c in cell:
            xs = np.clip(a + offs, 0, C - 1)
            ys = np.clip(b + offs, 0, C - 1)
            zs = np.clip(c + offs, 0, C - 1)
            cells = np.unique((xs[:, None, None] * C + ys[None, :, None]) * C + zs[None, None, :])
            cand = np.unique(np.concatenate([self.cell_idx[self.cell_ptr[k]:self.cell_ptr[k + 1]] for k in cells])) \
                if len(cells) else np.zeros(0, np.int32)
            idx.append(cand.astype(np.int32))
            ptr.append(ptr[-1] + len(cand))
        return np.array(ptr, dtype=np.int64), (np.concatenate(idx) if idx else np.zeros(0, np.int32))

    def phi(self, voxel_ids):
        """phi at the listed voxels, identical to sdf_oracle.sdf_voxels(..., all_faces=True); also returns the number of
        distance / ray candidates evaluated per voxel and the ring reached"""
        ids = np.ascontiguousarray(voxel_ids, dtype=np.int64)
        n = ids.shape[0]
        c = self.centres(ids).astype(np.float64)
        cell = np.clip(np.floor((c + 1.0) / self.h).astype(np.int64), 0, self.C - 1)
        sf = (self._project(c) - self.s_lo) * self.s_scale
        outside = ((sf < 0) | (sf >= self.R)).any(1)
        sb = np.clip(np.floor(sf).astype(np.int64), 0, self.R - 1)
        rbin = sb[:, 0] * self.R + sb[:, 1]
        cnt = np.where(outside, 0, self.ray_ptr[rbin + 1] - self.ray_ptr[rbin])
        ray_ptr = np.zeros(n + 1, dtype=np.int64)
        ray_ptr[1:] = np.cumsum(cnt)
        ray_idx = np.concatenate([self.ray_idx[self.ray_ptr[b]:self.ray_ptr[b] + k] for b, k in zip(rbin, cnt)] +
                                 [np.zeros(0, np.int32)]).astype(np.int32)
        if len(ray_idx) == 0:
            ray_idx = np.zeros(1, np.int32)
        lib = sdf_oracle._lib()
        out_d = np.full(n, 1000.0, dtype=np.float32)
        hits = np.zeros(n, dtype=np.int32)
        n_dist = np.zeros(n, dtype=np.int64)
        ring = np.zeros(n, dtype=np.int64)
        todo = np.arange(n)
        r = 1
        empty_ptr = np.zeros(n + 1, dtype=np.int64)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        first = True
        while len(todo):
            dptr, didx = self._ring_lists(cell[todo], r)
            d = np.zeros(len(todo), dtype=np.float32)
            hh = np.zeros(len(todo), dtype=np.int32)
            sub_ids = np.ascontiguousarray(ids[todo])
            if first:
                rp = np.zeros(len(todo) + 1, dtype=np.int64)
                rp[1:] = np.cumsum(ray_ptr[todo + 1] - ray_ptr[todo])
                ri = ray_idx
            else:
                rp, ri = np.zeros(len(todo) + 1, dtype=np.int64), np.zeros(1, np.int32)
            didx = np.ascontiguousarray(didx if len(didx) else np.zeros(1, np.int32))
            lib.sdf_ref_voxels_lists(P(d), P(hh), P(sub_ids), ctypes.c_long(len(todo)), P(self.faces), P(self.verts),
                                     ctypes.c_int(self.G), P(dptr), P(didx), P(rp), P(np.ascontiguousarray(ri)))
            if first:
                hits[todo] = hh
                first = False
            out_d[todo] = d
            n_dist[todo] = np.diff(dptr)
            ring[todo] = r
            # a triangle outside ring r is at least r * h away (minus a rounding margin); the whole grid is covered at r = C
            done = (d.astype(np.float64) < r * self.h * (1 - 1e-6)) | (r >= self.C)
            todo = todo[~done]
            r += 1
        del empty_ptr
        phi = np.where(hits % 2 == 0, np.float32(0), out_d).astype(np.float32)
        return phi, dict(dist_candidates=n_dist, ray_candidates=np.diff(ray_ptr), ring=ring)
