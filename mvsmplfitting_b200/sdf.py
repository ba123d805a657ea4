"""Drop-in for the reference's `sdf` extension package (sdf/sdf/sdf.py:8-26: SDFFunction, SDF, sdf):
`SDF()(faces_int32, vertices_in_[-1,1], grid_size) -> phi[B,G,G,G]`, no gradient.  The kernel is
mvs_sdf_grid in libmvsmpl (sm_100a) instead of the torch C++ extension that no longer builds."""
from __future__ import annotations

import torch
import torch.nn as nn

_ctx_cache = {}


def _ctx(device):
    from .context import FittingContext
    key = torch.device(device).index or 0
    if key not in _ctx_cache:
        _ctx_cache[key] = FittingContext(key)
    return _ctx_cache[key]


def sdf(faces, vertices, grid_size=32):
    """same call contract as the reference op: `faces.size(0)` is the number of triangles the kernel loops
    over (sdf_cuda_kernel.cu:314) -- so faces shaped [1,F,3], as fitting.py:367 passes them, means ONE."""
    if not vertices.is_cuda:
        raise RuntimeError("sdf: vertices must be a CUDA tensor (there is no CPU implementation)")
    with torch.no_grad():
        return _ctx(vertices.device).sdf_grid(faces.reshape(-1, 3), vertices, grid_size, num_faces=int(faces.shape[0]))


class SDF(nn.Module):
    def forward(self, faces, vertices, grid_size=32):
        return sdf(faces, vertices, grid_size)
