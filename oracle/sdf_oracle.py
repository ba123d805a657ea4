"""TEST INFRASTRUCTURE -- CPU restatement of the reference's SDF interpenetration
term: the voxel kernel (oracle/sdf_ref.c, restating sdf/sdf/csrc/sdf_cuda_kernel.cu)
plus the Python glue of code/utils/fitting.py:352-393 (bounding box, 1.2 x 0.5
scale, no-grad normalisation, grid_sample with its defaults, squared weighted sum).

Not a product path.  Pinned on the GPU box against the reference's own kernel compiled
unchanged (oracle/build_ref_sdf.sh -> oracle/_ref/libsdf_refcuda.so) by
tests/test_gpu_sdf_refpin.py, voxel for voxel.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = None
GRID_FN = None        # tests may plug in another voxeliser with sdf_grid's signature (the reference's own CUDA kernel,
                      # oracle/ref_sdf.py) for penetration_loss; None = the C restatement below


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp oracle/sdf_ref.c -> oracle/_build/libsdf_ref.so"""
    os.makedirs(_BUILD, exist_ok=True)
    so = os.path.join(_BUILD, "libsdf_ref.so")
    src = os.path.join(_HERE, "sdf_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", so, src, "-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.sdf_ref_grid.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4
        _LIB.sdf_ref_voxels.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    return _LIB


def sdf_grid(faces: np.ndarray, verts_norm: np.ndarray, grid_size: int, all_faces: bool = False) -> np.ndarray:
    """phi [B,G,G,G] float32.  `all_faces=False` reproduces the reference call
    fitting.py:367 (faces reshaped to [1,F,3] so the kernel sees num_faces == 1)."""
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    v = np.ascontiguousarray(verts_norm, dtype=np.float32)
    B, N = v.shape[0], v.shape[1]
    phi = np.zeros((B, grid_size, grid_size, grid_size), dtype=np.float32)
    nf = faces.shape[0] if all_faces else 1
    _lib().sdf_ref_grid(phi.ctypes.data, faces.ctypes.data, v.ctypes.data, B, nf, N, grid_size)
    return phi


def sdf_voxels(faces, verts_norm, grid_size, voxel_ids, all_faces=True):
    faces = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    v = np.ascontiguousarray(verts_norm, dtype=np.float32).reshape(-1, 3)
    ids = np.ascontiguousarray(voxel_ids, dtype=np.int64)
    out = np.zeros(ids.shape[0], dtype=np.float32)
    nf = faces.shape[0] if all_faces else 1
    _lib().sdf_ref_voxels(out.ctypes.data, ids.ctypes.data, ids.shape[0], faces.ctypes.data, v.ctypes.data,
                          nf, grid_size)
    return out


def penetration_loss(vertices: torch.Tensor, faces: torch.Tensor, coll_loss_weight: float,
                     grid_size: int = 128, all_faces: bool = False) -> torch.Tensor:
    """fitting.py:352-393 for one person.  vertices [N,3] (autograd-tracked, with transl).
    Gradient flows through the sample coordinates AND through the bounding-box
    centre/scale (min/max vertices), not through phi (sdf.py:17-19)."""
    v = vertices.unsqueeze(0)                                          # [1,N,3]
    lo = v[0].min(dim=0)[0]
    hi = v[0].max(dim=0)[0]
    boxes = torch.stack([lo, hi]).unsqueeze(0).to(torch.float32)       # fitting.py:282-288 (float32 zeros buffer)
    boxes = boxes.to(v.dtype) if v.dtype != torch.float32 else boxes
    center = boxes.mean(dim=1).unsqueeze(1)                            # [1,1,3]
    scale = (1 + 0.2) * 0.5 * (boxes[:, 1] - boxes[:, 0]).max(dim=-1)[0][:, None, None]
    with torch.no_grad():
        vn = (v - center) / scale
        phi = (GRID_FN or sdf_grid)(faces.numpy(), vn.to(torch.float32).numpy(), grid_size, all_faces=all_faces)
        phi = torch.from_numpy(phi).to(v.dtype)
    local = (v - center[0].unsqueeze(0)) / scale[0].unsqueeze(0)
    grid = local.view(1, -1, 1, 1, 3)
    val = torch.nn.functional.grid_sample(phi[0][None, None], grid, align_corners=False).view(1, -1)
    w = torch.tensor(coll_loss_weight, dtype=v.dtype)
    return (w * val.sum() / 1) ** 2
