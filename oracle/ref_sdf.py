"""TEST INFRASTRUCTURE -- the reference's OWN SDF CUDA kernel (sdf/sdf/csrc/sdf_cuda_kernel.cu, compiled unchanged into
oracle/_ref/libsdf_refcuda.so by oracle/build_ref_sdf.sh) behind a ctypes call, and a stand-in for the pybind module
`sdf.csrc` (sdf/sdf/csrc/sdf_cuda.cpp:14-28: input checks + forward to sdf_cuda) so that the reference's sdf/sdf/sdf.py
and code/utils/fitting.py:352-393 run unmodified on the GPU box.  Checker only; never used by the product path."""
from __future__ import annotations

import ctypes
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libsdf_refcuda.so")
_LIB = None


def available() -> bool:
    return os.path.exists(SO)


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(SO)
        _LIB.ref_sdf_cuda.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5
        _LIB.ref_sdf_cuda.restype = ctypes.c_int
    return _LIB


def sdf(phi: torch.Tensor, faces: torch.Tensor, vertices: torch.Tensor) -> torch.Tensor:
    """sdf.csrc.sdf(phi, faces, vertices) -> phi (sdf_cuda.cpp:14-28): CUDA + contiguous checked, phi written in place.
    faces.size(0) is what the kernel loops over (sdf_cuda_kernel.cu:314)."""
    for name, x in (("phi", phi), ("faces", faces), ("vertices", vertices)):
        assert x.is_cuda, name + " must be a CUDA tensor"
        assert x.is_contiguous(), name + " must be contiguous"
    assert faces.dtype == torch.int32 and phi.dtype == vertices.dtype and phi.dtype in (torch.float32, torch.float64)
    torch.cuda.synchronize(phi.device)                    # the reference launches on the legacy default stream (:321)
    with torch.cuda.device(phi.device):
        rc = _lib().ref_sdf_cuda(phi.data_ptr(), faces.data_ptr(), vertices.data_ptr(), int(phi.shape[0]), int(phi.shape[1]),
                                 int(faces.shape[0]), int(vertices.shape[1]), int(phi.dtype == torch.float64))
    if rc != 0:
        raise RuntimeError("reference sdf_cuda failed: cuda error %d" % rc)
    return phi


def grid(faces: torch.Tensor, verts_norm: torch.Tensor, grid_size: int, as_written: bool) -> torch.Tensor:
    """phi [B,G,G,G] from the reference kernel; as_written = faces handed over as [1,F,3] (fitting.py:367)."""
    f = faces.to(torch.int32).reshape(1, -1, 3).contiguous() if as_written else faces.to(torch.int32).reshape(-1, 3).contiguous()
    v = verts_norm.contiguous()
    phi = torch.zeros(v.shape[0], grid_size, grid_size, grid_size, dtype=v.dtype, device=v.device)
    return sdf(phi, f, v)


def install_csrc_stub():
    """registers `sdf.csrc` (the compiled extension of the reference's sdf package) backed by libsdf_refcuda.so"""
    if "sdf.csrc" not in sys.modules:
        m = types.ModuleType("sdf.csrc")
        m.sdf = sdf
        sys.modules["sdf.csrc"] = m
    return sys.modules["sdf.csrc"]
