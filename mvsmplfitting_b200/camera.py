"""Fixed-extrinsics perspective camera: drop-in for reference code/camera.py (create_camera :34-38,
PerspectiveCamera :41-117).  Inside the fitting closure the projection, its GMoF residual and the
adjoint run in the CUDA keypoint kernel; this module is the parameter container the caller builds in
init.py:108-131, plus a standalone forward() for visualisation code."""
from __future__ import annotations

import torch
import torch.nn as nn


def create_camera(camera_type="persp", **kwargs):
    if camera_type.lower() == "persp":
        return PerspectiveCamera(**kwargs)
    raise ValueError("Uknown camera type: {}".format(camera_type))


class PerspectiveCamera(nn.Module):
    FOCAL_LENGTH = 5000

    def __init__(self, rotation=None, translation=None, focal_length_x=None, focal_length_y=None, batch_size=1,
                 center=None, dtype=torch.float32, **kwargs):
        super().__init__()
        self.batch_size = batch_size
        self.dtype = dtype

        def focal(v):
            if v is None or type(v) == float:
                return torch.full([batch_size], self.FOCAL_LENGTH if v is None else v, dtype=dtype)
            return v
        self.register_buffer("focal_length_x", focal(focal_length_x))
        self.register_buffer("focal_length_y", focal(focal_length_y))
        self.register_buffer("center", torch.zeros([batch_size, 2], dtype=dtype) if center is None else center)
        if rotation is None:
            rotation = torch.eye(3, dtype=dtype).unsqueeze(0).repeat(batch_size, 1, 1)
        self.register_parameter("rotation", nn.Parameter(rotation, requires_grad=True))
        if translation is None:
            translation = torch.zeros([batch_size, 3], dtype=dtype)
        self.register_parameter("translation", nn.Parameter(translation, requires_grad=True))

    def numpy_params(self):
        """(R[3,3], t[3], f[2], c[2]) of camera 0 for mvs_set_cameras"""
        R = self.rotation.detach()[0].cpu().numpy()
        t = self.translation.detach()[0].cpu().numpy()
        f = [float(self.focal_length_x.reshape(-1)[0]), float(self.focal_length_y.reshape(-1)[0])]
        c = self.center.detach().reshape(-1, 2)[0].cpu().numpy()
        return R, t, f, c

    def forward(self, points):
        """points [B,N,3] -> [B,N,2]; x = R p + t, uv = f * x.xy / x.z + c  (camera.py:93-117).
        Off the hot path (overlays / debugging); the closure never calls it."""
        x = torch.einsum("bki,bji->bjk", self.rotation, points) + self.translation.unsqueeze(1)
        uv = x[..., :2] / x[..., 2:3]
        f = torch.stack([self.focal_length_x, self.focal_length_y], dim=-1).unsqueeze(1)
        return uv * f + self.center.unsqueeze(1)
