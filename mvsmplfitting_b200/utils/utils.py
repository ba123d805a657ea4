"""The few helpers of reference code/utils/utils.py that the fitting path touches:
JointMapper (:411-424), GMoF (:427-438), rel_change (:348-349), smpl_to_annotation (:441-466)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def rel_change(prev_val, curr_val):
    return (prev_val - curr_val) / max([np.abs(prev_val), np.abs(curr_val), 1])


class JointMapper(nn.Module):
    def __init__(self, joint_maps=None):
        super().__init__()
        if joint_maps is None:
            self.joint_maps = joint_maps
        else:
            self.register_buffer("joint_maps", torch.tensor(np.asarray(joint_maps), dtype=torch.long))

    def forward(self, joints, **kwargs):
        return joints if self.joint_maps is None else torch.index_select(joints, 1, self.joint_maps)


class GMoF(nn.Module):
    def __init__(self, rho=1):
        super().__init__()
        self.rho = rho

    def extra_repr(self):
        return "rho = {}".format(self.rho)

    def forward(self, residual):
        sq = residual ** 2
        return self.rho ** 2 * torch.div(sq, sq + self.rho ** 2)


def smpl_to_annotation(model_type="smplx", use_hands=False, use_face=False, use_face_contour=False, pose_format="lsp14"):
    from .. import synthetic as S
    if pose_format == "coco17" and model_type == "smpl":
        return S.JOINT_MAP_COCO17_SMPL.copy()
    if pose_format == "lsp14" and model_type == "smpllsp":
        return S.JOINT_MAP_LSP14.copy()
    raise ValueError("Unknown model type / joint format: {} / {}".format(model_type, pose_format))
