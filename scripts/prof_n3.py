"""workload for ncu: two closures in the all-faces SDF semantics over candidate lists (8 frames, G = 128) through the dense-round kernels"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from mvsmplfitting_b200.context import FittingContext  # noqa: E402

B = 8
model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(8)
fr = S.make_frames(model, cams, B, seed=1000)
ctx = FittingContext(0)
ctx.set_model(model); ctx.set_gmm_from_dict(gmm); ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
ctx.set_exec_mode(3)
ctx.set_loss(body_prior="gmm", interpenetration=True, coll_loss_weight=0.05, sdf_grid=128, sdf_all_faces=1, data_weight=500 / 1536,
             body_pose_weight=57.4, shape_weight=10.0, bending_prior_weight=3.17 * 57.4)
x = torch.tensor(S.pack_params(fr["init"]), device="cuda")
for _ in range(2):
    ctx.closure(x)
torch.cuda.synchronize()
ctx.close()
