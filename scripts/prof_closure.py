"""workload for ncu: a few closures (sparse / dense+SDF, B=256 and B=1) and a short L-BFGS stage"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsmplfitting_b200 import synthetic as S
from mvsmplfitting_b200.context import FittingContext
model = S.make_model(0); gmm = S.make_gmm(7)
mode = sys.argv[1] if len(sys.argv) > 1 else "closure"
for B in ((256, 1) if mode == "closure" else (256,)):
    cams = S.make_cameras(8)
    fr = S.make_frames(model, cams, B, seed=1000)
    ctx = FittingContext(0)
    ctx.set_model(model); ctx.set_gmm_from_dict(gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    x = torch.tensor(S.pack_params(fr["init"]), device="cuda")
    w = dict(data_weight=500 / 1536, body_pose_weight=57.4, shape_weight=10.0, bending_prior_weight=3.17 * 57.4)
    if mode == "closure":
        for sdf in (False, True):
            ctx.set_loss(body_prior="gmm", interpenetration=sdf, coll_loss_weight=1000.0 if sdf else 0.0, **w)
            for _ in range(3):
                ctx.closure(x)
            torch.cuda.synchronize()
    elif mode == "resident":          # frame-resident regime: one launch runs the whole stage
        ctx.set_loss(body_prior="gmm", interpenetration=False, **w)
        ctx.lbfgs_run(x, ctx.make_lbfgs_config(max_outer=2))
        torch.cuda.synchronize()
    else:
        ctx.set_loss(body_prior="gmm", interpenetration=True, coll_loss_weight=1000.0, **w)
        ctx.lbfgs_run(x, ctx.make_lbfgs_config(max_outer=2))
        torch.cuda.synchronize()
    ctx.close()
