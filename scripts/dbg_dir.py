"""debug: first L-BFGS step of the dense regime (exec mode 0) against the batched reference chain (mode 1), per frame"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mvsmplfitting_b200 import synthetic as S
from mvsmplfitting_b200.context import FittingContext
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = S.make_model(0); gmm = S.make_gmm(7); cams = S.make_cameras(8)
fr = S.make_frames(model, cams, B, seed=1000)
X0 = S.pack_params(fr["init"])
# start from a point where penetration is active: run stages 0-1 first (mode 0, resident)
res = {}
xs = None
for mode in (0, 1):
    ctx = FittingContext(0)
    ctx.set_model(model); ctx.set_gmm_from_dict(gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    st = bench.stage_table()
    if xs is None:
        x = torch.tensor(X0, device="cuda")
        for i in (0, 1):
            ctx.set_loss(body_prior="gmm", interpenetration=False, **st[i]); ctx.lbfgs_run(x)
        xs = x.clone()
    ctx.set_exec_mode(mode)
    ctx.set_loss(body_prior="gmm", interpenetration=True, sdf_grid=128, **st[2])
    o = ctx.closure(xs.clone()) if mode == 1 else None
    x = xs.clone()
    final, stt = ctx.lbfgs_run(x, ctx.make_lbfgs_config(max_outer=1, max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 1))
    res[mode] = ((x - xs).cpu().numpy(), final.cpu().numpy(), stt, o)
    ctx.close()
d0, f0, s0, _ = res[0]; d1, f1, s1, o = res[1]
cos = (d0 * d1).sum(1) / (np.linalg.norm(d0, axis=1) * np.linalg.norm(d1, axis=1) + 1e-30)
rel = np.linalg.norm(d0 - d1, axis=1) / (np.linalg.norm(d1, axis=1) + 1e-30)
print("stats", s0, s1)
print("entry loss rel diff max", np.nanmax(np.abs(f0 - f1) / np.abs(f1)))
bad = np.argsort(-rel)[:10]
for b in bad:
    print("frame %3d cos %.6f rel %.3e |d0| %.3e |d1| %.3e loss %.5e / %.5e" % (b, cos[b], rel[b], np.linalg.norm(d0[b]), np.linalg.norm(d1[b]), f0[b], f1[b]))
print("frames with rel > 1e-2:", int((rel > 1e-2).sum()), "of", B)
