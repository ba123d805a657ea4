"""closure time of the all-faces SDF modes through the dense-round kernels: python scripts/n3_time.py [B_lists=256] [B_brute=32]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from mvsmplfitting_b200.context import FittingContext  # noqa: E402

model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(8)
w = dict(data_weight=500 / 1536, body_pose_weight=57.4, shape_weight=10.0, bending_prior_weight=3.17 * 57.4)
for mode, B in ((0, int(sys.argv[1]) if len(sys.argv) > 1 else 256), (1, int(sys.argv[1]) if len(sys.argv) > 1 else 256),
                (2, int(sys.argv[2]) if len(sys.argv) > 2 else 32)):
    fr = S.make_frames(model, cams, B, seed=1000)
    ctx = FittingContext(0)
    ctx.set_model(model); ctx.set_gmm_from_dict(gmm); ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_exec_mode(3)
    ctx.set_loss(body_prior="gmm", interpenetration=True, coll_loss_weight=0.05 if mode else 1000.0, sdf_grid=128, sdf_all_faces=mode, **w)
    x = torch.tensor(S.pack_params(fr["init"]), device="cuda")
    ctx.closure(x)
    torch.cuda.synchronize()
    n = 3
    t0 = time.time()
    for _ in range(n):
        ctx.closure(x)
    torch.cuda.synchronize()
    print("sdf_all_faces=%d  B=%d  G=128: %.2f ms per closure (dense-round kernels), %.3f ms per frame" % (
        mode, B, (time.time() - t0) * 1e3 / n, (time.time() - t0) * 1e3 / n / B), flush=True)
    ctx.close()
