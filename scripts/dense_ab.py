"""A/B check of the dense rounds: the persistent cooperative kernel (MVS_DENSE_PERSISTENT=1) against the four-launch rounds
(default).  Usage: python scripts/dense_ab.py OUT.npz [B]  -- run once per variant, then
python scripts/dense_ab.py --compare A.npz B.npz.  A frame's arithmetic is identical in both, so the parameters must agree bit for bit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(out, B):
    import torch
    import bench
    from mvsmplfitting_b200 import synthetic as S
    from mvsmplfitting_b200.context import FittingContext
    model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(8)
    fr = S.make_frames(model, cams, B, seed=1000)
    ctx = FittingContext(0)
    ctx.set_model(model); ctx.set_gmm_from_dict(gmm); ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    stages = [ctx.make_loss_config(body_prior="gmm", interpenetration=True, sdf_grid=128, **st) for st in bench.stage_table()]
    X0 = S.pack_params(fr["init"])
    x = torch.tensor(X0, device="cuda")
    final, st = ctx.fit(x, stages)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        x.copy_(torch.tensor(X0, device="cuda"))
        torch.cuda.synchronize()
        t0 = time.time()
        final, st = ctx.fit(x, stages)
        torch.cuda.synchronize()
        ts.append((time.time() - t0) * 1e3)
    print("variant", "persistent" if os.environ.get("MVS_DENSE_PERSISTENT") else "multikernel", "B", B, "ms", [round(t, 2) for t in ts], st,
          "phases", ctx.dense_phase_times())
    np.savez(out, x=x.cpu().numpy(), final=final.cpu().numpy(), it=st["frame_iterations"], ev=st["frame_evals"])


if __name__ == "__main__":
    if sys.argv[1] == "--compare":
        a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
        same = np.array_equal(a["x"], b["x"]) and np.array_equal(a["final"], b["final"])
        print("bit-identical:", same, "iterations", int(a["it"]), int(b["it"]), "evals", int(a["ev"]), int(b["ev"]),
              "max |dx|", float(np.abs(a["x"] - b["x"]).max()))
        sys.exit(0 if same else 1)
    run(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
