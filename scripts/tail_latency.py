"""per-round latency of the dense regime with very few frames (the straggler tail): python scripts/tail_latency.py [B]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mvsmplfitting_b200 import synthetic as S
from mvsmplfitting_b200.context import FittingContext
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model = S.make_model(0); gmm = S.make_gmm(7); cams = S.make_cameras(8)
fr = S.make_frames(model, cams, B, seed=1000)
ctx = FittingContext(0)
ctx.set_model(model); ctx.set_gmm_from_dict(gmm)
ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
st = bench.stage_table()
x0 = torch.tensor(S.pack_params(fr["init"]), device="cuda")
cfgs = [ctx.make_loss_config(body_prior="gmm", interpenetration=True, sdf_grid=128, **s) for s in st]
names = [ctx.lib.mvs_kernel_name(k).decode() for k in range(18)]
for rep in range(3):
    x = x0.clone()
    ctx.fit(x, cfgs[:2])                      # stages 0-1 (frame-resident)
    torch.cuda.synchronize()
    prof = rep == 2
    if prof:
        ctx.profile(0xFFFFFFFF)
    t0 = time.time()
    _, s = ctx.fit(x, cfgs[2:])               # dense regime, stages 2-3 merged
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("B=%d  dense rounds %d  evals %d  %.1f us / round (%s)" % (B, s["rounds"], s["frame_evals"], dt * 1e6 / max(s["rounds"], 1),
                                                                     "events around every launch" if prof else "plain"))
    if prof:
        p = ctx.profile_read(); ctx.profile(0)
        tot = 0.0
        for k, (ms, n) in p.items():
            if n:
                print("   %-18s %6d launches  %7.2f us avg" % (k, n, ms * 1e3 / n)); tot += ms
        print("   sum of kernel time %.1f us / round" % (tot * 1e3 / max(s["rounds"], 1)))
