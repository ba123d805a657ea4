"""Times mvs_init_guess (3 launches: seed, rest-joint forward, init_guess_kernel) with CUDA events and checks it against
the oracle on a few frames.  First hardware run of that kernel: run under `timeout`.
    python scripts/init_time.py [frames] [views]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvsmplfitting_b200 import synthetic as S                 # noqa: E402
from mvsmplfitting_b200.context import FittingContext         # noqa: E402
from oracle import init_oracle as IO                          # noqa: E402  (checker only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
V = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model, cams = S.make_model(0), S.make_cameras(V)
fr = S.make_frames(model, cams, B, seed=5)
ctx = FittingContext(0)
ctx.set_model(model)
ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"])
ctx.set_batch(B)
ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
for _ in range(3):
    params, j3 = ctx.init_guess()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = ctx.launch_count()
e0.record()
for _ in range(20):
    params, j3 = ctx.init_guess()
e1.record()
torch.cuda.synchronize()
print("mvs_init_guess: %.1f us per call (%d frames x %d views), %d launches per call"
      % (e0.elapsed_time(e1) * 1e3 / 20, B, V, (ctx.launch_count() - n0) // 20))
z = lambda n: np.zeros((1, n))
rest = S.model_keypoints_np(model, z(10), z(3), z(69), z(3), np.ones((1, 1)), "smpllsp")[0]
ext = np.tile(np.eye(4), (V, 1, 1))
ext[:, :3, :3], ext[:, :3, 3] = cams["R"], cams["t"]
intr = np.tile(np.eye(3), (V, 1, 1))
intr[:, 0, 0], intr[:, 1, 1], intr[:, 0, 2], intr[:, 1, 2] = cams["f"][:, 0], cams["f"][:, 1], cams["c"][:, 0], cams["c"][:, 1]
x, j = params.cpu().numpy(), j3.cpu().numpy()
worst = 0.0
for b in range(0, B, max(1, B // 8)):
    kps = [np.concatenate([fr["gt_uv"][v, b], fr["conf"][v, b][:, None]], axis=1) for v in range(V)]
    o = IO.init_guess(ext, intr, kps, rest, True, 1.0, True)
    worst = max(worst, np.abs(j[b] - o["joints3d"]).max(), np.abs(x[b, 10:13] - o["global_orient"]).max(),
                np.abs(x[b, 82:85] - o["transl"]).max(), abs(x[b, 85] - o["scale"]))
print("max abs deviation from the oracle over the sampled frames: %.2e" % worst)
print("translation error of the guess vs ground truth (median over frames): %.3f"
      % np.median(np.abs(x[:, 82:85] - fr["gt"]["transl"]).max(axis=1)))
