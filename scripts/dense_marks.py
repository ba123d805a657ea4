"""per-mark cycle breakdown of the persistent dense-round kernel (CTA 0, thread 0) from the -DMVS_PHASE_DBG build:
   python -m mvsmplfitting_b200.build --debug && python scripts/dense_marks.py [B]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from mvsmplfitting_b200 import _lib, synthetic as S  # noqa: E402
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libmvsmpl_dbg.so")
from mvsmplfitting_b200.context import FittingContext  # noqa: E402
import bench  # noqa: E402

NAMES = {0: "frame: entry (scalars loaded)", 1: "frame: prologue (state, history issue, SDF scalars)", 2: "P1 rodrigues/J", 3: "P2 chain",
         4: "P3 A/Phi", 5: "P4 vposed gather", 6: "P5", 7: "P6 keypoints", 8: "P6 projection", 9: "P6 reduce", 10: "P6 scatter",
         11: "P7 dvp/dA", 12: "P8 dPhi", 14: "P9 skin_transform_bwd", 15: "reverse sweep", 16: "P10a",
         17: "P10 rodrigues_bwd/betas + gm_diff", 18: "GMM matvec", 19: "GMM ll", 20: "priors", 21: "angle", 22: "total/grad",
         23: "g_eval store + cp.async wait", 24: "L-BFGS advance + write-back", 25: "slot hand-out + x copy", 26: "rodrigues (next)",
         27: "chain (next)", 28: "A/Phi store (next)", 39: "round top (na read)", 40: "GEMM phase work (CTA 0)", 41: "GEMM barrier wait",
         42: "skin work", 43: "skin barrier wait", 44: "sdf work", 45: "sdf barrier wait", 46: "frame work (incl. marks 0-28)", 47: "frame barrier wait"}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(8)
fr = S.make_frames(model, cams, B, seed=1000)
ctx = FittingContext(0)
ctx.set_model(model); ctx.set_gmm_from_dict(gmm); ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
stages = [ctx.make_loss_config(body_prior="gmm", interpenetration=True, sdf_grid=128, **st) for st in bench.stage_table()]
lib = ctypes.CDLL(_lib.LIB_PATH)
clk, cnt = (ctypes.c_longlong * 64)(), (ctypes.c_longlong * 64)()
x0 = torch.tensor(S.pack_params(fr["init"]), device="cuda")
for rep in range(2):
    x = x0.clone()
    ctx.fit(x, stages)
    torch.cuda.synchronize()
    assert lib.mvs_debug_dense_clocks(clk, cnt, 64) == 0          # also resets
ph = ctx.dense_phase_times()
ghz = 1.965
rounds = max(cnt[41], 1)
print("B = %d, rounds %d (CTA 0 marks per round; us at %.3f GHz)" % (B, rounds, ghz))
tot = 0.0
for i in range(64):
    if cnt[i]:
        us = clk[i] / ghz / 1e3 / rounds
        tot += us if i >= 39 else 0.0
        print("%2d %-52s n=%6d  %8.2f us/round  (%.2f us per hit)" % (i, NAMES.get(i, ""), cnt[i], us, clk[i] / ghz / 1e3 / cnt[i]))
print("sum of the round-level marks (39..47): %.2f us/round; kernel phases (ms):" % tot, ph)
