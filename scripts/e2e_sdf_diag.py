"""GPU diagnostic for the end-to-end comparison with the SDF term: per frame, the reference's (torch-CUDA) per-stage counts and
final loss, the device's final loss in exec modes 0 (dense-regime kernels) and 1 (batched fp32 chain), and the device closure
evaluated at the REFERENCE's final parameters (do the two agree on the objective there?)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from mvsmplfitting_b200.context import FittingContext  # noqa: E402
from oracle import ref_fit as RF  # noqa: E402
from tests.test_gpu_fit_e2e import stage_cfgs  # noqa: E402

B, V = 8, 8
model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(V)
fr = S.make_frames(model, cams, B, seed=4200)
X0 = S.pack_params(fr["init"])
sc = RF.build_scene(model, gmm, cams, device="cuda")
runs = [RF.fit_frame(sc, fr, b, S.STAGE_WEIGHTS, interpenetration=True) for b in range(B)]
out = {"reference": [dict(per_stage=r["per_stage"], final=r["final_loss"]) for r in runs]}
Xref = np.stack([r["params"] for r in runs])
for mode in (0, 1):
    ctx = FittingContext(0)
    ctx.set_model(model); ctx.set_gmm_from_dict(gmm); ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_exec_mode(mode)
    cfgs = stage_cfgs(ctx, True)
    x = torch.tensor(X0, device="cuda")
    final, st = ctx.fit(x, cfgs)
    torch.cuda.synchronize()
    ctx.set_loss(config=cfgs[3])
    at_ref = ctx.closure(torch.tensor(Xref, device="cuda"), want_grad=False)["loss"].cpu().numpy()
    at_own = ctx.closure(x, want_grad=False)["loss"].cpu().numpy()
    out["device_mode%d" % mode] = dict(final=final.cpu().numpy().tolist(), stats=st, closure_at_reference_final=at_ref.tolist(),
                                       closure_at_own_final=at_own.tolist())
    ctx.close()
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "e2e_sdf_diag.json"), "w"), indent=1)
