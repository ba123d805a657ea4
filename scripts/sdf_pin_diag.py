"""GPU diagnostic: where does the closure-with-SDF gradient differ from the unmodified reference run?  Prints, per gradient
segment, the max-norm relative error of (a) the product (batched chain and dense-regime kernels) and (b) the reference's own
fp32 CUDA run against an fp64 evaluation of the same formulas (oracle glue in double, phi from the reference kernel)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from mvsmplfitting_b200.context import FittingContext  # noqa: E402
from oracle import closure_oracle as O, ref_harness as RH, ref_sdf, sdf_oracle  # noqa: E402
from oracle.lbfgs_oracle import PARAM_SEGMENTS  # noqa: E402

SEG = ("betas", "global_orient", "body_pose", "transl", "scale")
rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def main(B=6, grid=128, cw=1000.0, seed=2):
    model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(4)
    fr = S.make_frames(model, cams, B, seed=seed)
    w = dict(data_weight=500.0 / 1536, body_pose_weight=57.4, shape_weight=10.0, bending_prior_weight=3.17 * 57.4)
    X = S.pack_params(fr["init"])
    rm, rc, pr = RH.build_reference_model(model), RH.build_reference_cameras(cams), RH.build_reference_gmm(gmm)
    runs = [RH.reference_closure_eval(rm, rc, fr, b, w, pr, device="cuda", interpenetration=True, coll_loss_weight=cw) for b in range(B)]
    runs0 = [RH.reference_closure_eval(rm, rc, fr, b, w, pr, device="cuda", interpenetration=False) for b in range(B)]
    ref_loss = np.array([r["loss"] for r in runs]); ref_loss0 = np.array([r["loss"] for r in runs0])
    ref_grad = np.stack([np.concatenate([r["grads"][k] for k in SEG]) for r in runs])
    ref_grad0 = np.stack([np.concatenate([r["grads"][k] for k in SEG]) for r in runs0])

    def ref_grid(faces, vn, G_, all_faces=False):
        return ref_sdf.grid(torch.tensor(np.asarray(faces, dtype=np.int32), device="cuda"),
                            torch.tensor(np.asarray(vn, dtype=np.float32), device="cuda"), G_, as_written=not all_faces).cpu().numpy()
    sdf_oracle.GRID_FN = ref_grid
    out = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        om = O.OracleModel.from_numpy(model, dtype=dt)
        pri = O.OraclePriors.gmm_from_dict(gmm, dt)
        cfg = O.LossConfig(interpenetration=True, coll_loss_weight=cw, sdf_grid=grid, **w)
        out[name] = O.closure_eval_batch(om, cfg, pri, O.cams_to_torch(cams, dt), X, fr["gt_uv"], fr["conf"], fr["joint_weights"])
    a64 = out["f64"]
    res = {"pen_ref": (ref_loss - ref_loss0).tolist()}
    ours = {}
    for mode in (0, 3):
        ctx = FittingContext(0)
        ctx.set_model(model); ctx.set_gmm_from_dict(gmm); ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
        ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
        ctx.set_exec_mode(mode)
        ctx.set_loss(body_prior="gmm", interpenetration=True, coll_loss_weight=cw, sdf_grid=grid, **w)
        o = ctx.closure(torch.tensor(X, device="cuda"))
        torch.cuda.synchronize()
        ours[mode] = (o["loss"].cpu().numpy(), o["grad"].cpu().numpy())
        ctx.close()
    rows = {"reference_cuda_f32": (ref_loss, ref_grad), "oracle_f32_refphi": (out["f32"]["loss"], out["f32"]["grad"]),
            "ours_chain": ours[0], "ours_dense": ours[3]}
    for k, (l, g) in rows.items():
        res[k] = {"loss_vs_f64": rel(l, a64["loss"]), "loss_vs_refrun": rel(l, ref_loss),
                  "grad_vs_f64": {n: rel(g[:, a:e], a64["grad"][:, a:e]) for (a, e), n in zip(PARAM_SEGMENTS, SEG)},
                  "grad_vs_refrun": {n: rel(g[:, a:e], ref_grad[:, a:e]) for (a, e), n in zip(PARAM_SEGMENTS, SEG)}}
    # per-frame view of the worst segment
    a, e = PARAM_SEGMENTS[1]
    res["orient_per_frame"] = {"f64": a64["grad"][:, a:e].tolist(), "refrun": ref_grad[:, a:e].tolist(), "ours_dense": ours[3][1][:, a:e].tolist(),
                               "refrun_nosdf": ref_grad0[:, a:e].tolist()}
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sdf_pin_diag.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
