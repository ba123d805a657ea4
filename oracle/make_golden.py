"""TEST INFRASTRUCTURE -- writes tests/golden/*.npz by RUNNING THE UNMODIFIED
REFERENCE (/root/reference) on the deterministic synthetic inputs of
mvsmplfitting_b200/synthetic.py.  Run in the authoring container only:

    python -m oracle.make_golden

The reference has no tests or golden vectors of its own, so these
reference-run outputs are what pins the oracle (and through it the CUDA path).
Fixtures hold inputs that cannot be regenerated (demo cameras / keypoints shipped
with the reference) and the reference's outputs; the synthetic model itself is
regenerated from its seed and guarded by a checksum.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from oracle import ref_harness as H            # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def model_checksum(model: dict) -> str:
    h = hashlib.sha256()
    for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights", "lsp_regressor"):
        h.update(np.ascontiguousarray(model[k]).tobytes())
    return h.hexdigest()


def stage_weights(stage: int, H_img: int = 1536) -> dict:
    """non_linear_solver.py:148-180 with cfg_files/fit_smpl.yaml:40-59."""
    sw = S.STAGE_WEIGHTS
    bpw = sw["body_pose_prior_weights"][stage]
    return dict(data_weight=500.0 / H_img, body_pose_weight=bpw, shape_weight=sw["shape_weights"][stage],
                bending_prior_weight=3.17 * bpw)


def load_demo_scene():
    """The six demo cameras / keypoint files shipped with the reference
    (data/3DOH50K_Parameters.txt, data/keypoints/0000/Camera0*/00001_keypoints.json)."""
    ns = H.import_reference()
    extris, intris = ns.utils.load_camera_para(os.path.join(H.REF_ROOT, "data", "3DOH50K_Parameters.txt"))
    trans, rot = ns.utils.get_rot_trans(extris, photoscan=False)
    V = len(extris)
    cams = dict(
        R=np.stack(rot).astype(np.float32), t=np.stack(trans).astype(np.float32),
        f=np.stack([[intris[v][0][0], intris[v][0][0]] for v in range(V)]).astype(np.float32),   # init.py:111-118
        c=np.stack([intris[v][:2, 2] for v in range(V)]).astype(np.float32), H=1536, W=2048)
    kps = []
    for v in range(V):
        fn = os.path.join(H.REF_ROOT, "data", "keypoints", "0000", "Camera%02d" % v, "00001_keypoints.json")
        with open(fn) as f:
            d = json.load(f)
        kps.append(np.array(d["people"][0]["pose_keypoints_2d"], dtype=np.float32).reshape(-1, 3)[:17])
    kps = np.stack(kps)                                       # [V,17,3]
    return cams, kps


def triangulate_mean(cams, kps):
    """rough 3-D point seen at the mean keypoint of every view (linear least squares)"""
    A, b = [], []
    for v in range(cams["R"].shape[0]):
        uv = (kps[v, :, :2] * kps[v, :, 2:3]).sum(0) / kps[v, :, 2].sum()
        ray = np.array([(uv[0] - cams["c"][v, 0]) / cams["f"][v, 0], (uv[1] - cams["c"][v, 1]) / cams["f"][v, 1], 1.0])
        R, t = cams["R"][v].astype(np.float64), cams["t"][v].astype(np.float64)
        # (I - rr^T)(R X + t) = 0
        r = ray / np.linalg.norm(ray)
        P = np.eye(3) - np.outer(r, r)
        A.append(P @ R)
        b.append(-P @ t)
    X = np.linalg.lstsq(np.concatenate(A), np.concatenate(b), rcond=None)[0]
    return X


def run_case(name, model, cams, frames, weights, body_prior_kind, gmm, params, model_type="smpllsp",
             use_joints_conf=True, fix_shape=False):
    out = {}
    ns = H.import_reference()
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        rm = H.build_reference_model(model, dtype=dt, model_type=model_type)
        if fix_shape:
            rm.betas.requires_grad = False
        rc = H.build_reference_cameras(cams, dtype=dt)
        bp = H.build_reference_gmm(gmm, dtype=dt) if body_prior_kind == "gmm" else ns.prior.create_prior("l2")
        B = params["betas"].shape[0]
        res = [H.reference_closure_eval(rm, rc, frames, b, weights, bp, dtype=dt, params=params,
                                        use_joints_conf=use_joints_conf, fix_shape=fix_shape) for b in range(B)]
        out["loss_" + tag] = np.array([r["loss"] for r in res], dtype=np.float64)
        g = np.zeros((B, 86))
        for b, r in enumerate(res):
            o = 0
            for k, n in zip(S.PARAM_ORDER, S.PARAM_SIZES):
                if k in r["grads"]:
                    g[b, o:o + n] = r["grads"][k]
                o += n
        out["grad_" + tag] = g
        out["joints_" + tag] = np.stack([r["joints"] for r in res]).astype(np.float64)
        out["proj_" + tag] = np.stack([r["proj"] for r in res], axis=1).astype(np.float64)      # [V,B,17,2]
        verts = np.stack([r["vertices"] for r in res]).astype(np.float64)
        out["verts_head_" + tag] = verts[:, :64]
        out["verts_sum_" + tag] = verts.sum(axis=1)
        out["verts_face_" + tag] = verts[:, S.FACE_VERTEX_IDS]
    out["X"] = S.pack_params(params)
    out["gt_uv"] = frames["gt_uv"]
    out["conf"] = frames["conf"]
    out["joint_weights"] = frames["joint_weights"]
    for k in ("R", "t", "f", "c"):
        out["cam_" + k] = cams[k]
    out["weights"] = np.array([weights[k] for k in ("data_weight", "body_pose_weight", "shape_weight",
                                                    "bending_prior_weight")], dtype=np.float64)
    out["meta"] = np.array(json.dumps(dict(name=name, body_prior=body_prior_kind, model_type=model_type,
                                           use_joints_conf=use_joints_conf, fix_shape=fix_shape)))
    print("  case %-12s loss_f32 %s" % (name, out["loss_f32"]))
    return out


def run_lbfgs_case(model, cams, frames, gmm, frame, stage, start="init"):
    """Reference LBFGSLs + FittingMonitor.run_fitting for one frame, one stage."""
    ns = H.import_reference()
    dt = torch.float32
    rm = H.build_reference_model(model, dtype=dt)
    rc = H.build_reference_cameras(cams, dtype=dt)
    gp = H.build_reference_gmm(gmm, dtype=dt)
    w = stage_weights(stage)
    V = frames["gt_uv"].shape[0]
    gt = torch.tensor(frames["gt_uv"][:, frame:frame + 1])
    conf = [torch.tensor(frames["conf"][v, frame:frame + 1]) for v in range(V)]
    jw = torch.tensor(frames["joint_weights"]).unsqueeze(0)
    loss = ns.fitting.create_loss("smplify", rho=100.0, use_joints_conf=True, dtype=dt, body_pose_prior=gp,
                                  shape_prior=ns.prior.create_prior("l2"),
                                  angle_prior=ns.prior.create_prior("angle", dtype=dt),
                                  interpenetration=False, fix_shape=False)
    loss.reset_loss_weights({k: torch.tensor(v, dtype=dt) for k, v in w.items()})
    mon = ns.fitting.FittingMonitor(maxiters=30, ftol=1e-9, gtol=1e-9)
    params = [p for p in rm.parameters() if p.requires_grad]
    H.set_model_params(rm, frames[start], frame)
    opt, _ = ns.optim_factory.create_optimizer(params, optim_type="lbfgsls", lr=1.0, maxiters=30)
    closure = mon.create_fitting_closure(opt, rm, camera=rc, gt_joints=gt, joints_conf=conf, joint_weights=jw,
                                         loss=loss, create_graph=False, use_vposer=False, vposer=None,
                                         pose_embedding=None, return_verts=True, return_full_pose=True,
                                         use_3d=False)
    trace, xs = [], []

    def traced():
        xs.append(np.concatenate([p.detach().numpy().reshape(-1) for p in params]))
        l = closure()
        trace.append(float(l))
        return l
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        final = mon.run_fitting(opt, traced, params, rm, use_vposer=False)
    x_final = np.concatenate([p.detach().numpy().reshape(-1) for p in params])
    return dict(trace=np.array(trace), eval_x=np.stack(xs), x_final=x_final, final=np.float64(final),
                n_iter=np.int64(opt.state[params[0]]["n_iter"]), x0=S.pack_params(frames[start])[frame],
                weights=np.array([w[k] for k in ("data_weight", "body_pose_weight", "shape_weight",
                                                 "bending_prior_weight")]),
                frame=np.int64(frame), stage=np.int64(stage))


def main():
    os.makedirs(GOLD, exist_ok=True)
    model = S.make_model(0)
    chk = model_checksum(model)
    gmm = S.make_gmm(7)
    cases = {}

    cams8 = S.make_cameras(8)
    fr8 = S.make_frames(model, cams8, 3, seed=3)
    cases["gmm8_s3"] = run_case("gmm8_s3", model, cams8, fr8, stage_weights(3), "gmm", gmm, fr8["init"])
    cases["gmm8_s2"] = run_case("gmm8_s2", model, cams8, fr8, stage_weights(2), "gmm", gmm, fr8["init"])
    # perturbed ground truth, non-unit scale
    rng = np.random.RandomState(11)
    pert = {k: (v + rng.normal(0, 0.02, size=v.shape)).astype(np.float32) for k, v in fr8["gt"].items()}
    pert["scale"] = (1.0 + rng.uniform(-0.2, 0.3, size=pert["scale"].shape)).astype(np.float32)
    cases["gmm8_gt"] = run_case("gmm8_gt", model, cams8, fr8, stage_weights(3), "gmm", gmm, pert)

    cams4 = S.make_cameras(4)
    fr4 = S.make_frames(model, cams4, 2, seed=4)
    cases["l2_4_s0"] = run_case("l2_4_s0", model, cams4, fr4, stage_weights(0), "l2", gmm, fr4["init"])
    cases["l2_4_s3"] = run_case("l2_4_s3", model, cams4, fr4, stage_weights(3), "l2", gmm, fr4["init"])
    cases["l2_4_noconf"] = run_case("l2_4_noconf", model, cams4, fr4, stage_weights(3), "l2", gmm, fr4["init"],
                                    use_joints_conf=False)
    cases["l2_4_fixshape"] = run_case("l2_4_fixshape", model, cams4, fr4, stage_weights(3), "l2", gmm, fr4["init"],
                                      fix_shape=True)

    # 'smpl' model type: chain joints + face vertices, hips un-weighted (data_parser.py:354-356)
    fr_smpl = S.make_frames(model, cams4, 2, seed=5, model_type="smpl")
    fr_smpl["joint_weights"][11] = 0.0
    fr_smpl["joint_weights"][12] = 0.0
    cases["smpl_coco"] = run_case("smpl_coco", model, cams4, fr_smpl, stage_weights(3), "gmm", gmm, fr_smpl["init"],
                                  model_type="smpl")

    # demo scene shipped with the reference (6 cameras, 1 frame) on the synthetic model
    dcams, kps = load_demo_scene()
    X = triangulate_mean(dcams, kps)
    demo_frames = dict(gt_uv=kps[:, None, :, :2].copy(), conf=kps[:, None, :, 2].copy(),
                       joint_weights=np.ones(17, dtype=np.float32))
    rng = np.random.RandomState(5)
    demo_params = dict(betas=rng.normal(0, 0.5, (1, 10)).astype(np.float32),
                       global_orient=np.array([[0.2, 2.8, -0.3]], dtype=np.float32),
                       body_pose=rng.normal(0, 0.15, (1, 69)).astype(np.float32),
                       transl=X[None].astype(np.float32),
                       scale=np.array([[3.5]], dtype=np.float32))
    cases["demo6"] = run_case("demo6", model, dcams, demo_frames, stage_weights(1), "l2", gmm, demo_params)

    for name, c in cases.items():
        np.savez_compressed(os.path.join(GOLD, "closure_%s.npz" % name), model_checksum=np.array(chk), **c)

    # optimiser trajectories
    t0 = run_lbfgs_case(model, cams8, fr8, gmm, frame=0, stage=3)
    t1 = run_lbfgs_case(model, cams8, fr8, gmm, frame=1, stage=0)
    np.savez_compressed(os.path.join(GOLD, "lbfgs_traj_s3.npz"), model_checksum=np.array(chk),
                        gt_uv=fr8["gt_uv"], conf=fr8["conf"], **t0)
    np.savez_compressed(os.path.join(GOLD, "lbfgs_traj_s0.npz"), model_checksum=np.array(chk),
                        gt_uv=fr8["gt_uv"], conf=fr8["conf"], **t1)
    print("lbfgs traj: evals", len(t0["trace"]), len(t1["trace"]), "n_iter", t0["n_iter"], t1["n_iter"])
    with open(os.path.join(GOLD, "MODEL_CHECKSUM.txt"), "w") as f:
        f.write(chk + "\n")
    print("model checksum", chk)


if __name__ == "__main__":
    main()
