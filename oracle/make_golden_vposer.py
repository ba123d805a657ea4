"""TEST INFRASTRUCTURE -- writes tests/golden/vposer_s11.npz by RUNNING THE UNMODIFIED REFERENCE VPoser class
(/root/reference/code/model/VPoser.py) and the reference's fitting closure with use_vposer=True on the deterministic
synthetic decoder weights of mvsmplfitting_b200/synthetic.make_vposer.  Authoring container only:

    python -m oracle.make_golden_vposer
"""
from __future__ import annotations

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
from oracle.make_golden import stage_weights   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def reference_vposer(w: dict, dtype):
    H.import_reference()
    with H.in_reference_dir():
        from model.VPoser import VPoser
    vp = VPoser(num_neurons=512, latentD=32, data_shape=[1, 23, 3], use_cont_repr=True)
    sd = vp.state_dict()
    for ref_name, key in (("bodyprior_dec_fc1", "fc1"), ("bodyprior_dec_fc2", "fc2"), ("bodyprior_dec_out", "out")):
        sd[ref_name + ".weight"] = torch.tensor(w[key + "_w"])
        sd[ref_name + ".bias"] = torch.tensor(w[key + "_b"])
    vp.load_state_dict(sd)
    vp = vp.to(dtype)
    vp.eval()                                   # init.py:169
    return vp


def closure_with_vposer(model, cams, fr, frame, weights, vp, z, dtype):
    """fitting.py:162-203 with use_vposer=True: body_pose = vposer.decode(pose_embedding, 'aa')"""
    ns = H.import_reference()
    rm = H.build_reference_model(model, batch_size=1, dtype=dtype)
    rc = H.build_reference_cameras(cams, dtype=dtype)
    H.set_model_params(rm, fr["init"], frame)
    V = fr["gt_uv"].shape[0]
    gt = torch.tensor(fr["gt_uv"][:, frame:frame + 1], dtype=dtype)
    conf = [torch.tensor(fr["conf"][v, frame:frame + 1], dtype=dtype) for v in range(V)]
    jw = torch.tensor(fr["joint_weights"], dtype=dtype).unsqueeze(0)
    loss = ns.fitting.create_loss("smplify", rho=100.0, use_joints_conf=True, dtype=dtype,
                                  body_pose_prior=ns.prior.create_prior("l2"), shape_prior=ns.prior.create_prior("l2"),
                                  angle_prior=ns.prior.create_prior("angle", dtype=dtype), interpenetration=False,
                                  fix_shape=False)
    loss.reset_loss_weights({k: torch.tensor(v, dtype=dtype) for k, v in weights.items()})
    emb = torch.tensor(z, dtype=dtype).view(1, 32).requires_grad_(True)
    rm.body_pose.requires_grad = False
    plist = [q for q in rm.parameters() if q.requires_grad] + [emb]
    opt = torch.optim.SGD(plist, lr=0.0)
    mon = ns.fitting.FittingMonitor(maxiters=30, ftol=1e-9, gtol=1e-9)
    closure = mon.create_fitting_closure(opt, rm, camera=rc, gt_joints=gt, joints_conf=conf, joint_weights=jw, loss=loss,
                                         create_graph=False, use_vposer=True, vposer=vp, pose_embedding=emb,
                                         return_verts=True, return_full_pose=True, use_3d=False)
    total = closure()
    bp = vp.decode(emb, output_type="aa").view(1, -1)
    out = rm(return_verts=True, body_pose=bp, return_full_pose=True)
    g = {k: getattr(rm, k).grad.detach().numpy().reshape(-1).copy() for k in ("betas", "global_orient", "transl", "scale")}
    g["pose_embedding"] = emb.grad.detach().numpy().reshape(-1).copy()
    return float(total), g, out.joints.detach().numpy()[0].copy(), bp.detach().numpy()[0].copy()


def write_fixture(name, w, vp_for, n_codes=24):
    rng = np.random.RandomState(5)
    # latent codes: typical (unit normal) and large ones that push rotations past 90 / 180 degrees (all four
    # quaternion branches of rotation_matrix_to_quaternion)
    Z = np.concatenate([rng.normal(0, 1.0, size=(n_codes, 32)), rng.normal(0, 6.0, size=(n_codes, 32))]).astype(np.float32)
    Z[0] = 0.0                                  # decode(0): the known answer of SURVEY 8c for the shipped snapshot (max |aa| = 0.8225)
    out = dict(Z=Z)
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        vp = vp_for(dtype)
        with torch.no_grad():
            aa = vp.decode(torch.tensor(Z, dtype=dtype), output_type="aa").reshape(Z.shape[0], -1).numpy()
        out["aa_" + tag] = aa
    # closure cases: 2 frames x 4 views, stage-3 and stage-0 weights
    model = S.make_model(0)
    cams = S.make_cameras(4)
    fr = S.make_frames(model, cams, 2, seed=31)
    zc = rng.normal(0, 0.8, size=(2, 32)).astype(np.float32)
    out.update(closure_Z=zc, gt_uv=fr["gt_uv"], conf=fr["conf"], joint_weights=fr["joint_weights"],
               X=S.pack_params(fr["init"]), cam_R=cams["R"], cam_t=cams["t"], cam_f=cams["f"], cam_c=cams["c"])
    for stage in (3, 0):
        wts = stage_weights(stage)
        out["w%d" % stage] = np.array([wts["data_weight"], wts["body_pose_weight"], wts["shape_weight"], wts["bending_prior_weight"]])
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            vp = vp_for(dtype)
            for b in range(2):
                total, g, joints, bp = closure_with_vposer(model, cams, fr, b, wts, vp, zc[b], dtype)
                pre = "s%d_b%d_%s_" % (stage, b, tag)
                out[pre + "loss"] = np.array(total)
                out[pre + "joints"] = joints
                out[pre + "body_pose"] = bp
                for k, v in g.items():
                    out[pre + "g_" + k] = v
    np.savez_compressed(os.path.join(GOLD, name), **out)
    print("wrote %s:" % name, {k: np.asarray(v).shape for k, v in out.items() if not k.startswith("s")})
    print("stage-3 frame-0 loss f32/f64:", out["s3_b0_f32_loss"], out["s3_b0_f64_loss"], " max|decode(0)| =", np.abs(out["aa_f32"][0]).max())


def main():
    assert H.available(), "reference tree missing"
    w = S.make_vposer(11)
    write_fixture("vposer_s11.npz", w, lambda dtype: reference_vposer(w, dtype))
    # the snapshot the reference ships (priors/snapshots/poser_epoch091.pkl) through its own loader; the fixture holds inputs
    # and outputs only -- the weights stay in the reference tree (tests read them from there / from oracle/_ref/reference)
    real = H.load_reference_vposer()
    wr = H.vposer_weights_numpy(real)
    write_fixture("vposer_real.npz", wr, lambda dtype: reference_vposer(wr, dtype), n_codes=12)


if __name__ == "__main__":
    main()
