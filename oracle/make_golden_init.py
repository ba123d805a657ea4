"""TEST INFRASTRUCTURE -- writes tests/golden/init_s21.npz by RUNNING THE UNMODIFIED REFERENCE functions
`recompute3D` (code/utils/recompute3D.py:24), `umeyama` (code/utils/umeyama.py:18) and cv2.Rodrigues
(init_guess.py:86) on seeded inputs.  Authoring container only:

    python -m oracle.make_golden_init
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from oracle import ref_harness as H            # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def rodrigues(r):
    th = np.linalg.norm(r)
    k = r / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def main():
    import cv2
    H.import_reference()
    with H.in_reference_dir():
        from utils.recompute3D import recompute3D
        from utils.umeyama import umeyama
    rng = np.random.RandomState(21)
    model = S.make_model(0)
    out = {}
    # --- triangulation: ring cameras of the benchmark (4, 8, 16 views), noisy detections, some dropped (conf 0)
    for V in (2, 4, 8, 16):
        cams = S.make_cameras(V)
        fr = S.make_frames(model, cams, 6, seed=300 + V)
        ext = np.tile(np.eye(4), (V, 1, 1))
        ext[:, :3, :3] = cams["R"]
        ext[:, :3, 3] = cams["t"]
        intr = np.tile(np.eye(3), (V, 1, 1))
        intr[:, 0, 0], intr[:, 1, 1] = cams["f"][:, 0], cams["f"][:, 1]
        intr[:, 0, 2], intr[:, 1, 2] = cams["c"][:, 0], cams["c"][:, 1]
        conf = fr["conf"].copy()
        drop = rng.rand(*conf.shape) < 0.1
        conf[drop] = 0.0
        j3 = []
        for b in range(6):
            kps = [np.concatenate([fr["gt_uv"][v, b], conf[v, b][:, None]], axis=1)[None].astype(np.float64) for v in range(V)]
            j3.append(recompute3D(ext, intr, kps))
        out["tri%d_ext" % V], out["tri%d_int" % V] = ext, intr
        out["tri%d_uv" % V], out["tri%d_conf" % V] = fr["gt_uv"], conf
        out["tri%d_j3" % V] = np.stack(j3)
    # --- similarity alignment (reference function, as written) on torso-like 4-point and 17-point sets
    src_l, dst_l, est_l, rot_l, tr_l, sc_l = [], [], [], [], [], []
    for k in range(24):
        n = 4 if k % 2 == 0 else 17
        src = rng.normal(size=(n, 3)) * [0.25, 0.45, 0.12]
        R = rodrigues(rng.normal(size=3) * (0.3 + 0.1 * k))
        s = 0.5 + rng.rand() * 2
        dst = s * (src @ R.T) + rng.normal(size=3) + rng.normal(size=(n, 3)) * 0.01
        est = bool(k % 3)
        rot, tr, sc = umeyama(src.copy(), dst.copy(), est)
        src_l.append(np.pad(src, ((0, 17 - n), (0, 0)))); dst_l.append(np.pad(dst, ((0, 17 - n), (0, 0))))
        est_l.append(est); rot_l.append(rot); tr_l.append(tr); sc_l.append(sc)
    out.update(um_n=np.array([4 if k % 2 == 0 else 17 for k in range(24)]), um_src=np.stack(src_l), um_dst=np.stack(dst_l),
               um_est=np.array(est_l), um_rot=np.stack(rot_l), um_trans=np.stack(tr_l), um_scale=np.array(sc_l, dtype=np.float64))
    # --- cv2.Rodrigues, matrix -> vector: generic, tiny angle, near pi
    Rs, rv = [], []
    for k in range(40):
        ang = [1e-7, 1e-3, 0.5, 2.0, 3.0, np.pi - 1e-4, np.pi - 1e-7, np.pi][k % 8]
        ax = rng.normal(size=3)
        R = rodrigues(ax / np.linalg.norm(ax) * ang)
        Rs.append(R)
        rv.append(cv2.Rodrigues(R)[0].reshape(3))
    out.update(rod_R=np.stack(Rs), rod_r=np.stack(rv))
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "init_s21.npz"), **out)
    print("wrote", os.path.join(GOLD, "init_s21.npz"), {k: v.shape for k, v in out.items()})
    whole_init_guess()


def whole_init_guess():
    """tests/golden/init_guess_ref.npz: the reference's own init_guess + fix_params (code/utils/init_guess.py:18-107,190-212),
    UNMODIFIED, on the synthetic model: multi-view (triangulation) and single-view (depth guess, :54-78) inputs.  The file
    hard-codes `.cuda()` at :38; Tensor.cuda is a no-op while it runs here."""
    import torch
    ns = H.import_reference()
    with H.in_reference_dir():
        from utils import init_guess as RIG
    model = S.make_model(0)
    rm = H.build_reference_model(model)
    out = {}
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        for tag, V, est in (("mv4", 4, True), ("sv", 1, False), ("sv_est", 1, True)):
            cams = S.make_cameras(max(V, 4))
            cams = {k: v[:V] for k, v in cams.items() if isinstance(v, np.ndarray)}
            fr = S.make_frames(model, cams, 4, seed=700 + V)
            ext = np.tile(np.eye(4), (V, 1, 1))
            ext[:, :3, :3], ext[:, :3, 3] = cams["R"], cams["t"]
            intr = np.tile(np.eye(3), (V, 1, 1))
            intr[:, 0, 0], intr[:, 1, 1], intr[:, 0, 2], intr[:, 1, 2] = cams["f"][:, 0], cams["f"][:, 1], cams["c"][:, 0], cams["c"][:, 1]
            res = []
            for b in range(4):
                kps = [np.concatenate([fr["gt_uv"][v, b], fr["conf"][v, b][:, None]], axis=1)[None].astype(np.float32) for v in range(V)]
                setting = dict(model=rm, dtype=torch.float32, batch_size=1, device=torch.device("cpu"), fix_scale=not est,
                               fixed_scale=None, extris=ext, intris=intr)
                data = {"keypoints": kps, "3d_joint": None}
                RIG.init_guess(setting, data, use_torso=True, model_type="smpllsp", use_vposer=False, use_3d=False)
                RIG.fix_params(setting, scale=None, shape=None)
                res.append(np.concatenate([rm.betas.detach().numpy().reshape(-1), rm.global_orient.detach().numpy().reshape(-1),
                                           rm.body_pose.detach().numpy().reshape(-1), rm.transl.detach().numpy().reshape(-1),
                                           rm.scale.detach().numpy().reshape(-1)]))
            out[tag + "_ext"], out[tag + "_int"] = ext, intr
            out[tag + "_uv"], out[tag + "_conf"] = fr["gt_uv"], fr["conf"]
            out[tag + "_params"] = np.stack(res)
    finally:
        torch.Tensor.cuda = cuda
    np.savez_compressed(os.path.join(GOLD, "init_guess_ref.npz"), **out)
    print("wrote init_guess_ref.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
