"""CPU: host-side sequence I/O: camera file, keypoint JSON, result pickle / obj, batch driver."""
import json
import os
import pickle
import sys

import numpy as np
import pytest

from mvsmplfitting_b200 import seqio

REF = "/root/reference"


def write_demo(tmp, V=3, B=4, missing=((1, 2),), seed=0):
    rng = np.random.default_rng(seed)
    kp = rng.uniform(0, 2000, size=(V, B, 19, 3)).astype(np.float32)        # 19 > 17: the reader keeps the first 17
    kp[..., 2] = rng.uniform(0, 1, size=(V, B, 19))
    for v in range(V):
        d = os.path.join(tmp, "keypoints", "0007", "Camera%02d" % v)
        os.makedirs(d)
        for b in range(B):
            if (v, b) in missing:
                continue
            people = [{"pose_keypoints_2d": kp[v, b].reshape(-1).tolist()}]
            with open(os.path.join(d, "%05d_keypoints.json" % (b + 1)), "w") as f:
                json.dump({"version": 1.1, "people": people}, f)
    ext = np.tile(np.eye(4), (V, 1, 1))
    ext[:, :3, :] = rng.normal(size=(V, 3, 4))
    intr = np.tile(np.eye(3), (V, 1, 1))
    intr[:, 0, 0] = intr[:, 1, 1] = rng.uniform(2000, 2500, V)
    intr[:, :2, 2] = rng.uniform(700, 1100, (V, 2))
    with open(os.path.join(tmp, "cams.txt"), "w") as f:
        for v in range(V):
            f.write("%d\n" % v)
            for r in intr[v]:
                f.write(" ".join(repr(float(x)) for x in r) + " \n")
            f.write("0 0\n")
            for r in ext[v, :3]:
                f.write(" ".join(repr(float(x)) for x in r) + " \n")
            f.write("\n")
    return kp, ext, intr


def test_camera_file_and_keypoint_folders(tmp_path):
    kp, ext, intr = write_demo(str(tmp_path))
    e, i = seqio.load_camera_para(str(tmp_path / "cams.txt"))
    assert np.array_equal(e, ext) and np.array_equal(i, intr)
    cams = seqio.camera_arrays(e, i, views=[0, 2])
    assert cams["R"].shape == (2, 3, 3) and cams["R"].dtype == np.float32
    assert np.allclose(cams["t"][1], ext[2, :3, 3]) and np.allclose(cams["f"][1], intr[2, 0, 0]) and np.allclose(cams["c"][0], intr[0, :2, 2])
    seq = seqio.load_sequence(str(tmp_path / "keypoints"), "0007")
    assert seq["cameras"] == ["Camera00", "Camera01", "Camera02"] and seq["frames"] == ["00001", "00002", "00003", "00004"]
    assert seq["gt_uv"].shape == (3, 4, 17, 2) and seq["conf"].shape == (3, 4, 17)
    assert not seq["present"][1, 2] and seq["present"].sum() == 11
    assert (seq["conf"][1, 2] == 0).all() and (seq["gt_uv"][1, 2] == 0).all()
    assert np.array_equal(seq["gt_uv"][2, 3], kp[2, 3, :17, :2]) and np.array_equal(seq["conf"][0, 0], kp[0, 0, :17, 2])
    sub = seqio.load_sequence(str(tmp_path / "keypoints"), "0007", cameras=["Camera02"], frames=["00002", "00009"])
    assert sub["gt_uv"].shape == (1, 2, 17, 2) and sub["present"].tolist() == [[True, False]]
    assert seqio.read_keypoints(str(tmp_path / "keypoints" / "0007" / "Camera00" / "00001_keypoints.json"), person=1) is None


def test_joint_weights():
    w = seqio.joint_weights("coco17", True)
    assert w[11] == 0 and w[12] == 0 and w.sum() == 15
    assert seqio.joint_weights("lsp14", True).sum() == 17 and seqio.joint_weights("lsp14", False).sum() == 15


def test_result_pickle_and_obj(tmp_path):
    x = np.arange(86, dtype=np.float32) + 1
    r = seqio.result_from_params(x, 12.5)
    assert [r[k].shape for k in ("betas", "global_orient", "body_pose", "transl", "scale")] == [(1, 10), (1, 3), (1, 69), (1, 3), (1,)]
    verts = np.random.default_rng(0).normal(size=(5, 3))
    faces = np.array([[0, 1, 2], [2, 3, 4]])
    out = seqio.save_results(str(tmp_path / "res"), "0007", "00001", r, verts=verts, faces=faces, mesh_folder=str(tmp_path / "mesh"))
    assert out == str(tmp_path / "res" / "0007" / "00001" / "000.pkl")
    with open(out, "rb") as f:
        raw = f.read()
    assert raw[:2] == b"\x80\x02"                                    # protocol 2, utils.py:862
    got = pickle.loads(raw)
    bp = got["body_pose"][0]
    assert (bp[18:24] == 0).all() and (bp[27:33] == 0).all() and (bp[57:] == 0).all() and (bp[:18] == x[13:31]).all()
    assert got["pose"].shape == (1, 72) and (got["pose"][0, :3] == x[10:13]).all() and got["loss"] == 12.5
    lines = open(tmp_path / "mesh" / "0007" / "00001" / "000.obj").read().splitlines()
    assert len(lines) == 7 and lines[5] == "f 1 2 3" and lines[6] == "f 3 4 5"
    assert np.allclose([float(t) for t in lines[0].split()[1:]], verts[0], atol=1e-7)


def test_fit_sequence_drives_the_context_in_batch_order(tmp_path):
    import torch
    write_demo(str(tmp_path))
    seq = seqio.load_sequence(str(tmp_path / "keypoints"), "0007")

    class StubCtx:                                  # records what the driver asks of a FittingContext
        B, calls = 4, []

        def set_keypoints(self, gt_uv, conf, jw):
            self.calls.append(("kp", gt_uv.shape, conf.shape, jw.sum()))

        def init_guess(self, **kw):
            self.calls.append(("init", kw["estimate_scale"], kw["use_torso"], kw["hip_seed"]))
            return torch.arange(4 * 86, dtype=torch.float32).reshape(4, 86), None

        def fit(self, params, stage_cfgs, opt_cfg):
            self.calls.append(("fit", len(stage_cfgs)))
            params += 1
            return torch.tensor([1.0, 2.0, 3.0, 4.0]), dict(frame_iterations=7)

        def forward_only(self, params, want_verts=True):
            self.calls.append(("fwd", float(params[0, 13 + 18]), float(params[0, 13 + 17])))
            return dict(verts=torch.zeros(4, 5, 3))

    ctx = StubCtx()
    x, loss, st = seqio.fit_sequence(ctx, seq, stage_cfgs=[1, 2, 3, 4], result_folder=str(tmp_path / "res"),
                                     mesh_folder=str(tmp_path / "mesh"), faces=np.array([[0, 1, 2]]))
    assert [c[0] for c in ctx.calls] == ["kp", "init", "fit", "fwd"]
    assert ctx.calls[0][1:] == ((3, 4, 17, 2), (3, 4, 17), 15.0) and ctx.calls[1][1:] == (False, True, 1.0)
    assert ctx.calls[3][1] == 0.0 and ctx.calls[3][2] == 13 + 17 + 1          # mesh from the saved (zeroed) pose
    assert x.shape == (4, 86) and x[1, 0] == 87 and loss.tolist() == [1, 2, 3, 4] and st["frame_iterations"] == 7
    for b, fr in enumerate(seq["frames"]):
        got = pickle.load(open(tmp_path / "res" / "0007" / fr / "000.pkl", "rb"))
        assert got["loss"] == b + 1 and got["transl"][0, 0] == b * 86 + 82 + 1
        assert os.path.exists(tmp_path / "mesh" / "0007" / fr / "000.obj")


@pytest.mark.needs_reference
def test_readers_equal_the_reference_on_its_demo_data():
    """the reference's own load_camera_para / read_keypoints on the data it ships (authoring container only)"""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_harness as H
    H.import_reference()
    with H.in_reference_dir():
        from utils.utils import load_camera_para
        from utils.data_parser import read_keypoints
    e_ref, i_ref = load_camera_para(os.path.join(REF, "data", "3DOH50K_Parameters.txt"))
    e, i = seqio.load_camera_para(os.path.join(REF, "data", "3DOH50K_Parameters.txt"))
    assert e.shape == (6, 4, 4) and np.array_equal(e, e_ref) and np.array_equal(i, i_ref)
    seq = seqio.load_sequence(os.path.join(REF, "data", "keypoints"), "0000")
    assert len(seq["cameras"]) == 6 and seq["frames"] == ["00001"] and seq["present"].all()
    for v, cam in enumerate(seq["cameras"]):
        ref = read_keypoints(os.path.join(REF, "data", "keypoints", "0000", cam, "00001_keypoints.json"),
                             use_hands=False, use_face=False).keypoints[0]
        assert np.array_equal(seq["gt_uv"][v, 0], ref[:, :2]) and np.array_equal(seq["conf"][v, 0], ref[:, 2])
