"""Host-side data formats on either side of the fitting path: the reference's camera file,
per-view keypoint JSON files and per-frame result pickles / meshes, read and written for WHOLE SEQUENCES so that the
frames of a sequence go through `mvs_init_guess` + `mvs_fit` as one batch instead of the reference's per-frame loop
(code/main.py:32-89).

Reference interfaces mirrored (file:line of /root/reference/code):
  load_camera_para     utils/utils.py:352-394      camera file -> extrinsics [V,4,4], intrinsics [V,3,3]
  read_keypoints       utils/data_parser.py:42-90  OpenPose-style JSON -> [17,3] (u, v, confidence), body part only
                                                   (use_hands = use_face = False is the only configuration of the path)
  FittingData          utils/data_parser.py:257-433  folder layout <keypoints>/<serial>/<camera>/<frame>_keypoints.json,
                                                   joint weights (:340-358)
  save_results         utils/utils.py:729-761,856-890  result dict, pickle protocol 2 at <results>/<serial>/<frame>/000.pkl,
                                                   optional .obj mesh

Plain numpy / stdlib for the formats; the device work is behind FittingContext (torch only carries its buffers)."""
from __future__ import annotations

import json
import os
import pickle

import numpy as np
import torch

NUM_BODY_JOINTS = 17                     # data_parser.py:259


def load_camera_para(path: str):
    """utils.py:352-394: lines of 3 numbers build the intrinsics, lines of 4 numbers the [R | t] rows of the
    extrinsics (a [0, 0, 0, 1] row is appended), anything else (camera index lines, blanks) is skipped."""
    intr_rows, ext_rows = [], []
    with open(path) as f:
        for line in f:
            words = line.split()
            if len(words) == 3:
                intr_rows.append([float(w) for w in words])
            elif len(words) == 4:
                ext_rows.append([float(w) for w in words])
    intris = [intr_rows[i:i + 3] for i in range(0, len(intr_rows) - len(intr_rows) % 3, 3)]
    extris = [ext_rows[i:i + 3] + [[0.0, 0.0, 0.0, 1.0]] for i in range(0, len(ext_rows) - len(ext_rows) % 3, 3)]
    return np.array(extris), np.array(intris)


def camera_arrays(extris, intris, views=None) -> dict:
    """(extrinsics, intrinsics) -> the R [V,3,3], t [V,3], f [V,2], c [V,2] float32 arrays FittingContext.set_cameras
    takes (what create_camera receives per view at init.py:108-131)."""
    extris, intris = np.asarray(extris, np.float64), np.asarray(intris, np.float64)
    if views is not None:
        extris, intris = extris[list(views)], intris[list(views)]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(R=f32(extris[:, :3, :3]), t=f32(extris[:, :3, 3]), f=f32(np.stack([intris[:, 0, 0], intris[:, 1, 1]], 1)),
                c=f32(np.stack([intris[:, 0, 2], intris[:, 1, 2]], 1)))


def read_keypoints(path: str, person: int = 0):
    """data_parser.py:42-90 with use_hands = use_face = False: the first 17 (u, v, confidence) triplets of
    `pose_keypoints_2d` of one person as float32 [17,3]; None if the file lists fewer people."""
    with open(path) as f:
        data = json.load(f)
    people = data.get("people", [])
    if person >= len(people):
        return None
    kp = np.array(people[person]["pose_keypoints_2d"], dtype=np.float32).reshape(-1, 3)[:NUM_BODY_JOINTS]
    return kp


def joint_weights(pose_format: str = "coco17", use_hip: bool = True) -> np.ndarray:
    """FittingData.get_joint_weights, data_parser.py:340-358: hips (11, 12) are ignored unless lsp14 with use_hip"""
    w = np.ones(NUM_BODY_JOINTS, dtype=np.float32)
    if pose_format != "lsp14" or not use_hip:
        w[11] = 0.0
        w[12] = 0.0
    return w


def list_sequence(keyp_folder: str, serial: str):
    """(cameras, frames): sorted camera folders of one serial and the sorted union of frame names seen in any of them
    (FittingData.__init__ sorts the same way, data_parser.py:296-313, on the image folders)."""
    root = os.path.join(keyp_folder, serial)
    cameras = sorted(d for d in os.listdir(root) if os.path.isdir(os.path.join(root, d)))
    frames = set()
    for cam in cameras:
        for fn in os.listdir(os.path.join(root, cam)):
            if fn.endswith("_keypoints.json") and not fn.startswith("."):
                frames.add(fn[: -len("_keypoints.json")])
    return cameras, sorted(frames)


def load_sequence(keyp_folder: str, serial: str, cameras=None, frames=None, person: int = 0) -> dict:
    """All detections of a sequence as the arrays the device path takes: gt_uv [V,B,17,2], conf [V,B,17] float32.
    A (view, frame) without a keypoint file or without that person gets confidence 0 for every joint -- the reference
    drops such a view for that frame (main.py:45-56); with confidence 0 it weighs (w * conf)^2 = 0 in the data term
    and 1e-6 in the triangulation (recompute3D.py:49)."""
    all_cams, all_frames = list_sequence(keyp_folder, serial)
    cameras = list(cameras) if cameras is not None else all_cams
    frames = list(frames) if frames is not None else all_frames
    V, B = len(cameras), len(frames)
    gt_uv = np.zeros((V, B, NUM_BODY_JOINTS, 2), dtype=np.float32)
    conf = np.zeros((V, B, NUM_BODY_JOINTS), dtype=np.float32)
    present = np.zeros((V, B), dtype=bool)
    for v, cam in enumerate(cameras):
        for b, fr in enumerate(frames):
            fn = os.path.join(keyp_folder, serial, cam, fr + "_keypoints.json")
            if not os.path.exists(fn):
                continue
            kp = read_keypoints(fn, person)
            if kp is None or kp.shape[0] < NUM_BODY_JOINTS:
                continue
            gt_uv[v, b], conf[v, b], present[v, b] = kp[:, :2], kp[:, 2], True
    return dict(gt_uv=gt_uv, conf=conf, present=present, cameras=cameras, frames=frames, serial=serial)


PARAM_SLICES = (("betas", 0, 10), ("global_orient", 10, 13), ("body_pose", 13, 82), ("transl", 82, 85), ("scale", 85, 86))


def result_from_params(x, loss: float, pose_embedding=None, body_pose=None) -> dict:
    """One frame's 86-vector -> the dict non_linear_solver returns (non_linear_solver.py:283-287): the model's named
    parameters with their reference shapes ([1,10], [1,3], [1,69], [1,3], [1]), 'loss', 'pose_embedding'.  body_pose: the
    decoded pose when the pose slot of x holds a VPoser latent code."""
    x = np.asarray(x, dtype=np.float32).reshape(-1)
    out = {}
    for name, a, b in PARAM_SLICES:
        out[name] = x[a:b].copy() if name == "scale" else x[a:b].reshape(1, -1).copy()
    if body_pose is not None:
        out["body_pose"] = np.asarray(body_pose, np.float32).reshape(1, 69).copy()
    out["loss"] = float(loss)
    out["pose_embedding"] = None if pose_embedding is None else np.asarray(pose_embedding, np.float32).reshape(1, -1)
    return out


def finalize_result(result: dict) -> dict:
    """save_results' post-processing (utils.py:741-768): ankles (body_pose[18:24] = SMPL joints 7, 8), feet (27:33 = joints
    10, 11) and wrists + hands (57: = joints 20-23) are zeroed, 'pose' = [global_orient | body_pose].  With VPoser the
    'body_pose' entry must already hold the DECODED pose (result_from_params(..., body_pose=...)), as :741-743 decodes the
    latent code before zeroing; 'pose_embedding' keeps the code."""
    bp = result["body_pose"]
    bp[:, 18:24] = 0.0
    bp[:, 27:33] = 0.0
    bp[:, 57:] = 0.0
    result["pose"] = np.hstack((result["global_orient"], bp))
    return result


def write_obj(path: str, verts, faces) -> None:
    """Wavefront .obj as trimesh's exporter writes a bare mesh (utils.py:884-890): `v x y z` lines, 1-based `f a b c`"""
    verts, faces = np.asarray(verts, np.float64).reshape(-1, 3), np.asarray(faces, np.int64).reshape(-1, 3)
    with open(path, "w") as f:
        for v in verts:
            f.write("v %.8f %.8f %.8f\n" % (v[0], v[1], v[2]))
        for t in faces + 1:
            f.write("f %d %d %d\n" % (t[0], t[1], t[2]))


def save_results(result_folder: str, serial: str, fn: str, result: dict, person_id: int = 0, verts=None, faces=None,
                 mesh_folder: str | None = None) -> str:
    """utils.py:856-863 (+ :884-890 when a mesh is given): <result_folder>/<serial>/<fn>/000.pkl, pickle protocol 2"""
    d = os.path.join(result_folder, serial, fn)
    os.makedirs(d, exist_ok=True)
    out = os.path.join(d, "%03d.pkl" % person_id)
    with open(out, "wb") as f:
        pickle.dump(finalize_result(result), f, protocol=2)
    if verts is not None and faces is not None and mesh_folder is not None:
        md = os.path.join(mesh_folder, serial, fn)
        os.makedirs(md, exist_ok=True)
        write_obj(os.path.join(md, "%03d.obj" % person_id), verts, faces)
    return out


def _uses_latent_pose(stage_cfgs) -> bool:
    """True when the stages optimise VPoser's latent code on the device (use_vposer = 2); mixing encodings is refused"""
    kinds = {int(c.use_vposer) for c in stage_cfgs}
    if 1 in kinds:
        raise NotImplementedError("use_vposer = 1 (pose decoded by the caller) has no whole-sequence path: upload the decoder "
                                  "(FittingContext.set_vposer) and use use_vposer = 2")
    if kinds == {0, 2}:
        raise ValueError("stage configurations mix axis-angle and latent-code poses")
    return kinds == {2}


def _check_missing_views(seq: dict, stage_cfgs) -> None:
    """a (view, frame) without detections carries confidence 0; that only removes it from the data term when the loss uses the
    confidences (the reference drops such a view from the camera list instead, main.py:45-56)"""
    if not seq["present"].all() and not all(int(c.use_joints_conf) for c in stage_cfgs):
        raise ValueError("sequence has (view, frame) pairs without detections: every stage needs use_joints_conf = 1")


def _export(ctx, seq, params, loss, latent, result_folder, mesh_folder, faces, frames=None, rows=None):
    """per-frame result pickles (+ meshes) of `rows` of the batch (default: all), named by `frames`"""
    x = params.cpu().numpy()
    B = x.shape[0]
    rows = list(range(B)) if rows is None else list(rows)
    frames = seq["frames"] if frames is None else frames
    decoded = ctx.vposer_decode(params).cpu().numpy() if latent else None      # utils.py:741-743
    verts = None
    if mesh_folder is not None:          # the mesh is rebuilt from the SAVED pose (extremities zeroed), utils.py:865-871
        saved = params.clone()
        if latent:
            saved[:, 13:82] = torch.as_tensor(decoded, device=saved.device)
        for a, b in ((18, 24), (27, 33), (57, 69)):
            saved[:, 13 + a:13 + b] = 0.0
        verts = ctx.forward_only(saved, want_verts=True)["verts"].cpu().numpy()
    for b, fr in zip(rows, frames):
        res = result_from_params(x[b], loss[b], pose_embedding=x[b, 13:45] if latent else None,
                                 body_pose=decoded[b] if latent else None)
        save_results(result_folder, seq["serial"], fr, res, verts=None if verts is None else verts[b], faces=faces,
                     mesh_folder=mesh_folder)


def fit_sequence(ctx, seq: dict, stage_cfgs, opt_cfg=None, estimate_scale: bool = False, fixed_scale: float = 1.0,
                 use_hip: bool = True, pose_format: str = "coco17", result_folder: str | None = None,
                 mesh_folder: str | None = None, faces=None, umeyama_as_written: bool = False):
    """The frame loop of main.py:32-89 for one sequence as ONE batch (every frame from its own initial guess, is_seq = False):
    upload the detections, initial guess on the device (init_guess + fix_params), all stages (`mvs_fit`), results per
    frame.  `ctx` is a FittingContext whose model, priors (GMM / VPoser decoder), cameras and batch (= number of frames) are
    set.  With use_vposer = 2 stage configurations (the reference's default, cfg_files/fit_smpl.yaml) the latent code
    starts at 0 (init_guess.py:96-98), and results carry the decoded pose plus 'pose_embedding' (utils.py:741-759).
    Returns (params [B,86] numpy -- pose slot = latent code under VPoser --, final_loss [B], stats)."""
    B = len(seq["frames"])
    assert ctx.B == B, "context batch %d != frames %d" % (ctx.B, B)
    latent = _uses_latent_pose(stage_cfgs)
    _check_missing_views(seq, stage_cfgs)
    ctx.set_keypoints(seq["gt_uv"], seq["conf"], joint_weights(pose_format, use_hip))
    ctx.set_loss(config=stage_cfgs[0])           # mvs_init_guess zeroes the pose slot when it holds a latent code
    params, _ = ctx.init_guess(estimate_scale=estimate_scale, fixed_scale=fixed_scale, use_torso=True, hip_seed=1.0,
                               want_joints3d=False, umeyama_as_written=umeyama_as_written)
    final, stats = ctx.fit(params, stage_cfgs, opt_cfg)
    loss = final.cpu().numpy()
    if result_folder is not None:
        _export(ctx, seq, params, loss, latent, result_folder, mesh_folder, faces)
    return params.cpu().numpy(), loss, stats


def fit_sequences(ctx, seqs, stage_cfgs, opt_cfg=None, estimate_scale: bool = False, fixed_scale: float = 1.0,
                  use_hip: bool = True, pose_format: str = "coco17", result_folder: str | None = None,
                  mesh_folder: str | None = None, faces=None, umeyama_as_written: bool = False, reinit_loss: float = 5000.0):
    """main.py:32-89 with is_seq = True for S sequences in lock-step: the batch of step t is frame t of every sequence
    (context batch = S; same cameras).  A sequence's first frame gets the initial guess and every stage; a later frame
    starts from the previous frame's result as load_init + fix_params leave it (init_guess.py:137-166,190-212: betas,
    global_orient, transl, scale kept; body pose back to the hip seed, or the previous latent code under VPoser), skips
    stages 0 and 1 and damps stage 2's pose prior (non_linear_solver.py:157-162, `mvs_fit_seq`) -- unless the previous loss
    was above 5000, which re-runs the initial guess for that sequence (:141-144).  Sequences may differ in length.
    Returns a list (per sequence) of (params [T_s,86], loss [T_s]) and the summed stats."""
    S_ = len(seqs)
    assert ctx.B == S_, "context batch %d != sequences %d" % (ctx.B, S_)
    latent = _uses_latent_pose(stage_cfgs)
    for q in seqs:
        _check_missing_views(q, stage_cfgs)
    T = max(len(q["frames"]) for q in seqs)
    V = seqs[0]["gt_uv"].shape[0]
    jw = joint_weights(pose_format, use_hip)
    prev = None
    prev_loss = np.full(S_, np.inf, np.float32)
    out_x = [np.zeros((len(q["frames"]), 86), np.float32) for q in seqs]
    out_l = [np.zeros(len(q["frames"]), np.float32) for q in seqs]
    tot = dict(frame_iterations=0, frame_evals=0, rounds=0, frames_nan=0)
    for t in range(T):
        live = np.array([t < len(q["frames"]) for q in seqs])
        gt = np.zeros((V, S_, NUM_BODY_JOINTS, 2), np.float32)
        cf = np.zeros((V, S_, NUM_BODY_JOINTS), np.float32)
        for s_, q in enumerate(seqs):
            if live[s_]:
                gt[:, s_], cf[:, s_] = q["gt_uv"][:, t], q["conf"][:, t]
            else:                              # a finished sequence idles on its last frame (results are not kept)
                gt[:, s_], cf[:, s_] = q["gt_uv"][:, -1], q["conf"][:, -1]
        ctx.set_keypoints(gt, cf, jw)
        ctx.set_loss(config=stage_cfgs[0])
        cold, _ = ctx.init_guess(estimate_scale=estimate_scale, fixed_scale=fixed_scale, use_torso=True, hip_seed=1.0,
                                 want_joints3d=False, umeyama_as_written=umeyama_as_written)
        warm = np.zeros(S_, bool) if prev is None else (prev_loss <= reinit_loss)
        params = cold
        if warm.any():
            carried = prev.clone()
            if not latent:                     # fix_params: the body pose restarts from the hip seed
                carried[:, 13:82] = cold[:, 13:82]
            wm = torch.as_tensor(warm, device=params.device)
            params = torch.where(wm[:, None], carried, cold).contiguous()
        final, st = ctx.fit(params, stage_cfgs, opt_cfg, warm=warm if warm.any() else None)
        for k in tot:
            tot[k] += st[k]
        prev, prev_loss = params, final.cpu().numpy()
        x = params.cpu().numpy()
        for s_, q in enumerate(seqs):
            if live[s_]:
                out_x[s_][t], out_l[s_][t] = x[s_], prev_loss[s_]
        if result_folder is not None:
            for s_, q in enumerate(seqs):
                if live[s_]:
                    _export(ctx, q, params, prev_loss, latent, result_folder, mesh_folder, faces, frames=[q["frames"][t]], rows=[s_])
    return list(zip(out_x, out_l)), tot
