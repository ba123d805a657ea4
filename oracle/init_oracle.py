"""TEST INFRASTRUCTURE (never imported by the product) -- CPU restatement of the reference's per-frame initial guess
(SURVEY §8f row N2): multi-view triangulation of the 17 joints, similarity alignment of the model's rest joints to
them, rotation matrix -> axis-angle.

Pinned by tests/golden/init_s21.npz (written by oracle/make_golden_init.py from the reference's own
`recompute3D`, `umeyama` and from cv2.Rodrigues, the third-party routine `init_guess.py:86` calls;
opencv-python 4.x as installed in the authoring container):

* `triangulate`           follows code/utils/recompute3D.py:24-61 line by line (parity target: as written).
* `umeyama_as_written`    follows code/utils/umeyama.py:18-109 line by line.  Its full-rank branch multiplies by
                          `V.T` where numpy's `svd` already returned V^H (umeyama.py:67), so its rotation depends on
                          the SIGN CONVENTION of the LAPACK build behind numpy (flipping the sign of one singular
                          pair changes U diag(d) Vh^T); the authors' two-candidate patch (:77-107) does not repair
                          that.  It is restated to PIN this file against the reference on this container's LAPACK,
                          and is not a portable parity target.
* `umeyama`               the published algorithm (Umeyama, PAMI 1991, eq. 38-43; scikit-image `_umeyama`):
                          R = U diag(d) Vh.  This is what the device routine implements ("intended semantics").
* `rotmat_to_aa`          cv2.Rodrigues (matrix -> vector) as published in OpenCV calib3d (cvRodrigues2).
"""
from __future__ import annotations

import numpy as np

TORSO = (5, 6, 11, 12)                     # init_guess.py:82-84 (use_torso=True is what main.py:77 passes)


def triangulate(extris, intris, keypoints):
    """recompute3D.py:24-61.  extris [V,4,4] world->camera, intris [V,3,3], keypoints [V,17,3] (u, v, conf)
    -> [17,3] world points.  Normal equations of the point closest to the V viewing rays, weighted by conf + 1e-6;
    AtA is rounded to float32 before the solve (recompute3D.py:52)."""
    kp = np.array(keypoints, dtype=np.float64).reshape(len(extris), -1, 3)
    conf = kp[:, :, 2].copy()
    kp[:, :, 2] = 1.0
    K = kp.shape[1]
    AtA = np.zeros((K, 3, 3))
    Atb = np.zeros((K, 3))
    for v in range(len(extris)):
        Kinv = np.linalg.inv(np.asarray(intris[v], dtype=np.float64))
        R = np.asarray(extris[v], dtype=np.float64)[:3, :3]
        t = np.asarray(extris[v], dtype=np.float64)[:3, 3]
        for i in range(K):
            n = Kinv @ kp[v, i]
            n = n / np.linalg.norm(n)
            P = R.T @ (np.eye(3) - np.outer(n, n))
            w = conf[v, i] + 1e-6
            AtA[i] += (P @ R) * w
            Atb[i] += (-P @ t) * w
    AtA = AtA.astype(np.float32)
    return np.stack([np.linalg.solve(AtA[i], Atb[i]) for i in range(K)])


def svd_sign_normalised(A):
    """numpy's SVD with the sign freedom of every singular pair (u_i, v_i) fixed: the largest-magnitude component of v_i is
    positive.  The convention of the device's as-written variant (mvs_init.cuh: umeyama_fit(as_written)); numpy's own
    signs are whatever the LAPACK build produces."""
    U, S, Vh = np.linalg.svd(A)
    U, Vh = U.copy(), Vh.copy()
    for i in range(Vh.shape[0]):
        if Vh[i, np.argmax(np.abs(Vh[i]))] < 0:
            Vh[i] *= -1
            U[:, i] *= -1
    return U, S, Vh


def umeyama_as_written(src, dst, estimate_scale, svd=np.linalg.svd):
    """umeyama.py:18-109, quirks kept: `V` is numpy's V^H and the full-rank branch transposes it (:67); `rot` is a
    VIEW of T, so negating its first two columns for the second candidate (:80-82) also changes the T that the
    returned translation is computed from (:100)."""
    src, dst = np.asarray(src, np.float64), np.asarray(dst, np.float64)
    num, dim = src.shape
    sm, dm = src.mean(0), dst.mean(0)
    sd, dd = src - sm, dst - dm
    A = dd.T @ sd / num
    d = np.ones(dim)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    T = np.eye(dim + 1)
    U, S, V = svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.nan * T
    if rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(V) > 0:
            T[:dim, :dim] = U @ V
        else:
            s = d[dim - 1]
            d[dim - 1] = -1
            T[:dim, :dim] = U @ np.diag(d) @ V
            d[dim - 1] = s
    else:
        T[:dim, :dim] = U @ np.diag(d) @ V.T
    scale = 1.0 / sd.var(axis=0).sum() * (S @ d) if estimate_scale else 1.0
    homo = np.insert(src, 3, 1, axis=1).T
    rot = T[:dim, :dim]
    rots, losses = [], []
    for i in range(2):
        if i == 1:
            rot[:, :2] *= -1
        M = np.eye(dim + 1)
        M[:dim, :dim] = rot * scale
        M[:dim, dim] = dm - scale * (T[:dim, :dim] @ sm)
        losses.append(np.linalg.norm((M @ homo).T[:, :3] - dst))
        rots.append(rot.copy())
    trans = dm - scale * (T[:dim, :dim] @ sm)
    return (rots[1] if losses[0] > losses[1] else rots[0]), trans, scale


def umeyama(src, dst, estimate_scale):
    """Umeyama 1991 eq. 38-43 (the algorithm umeyama.py:16 says it took from scikit-image): least-squares similarity
    dst ~ scale * R src + t.  Returns (R [3,3], t [3], scale)."""
    src, dst = np.asarray(src, np.float64), np.asarray(dst, np.float64)
    num, dim = src.shape
    sm, dm = src.mean(0), dst.mean(0)
    sd, dd = src - sm, dst - dm
    A = dd.T @ sd / num
    d = np.ones(dim)
    if np.linalg.det(A) < 0:
        d[dim - 1] = -1
    U, S, Vh = np.linalg.svd(A)
    rank = np.linalg.matrix_rank(A)
    if rank == 0:
        return np.full((3, 3), np.nan), np.full(3, np.nan), np.nan
    if rank == dim - 1:
        if np.linalg.det(U) * np.linalg.det(Vh) > 0:
            R = U @ Vh
        else:
            dd_ = d.copy()
            dd_[dim - 1] = -1
            R = U @ np.diag(dd_) @ Vh
    else:
        R = U @ np.diag(d) @ Vh
    scale = 1.0 / sd.var(axis=0).sum() * (S @ d) if estimate_scale else 1.0
    return R, dm - scale * (R @ sm), scale


def rotmat_to_aa(R):
    """cv2.Rodrigues(R)[0] for a 3x3 input (OpenCV calib3d cvRodrigues2, matrix branch), without its initial
    SVD re-orthonormalisation (the callers pass rotations)."""
    R = np.asarray(R, np.float64)
    r = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.sqrt((r @ r) * 0.25)
    c = min(1.0, max(-1.0, (R[0, 0] + R[1, 1] + R[2, 2] - 1.0) * 0.5))
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        rx = np.sqrt(max((R[0, 0] + 1) * 0.5, 0.0))
        ry = np.sqrt(max((R[1, 1] + 1) * 0.5, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        rz = np.sqrt(max((R[2, 2] + 1) * 0.5, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(rx) < abs(ry) and abs(rx) < abs(rz) and ((R[1, 2] > 0) != (ry * rz > 0)):
            rz = -rz
        v = np.array([rx, ry, rz])
        return v * (theta / np.linalg.norm(v))
    return r * (theta / (2.0 * s))


def single_view_joints(extri, intri, keypoints0, rest_joints):
    """init_guess.py:54-78 line by line: depth guess for single-view input.  keypoints0 [17,3] (u, v, conf) of the only view.
    Quirks kept: the 2-D torso height is the L-shoulder -- L-hip difference twice (:66), taken over (u, v, conf) rows."""
    joints = np.asarray(rest_joints, np.float64)
    extri, intri = np.asarray(extri, np.float64), np.asarray(intri, np.float64)
    torso3d = joints[[5, 6, 11, 12]]
    torso2d = np.asarray(keypoints0, np.float64)[[5, 6, 11, 12]]
    torso3d = np.insert(torso3d, 3, 1, axis=1).T
    torso3d = (np.dot(extri, torso3d).T)[:, :3]
    diff3d = np.array([torso3d[0] - torso3d[2], torso3d[1] - torso3d[3]])
    mean_height3d = np.mean(np.sqrt(np.sum(diff3d ** 2, axis=1)))
    diff2d = np.array([torso2d[0] - torso2d[2], torso2d[0] - torso2d[2]])
    mean_height2d = np.mean(np.sqrt(np.sum(diff2d ** 2, axis=1)))
    est_d = intri[0][0] * (mean_height3d / mean_height2d)
    cam_joints = np.dot(extri, np.insert(joints.copy(), 3, 1, axis=1).T)
    cam_joints[2, :] += est_d
    return (np.dot(np.linalg.inv(extri), cam_joints).T)[:, :3]


def init_guess(extris, intris, keypoints, rest_joints, estimate_scale, fixed_scale=1.0, use_torso=True,
               as_written=False, svd=np.linalg.svd):
    """init_guess.py:18-107: returns dict(joints3d [17,3], global_orient [3], transl [3], scale).
    rest_joints [17,3] = the model's joints at zero pose / zero shape / scale = fixed_scale (init_guess.py:36-52)."""
    if len(keypoints) == 1:
        j3 = single_view_joints(extris[0], intris[0], np.asarray(keypoints[0]).reshape(-1, 3), rest_joints)
    else:
        j3 = triangulate(extris, intris, keypoints)
    sel = list(TORSO) if use_torso else list(range(j3.shape[0]))
    fn = (lambda a, b, e: umeyama_as_written(a, b, e, svd=svd)) if as_written else umeyama
    R, t, s = fn(np.asarray(rest_joints, np.float64)[sel], j3[sel], estimate_scale)
    return dict(joints3d=j3, global_orient=rotmat_to_aa(R), transl=t, scale=(s if estimate_scale else fixed_scale), R=R)
