"""CPU: the device routines of the initial guess (mvs_init.cuh: triangulate_point, umeyama_fit, rotmat_to_aa),
compiled for the host, against the reference-run golden vectors (tests/golden/init_s21.npz) and the pinned
oracle (oracle/init_oracle.py)."""
import os

import numpy as np
import pytest

from oracle import init_oracle as IO
from tests import hostsim as HS

GOLD = os.path.join(os.path.dirname(__file__), "golden", "init_s21.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def cams_of(gold, V):
    ext, intr = gold["tri%d_ext" % V], gold["tri%d_int" % V]
    return dict(R=ext[:, :3, :3], t=ext[:, :3, 3], f=np.stack([intr[:, 0, 0], intr[:, 1, 1]], 1),
                c=np.stack([intr[:, 0, 2], intr[:, 1, 2]], 1))


@pytest.mark.parametrize("V", [2, 4, 8, 16])
@pytest.mark.parametrize("use_double,tol", [(True, 1e-5), (False, 1e-4)])
def test_triangulation_matches_reference(V, use_double, tol, gold):
    """parity bar 1e-4 relative (the cameras travel as float32 like everywhere in the library; the reference rounds
    AtA to float32 itself, recompute3D.py:52)"""
    cams = cams_of(gold, V)
    if V == 2 and not use_double:
        pytest.skip("two opposite cameras: the rays are anti-parallel, float32 normal equations are too ill-conditioned "
                    "to compare (the product kernel accumulates in double)")
    for b in range(gold["tri%d_j3" % V].shape[0]):
        conf = gold["tri%d_conf" % V][:, b]
        X = HS.triangulate(cams, gold["tri%d_uv" % V][:, b], conf, use_double)
        ref = gold["tri%d_j3" % V][b]
        seen = (conf > 0).sum(0) >= (2 if use_double else 3)      # a joint seen once is a ray, not a point
        assert seen.sum() >= 8
        assert np.abs(X - ref)[seen].max() / np.abs(ref[seen]).max() < tol, (V, b)


@pytest.mark.parametrize("use_double,tol", [(True, 1e-9), (False, 2e-4)])
def test_umeyama_matches_published_algorithm(use_double, tol, gold):
    for k in range(len(gold["um_n"])):
        n = int(gold["um_n"][k])
        src, dst, est = gold["um_src"][k, :n], gold["um_dst"][k, :n], bool(gold["um_est"][k])
        R, t, s = IO.umeyama(src, dst, est)
        got = HS.umeyama(src, dst, est, use_double)
        assert got is not None
        assert np.abs(got[0] - R).max() < tol and np.abs(got[1] - t).max() < tol * 10 and abs(got[2] - s) < tol * 10
        assert got[2] == pytest.approx(float(gold["um_scale"][k]), rel=max(tol, 1e-9))      # reference scale
        aa = IO.rotmat_to_aa(R)
        assert np.abs(got[3] - aa).max() < tol * 10


def test_umeyama_recovers_exact_similarities_and_reflections_are_refused():
    rng = np.random.default_rng(5)
    for k in range(50):
        n = 4 if k % 2 else 17
        src = rng.normal(size=(n, 3)) * [0.3, 0.5, 0.02 if k % 3 == 0 else 0.2]      # every third set nearly planar
        aa = rng.normal(size=3)
        aa *= rng.uniform(0.1, 3.0) / np.linalg.norm(aa)      # angle below pi: the axis-angle is unique
        th = np.linalg.norm(aa)
        K = np.array([[0, -aa[2], aa[1]], [aa[2], 0, -aa[0]], [-aa[1], aa[0], 0]]) / th
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        s, t = rng.uniform(0.3, 3), rng.normal(size=3) * 2
        dst = s * src @ R.T + t
        got = HS.umeyama(src, dst, True, True)
        assert np.abs(got[0] - R).max() < 1e-8 and np.abs(got[1] - t).max() < 1e-8 and abs(got[2] - s) < 1e-8
        th2 = np.linalg.norm(got[3])
        assert abs(th2 - th) < 1e-7 and np.abs(got[3] - aa).max() < 1e-6
        mirrored = dst * [1, 1, -1]                       # no proper rotation maps src there: det(R) must stay +1
        Rm = HS.umeyama(src, mirrored, True, True)[0]
        assert abs(np.linalg.det(Rm) - 1) < 1e-9
        Ro = IO.umeyama(src, mirrored, True)[0]
        assert np.abs(Rm - Ro).max() < 1e-7
    line = np.outer(np.linspace(0, 1, 4), [1.0, 2.0, 3.0])                      # rank 1: refused
    assert HS.umeyama(line, line * 2, True, True) is None


def test_rotmat_to_aa_matches_cv2(gold):
    aa = HS.rotmat_to_aa(gold["rod_R"])
    for mine, r in zip(aa, gold["rod_r"]):
        ang = np.linalg.norm(r)
        if ang > np.pi - 1e-5:
            assert abs(np.linalg.norm(mine) - ang) < 1e-6 and min(np.abs(mine - r).max(), np.abs(mine + r).max()) < 1e-5
        else:
            assert np.abs(mine - r).max() < 1e-9 + 1e-7 * ang


@pytest.mark.parametrize("V,est,torso", [(4, True, True), (8, False, True), (16, True, False)])
def test_whole_initial_guess_matches_oracle(V, est, torso):
    """the data flow of init_guess_kernel on the host: triangulate every keypoint, align the (torso) rest joints"""
    from mvsmplfitting_b200 import synthetic as S
    model, cams = S.make_model(0), S.make_cameras(V)
    fr = S.make_frames(model, cams, 6, seed=900 + V)
    z = lambda n: np.zeros((1, n))
    rest = S.model_keypoints_np(model, z(10), z(3), z(69), z(3), np.ones((1, 1)), "smpllsp")[0].astype(np.float32)
    ext = np.tile(np.eye(4), (V, 1, 1))
    ext[:, :3, :3], ext[:, :3, 3] = cams["R"], cams["t"]
    intr = np.tile(np.eye(3), (V, 1, 1))
    intr[:, 0, 0], intr[:, 1, 1], intr[:, 0, 2], intr[:, 1, 2] = cams["f"][:, 0], cams["f"][:, 1], cams["c"][:, 0], cams["c"][:, 1]
    sel = list(IO.TORSO) if torso else list(range(17))
    for b in range(6):
        kps = [np.concatenate([fr["gt_uv"][v, b], fr["conf"][v, b][:, None]], axis=1) for v in range(V)]
        o = IO.init_guess(ext, intr, kps, rest, est, 1.0, torso)
        j3 = HS.triangulate(cams, fr["gt_uv"][:, b], fr["conf"][:, b], True)
        assert np.abs(j3 - o["joints3d"]).max() / np.abs(o["joints3d"]).max() < 1e-5
        R, t, s, aa = HS.umeyama(rest[sel], j3[sel], est, True)
        assert np.abs(aa - o["global_orient"]).max() < 1e-4 and np.abs(t - o["transl"]).max() < 1e-4
        assert abs(s - (o["scale"] if est else 1.0)) < 1e-4
        if torso:      # the point of the guess: the frame's true translation is recovered to a few cm
            assert np.abs(t - fr["gt"]["transl"][b]).max() < 0.5


def test_umeyama_as_written_matches_the_file_under_the_fixed_sign_convention(gold):
    """umeyama_fit(as_written): the expression of code/utils/umeyama.py (transposed V^H at :67, two-candidate patch, translation
    from the negated candidate) with the singular pairs' signs fixed; the oracle restatement of the file is pinned bit for bit
    to the reference run with numpy's own SVD (tests/test_init_golden.py) and takes the same convention through svd=."""
    differs = 0
    for k in range(len(gold["um_n"])):
        n = int(gold["um_n"][k])
        src, dst, est = gold["um_src"][k, :n], gold["um_dst"][k, :n], bool(gold["um_est"][k])
        R, t, s = IO.umeyama_as_written(src, dst, est, svd=IO.svd_sign_normalised)
        got = HS.umeyama(src, dst, est, True, as_written=True)
        assert np.abs(got[0] - R).max() < 1e-9 and np.abs(got[1] - t).max() < 1e-8 and abs(got[2] - s) < 1e-9
        assert abs(np.linalg.det(got[0]) - 1) < 1e-9
        gf = HS.umeyama(src, dst, est, False, as_written=True)
        assert np.abs(gf[0] - R).max() < 2e-3
        differs += np.abs(got[0] - IO.umeyama(src, dst, est)[0]).max() > 1e-3
    assert differs >= len(gold["um_n"]) // 2      # it really is a different rotation from the published algorithm's


def test_single_view_depth_guess_matches_the_restatement():
    from mvsmplfitting_b200 import synthetic as S
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "init_guess_ref.npz"))
    model = S.make_model(0)
    z = lambda n: np.zeros((1, n))
    rest = S.model_keypoints_np(model, z(10), z(3), z(69), z(3), np.ones((1, 1)), "smpllsp")[0].astype(np.float32)
    ext, intr = g["sv_ext"][0], g["sv_int"][0]
    cam = dict(R=ext[:3, :3], t=ext[:3, 3], f=np.array([intr[0, 0], intr[1, 1]]), c=np.array([intr[0, 2], intr[1, 2]]))
    for b in range(g["sv_uv"].shape[1]):
        kp = np.concatenate([g["sv_uv"][0, b], g["sv_conf"][0, b][:, None]], axis=1)
        ref = IO.single_view_joints(ext, intr, kp, rest)
        got = HS.single_view_joints(cam, rest, g["sv_uv"][0, b], g["sv_conf"][0, b])
        assert np.abs(got - ref).max() < 1e-5 * np.abs(ref).max()
        # whole single-view guess, as written, under the fixed sign convention
        o = IO.init_guess(ext[None], intr[None], [kp], rest, False, 1.0, True, as_written=True, svd=IO.svd_sign_normalised)
        R, t, s, aa = HS.umeyama(rest[list(IO.TORSO)], got[list(IO.TORSO)], False, True, as_written=True)
        assert np.abs(aa - o["global_orient"]).max() < 1e-4 and np.abs(t - o["transl"]).max() < 1e-3
        # published algorithm: the guess is a pure translation along the optical axis
        Rp = HS.umeyama(rest[list(IO.TORSO)], got[list(IO.TORSO)], False, True)[0]
        assert np.abs(Rp - np.eye(3)).max() < 1e-6
