"""CPU: oracle/init_oracle.py (restatement of recompute3D.py, umeyama.py and cv2.Rodrigues) against outputs of the
reference's own functions (tests/golden/init_s21.npz, written by oracle/make_golden_init.py)."""
import os

import numpy as np
import pytest

from oracle import init_oracle as IO
from tests import golden_util as G

GOLD = os.path.join(os.path.dirname(__file__), "golden", "init_s21.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


@pytest.mark.parametrize("V", [2, 4, 8, 16])
def test_triangulation_is_the_reference_bit_for_bit(V, gold):
    for b in range(gold["tri%d_j3" % V].shape[0]):
        kps = [np.concatenate([gold["tri%d_uv" % V][v, b], gold["tri%d_conf" % V][v, b][:, None]], axis=1) for v in range(V)]
        j3 = IO.triangulate(gold["tri%d_ext" % V], gold["tri%d_int" % V], kps)
        assert np.array_equal(j3, gold["tri%d_j3" % V][b])


def test_umeyama_as_written_is_the_reference(gold):
    for k in range(len(gold["um_n"])):
        n = int(gold["um_n"][k])
        rot, tr, sc = IO.umeyama_as_written(gold["um_src"][k, :n], gold["um_dst"][k, :n], bool(gold["um_est"][k]))
        assert np.array_equal(rot, gold["um_rot"][k]) and np.array_equal(tr, gold["um_trans"][k])
        assert sc == gold["um_scale"][k]


def test_published_umeyama_recovers_the_transform_the_reference_misses(gold):
    """the data are noisy similarity transforms: the published algorithm's residual is at the noise level and never
    above the as-written variant's; the scale (which does not depend on the transposed factor) is identical"""
    worse = 0
    for k in range(len(gold["um_n"])):
        n = int(gold["um_n"][k])
        src, dst, est = gold["um_src"][k, :n], gold["um_dst"][k, :n], bool(gold["um_est"][k])
        R, t, s = IO.umeyama(src, dst, est)
        assert abs(np.linalg.det(R) - 1) < 1e-9 and np.abs(R @ R.T - np.eye(3)).max() < 1e-9
        assert s == pytest.approx(float(gold["um_scale"][k]), rel=1e-12)
        res = np.linalg.norm(s * src @ R.T + t - dst)
        res_ref = np.linalg.norm(gold["um_scale"][k] * src @ gold["um_rot"][k].T + gold["um_trans"][k] - dst)
        assert res <= res_ref * (1 + 1e-9)
        if est:
            assert res < 0.01 * np.sqrt(3 * n) * 2.5
        worse += res_ref > 2 * res
    assert worse >= len(gold["um_n"]) // 2       # the LAPACK-sign-dependent rotation is off for most inputs


def test_rotmat_to_aa_is_cv2_rodrigues(gold):
    for R, r in zip(gold["rod_R"], gold["rod_r"]):
        mine = IO.rotmat_to_aa(R)
        ang = np.linalg.norm(r)
        if ang > np.pi - 1e-5:                   # at pi the axis sign is a convention of the branch; compare rotations
            assert abs(np.linalg.norm(mine) - ang) < 1e-6
            assert min(np.abs(mine - r).max(), np.abs(mine + r).max()) < 1e-5
        else:
            assert np.abs(mine - r).max() < 1e-9 + 1e-7 * ang


REF = os.path.join(os.path.dirname(__file__), "golden", "init_guess_ref.npz")


@pytest.mark.parametrize("tag,est", [("mv4", True), ("sv", False), ("sv_est", True)])
def test_whole_init_guess_as_written_is_the_reference_run(tag, est, syn_model):
    """init_guess + fix_params of the reference, run unmodified on the synthetic model (oracle/make_golden_init.py), against
    the restatement with as_written=True and numpy's own SVD (same LAPACK as the authoring run): multi-view triangulation and
    the single-view depth guess (init_guess.py:54-78)."""
    from mvsmplfitting_b200 import synthetic as S
    g = np.load(REF)
    z = lambda n: np.zeros((1, n))
    rest = S.model_keypoints_np(syn_model, z(10), z(3), z(69), z(3), np.ones((1, 1)), "smpllsp")[0]
    V = g[tag + "_ext"].shape[0]
    for b in range(g[tag + "_params"].shape[0]):
        kps = [np.concatenate([g[tag + "_uv"][v, b], g[tag + "_conf"][v, b][:, None]], axis=1) for v in range(V)]
        o = IO.init_guess(g[tag + "_ext"], g[tag + "_int"], kps, rest, est, 1.0, True, as_written=True)
        x = g[tag + "_params"][b]
        assert np.abs(o["global_orient"] - x[10:13]).max() < 2e-4, (tag, b)       # the reference's rest joints are float32
        assert np.abs(o["transl"] - x[82:85]).max() < 2e-4 * max(1.0, np.abs(x[82:85]).max())
        assert abs(o["scale"] - x[85]) < 1e-4 * x[85]
        assert (x[:10] == 0).all() and (x[13:19] == 1).all() and (x[19:82] == 0).all()      # fix_params
