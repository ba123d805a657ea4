"""CPU: the device routine behind use_vposer = 2 (6D -> rotation -> quaternion -> axis-angle and its hand-written
adjoint, mvs_math.cuh cont6d_to_aa_fwd/_bwd), compiled for the host, against the pinned restatement of
VPoser.py:29-174 (oracle/vposer_oracle.py, checked against the reference class in test_vposer_golden.py) and
its autograd gradient.  All four quaternion branches are exercised."""
import numpy as np
import pytest
import torch

from oracle import vposer_oracle as VO
from tests import golden_util as G
from tests.hostsim import cont6d_to_aa


def six_vectors(n, seed):
    rng = np.random.default_rng(seed)
    o = rng.normal(size=(n, 6))
    o[: n // 4] *= 0.05                      # short, nearly parallel columns: stresses both normalisations
    o[n // 4: n // 2, 1::2] += 3.0 * o[n // 4: n // 2, 0::2]
    return o


@pytest.mark.parametrize("use_double,tol", [(True, 1e-10), (False, 2e-4)])
def test_cont6d_matches_oracle_and_autograd(use_double, tol):
    o = six_vectors(4000, 3)
    rng = np.random.default_rng(4)
    daa = rng.normal(size=(o.shape[0], 3))
    aa, d_o, br = cont6d_to_aa(o, daa, use_double)
    assert set(np.unique(br)) == {0, 1, 2, 3}
    x = torch.tensor(o, dtype=torch.float64, requires_grad=True)
    ref = VO.quaternion_to_angle_axis(VO.matrot_to_quaternion(VO.cont6d_to_matrot(x)))
    (ref * torch.tensor(daa)).sum().backward()
    ref, gref = ref.detach().numpy(), x.grad.numpy()
    # away from the branch surfaces and from theta = pi the map is smooth; fp32 rows that sit on a branch
    # boundary may legitimately pick the neighbouring formula (same rotation, same aa up to rounding)
    ang = np.linalg.norm(ref, axis=1)
    ok = ang < np.pi - 1e-2
    assert ok.mean() > 0.95
    assert G.relmax(aa[ok], ref[ok]) < tol
    scale = np.abs(gref[ok]).max(axis=1, keepdims=True) + 1e-12
    err = np.abs(d_o[ok] - gref[ok]) / scale
    assert np.quantile(err.max(axis=1), 0.999) < (1e-8 if use_double else 2e-3)


def test_cont6d_adjoint_against_finite_differences():
    o = six_vectors(200, 9)
    rng = np.random.default_rng(10)
    daa = rng.normal(size=(o.shape[0], 3))
    aa, d_o, br = cont6d_to_aa(o, daa, True)
    h = 1e-6
    for k in range(6):
        e = np.zeros(6); e[k] = h
        ap, _, bp = cont6d_to_aa(o + e, daa, True)
        am, _, bm = cont6d_to_aa(o - e, daa, True)
        same = (bp == br) & (bm == br) & (np.linalg.norm(aa, axis=1) < np.pi - 1e-2)
        fd = ((ap - am) * daa).sum(1) / (2 * h)
        assert same.sum() > 150
        assert np.abs(fd[same] - d_o[same, k]).max() / (np.abs(d_o[same]).max() + 1e-12) < 1e-5
