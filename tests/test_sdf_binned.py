"""CPU: the culling scheme of the accelerated all-faces SDF mode (oracle/sdf_binned.py) reproduces the
brute-force restatement of the reference kernel (oracle/sdf_ref.c, all 13 776 faces per voxel) BIT FOR BIT at the voxels
the fused SDF term samples, with more than two orders of magnitude fewer triangle tests."""
import numpy as np
import pytest

from mvsmplfitting_b200 import synthetic as S
from oracle import sdf_binned, sdf_oracle


def posed_vertices(model, seed):
    rng = np.random.RandomState(seed)
    from oracle import closure_oracle as O
    import torch
    om = O.OracleModel.from_numpy(model, dtype=torch.float64)
    x = np.zeros(86)
    x[:10] = rng.normal(0, 1, 10)
    x[10:13] = rng.normal(0, 0.3, 3)
    x[13:82] = rng.normal(0, 0.35, 69)          # strong pose: limbs come close to the torso
    x[85] = 1.0
    t = lambda a: torch.tensor(a[None])
    verts, _, _ = O.smpl_forward(om, t(x[:10]), t(x[10:13]), t(x[13:82]), t(x[82:85]), t(x[85:86]))
    return verts.numpy()


def sample_voxels(vn, G, step):
    """the <= 8 voxels around every `step`-th vertex, as grid_sample (align_corners=False) addresses them"""
    g = (vn[::step].astype(np.float64) + 1.0) * 0.5 * G - 0.5
    base = np.floor(g).astype(np.int64)
    ids = set()
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                ijk = base + [dx, dy, dz]
                ok = ((ijk >= 0) & (ijk < G)).all(1)
                for i, j, k in ijk[ok]:
                    ids.add(int((k * G + j) * G + i))
    return np.array(sorted(ids), dtype=np.int64)


@pytest.mark.parametrize("G,step,seed", [(32, 5, 1), (128, 16, 2)])
def test_binned_phi_is_the_brute_force_bit_for_bit(G, step, seed, syn_model):
    verts = posed_vertices(syn_model, seed)
    lo, hi = verts.min(0), verts.max(0)
    vn = ((verts - 0.5 * (lo + hi)) / (0.6 * (hi - lo).max())).astype(np.float32)       # fitting.py:356-364
    faces = syn_model["f"]
    ids = sample_voxels(vn, G, step)
    brute = sdf_oracle.sdf_voxels(faces, vn, G, ids, all_faces=True)
    bs = sdf_binned.BinnedSdf(faces, vn, G)
    phi, info = bs.phi(ids)
    assert np.array_equal(phi, brute)
    assert (brute > 0).sum() > 20                      # inside voxels exist, the parity rule is exercised
    F = faces.reshape(-1, 3).shape[0]
    tests = info["dist_candidates"].mean() + info["ray_candidates"].mean()
    assert tests < 2 * F / 100                         # > 100x fewer triangle tests than 2 F per voxel
    assert info["ring"].max() <= (3 if G == 128 else 12)      # at the product's G = 128 the first ring almost always decides
    if G == 128:
        assert info["ring"].mean() < 1.2
