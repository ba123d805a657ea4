"""CPU: the oracle (oracle/closure_oracle.py, oracle/lbfgs_oracle.py) against the
reference-run golden fixtures in tests/golden/ (written by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from mvsmplfitting_b200 import synthetic as S
from oracle import closure_oracle as O
from oracle import lbfgs_oracle as L
from tests import golden_util as G

TOL = 1e-4      # north-star tolerance (relative, fp32); the oracle is far inside it


def test_synthetic_model_is_the_one_the_fixtures_used(syn_model):
    assert G.model_checksum(syn_model) == open(G.GOLD + "/MODEL_CHECKSUM.txt").read().strip()


@pytest.mark.parametrize("name", G.closure_cases())
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_oracle_closure_matches_reference(name, dtype, syn_model, syn_gmm):
    c = G.load_case(name)
    assert str(c["model_checksum"]) == G.model_checksum(syn_model)
    tag = "f32" if dtype == torch.float32 else "f64"
    om = O.OracleModel.from_numpy(syn_model, dtype=dtype, model_type=c["meta"]["model_type"])
    pri = O.OraclePriors.gmm_from_dict(syn_gmm, dtype) if c["meta"]["body_prior"] == "gmm" else O.OraclePriors("l2")
    cfg = O.LossConfig(use_joints_conf=c["meta"]["use_joints_conf"], fix_shape=c["meta"]["fix_shape"], **c["w"])
    r = O.closure_eval_batch(om, cfg, pri, O.cams_to_torch(c["cams"], dtype), c["X"], c["gt_uv"], c["conf"],
                             c["joint_weights"], want_verts=True)
    tol = 2e-6 if dtype == torch.float32 else 1e-12
    assert G.relmax(r["loss"], c["loss_" + tag]) < tol
    g_ref = c["grad_" + tag].copy()
    if c["meta"]["fix_shape"]:
        r["grad"][:, :10] = 0.0        # frozen betas carry no grad in the reference
    for a, b in L.PARAM_SEGMENTS:
        if np.abs(g_ref[:, a:b]).max() > 0:
            assert G.relmax(r["grad"][:, a:b], g_ref[:, a:b]) < 50 * tol
    assert G.relmax(r["joints"], c["joints_" + tag]) < tol
    assert G.relmax(r["proj"], c["proj_" + tag]) < tol
    assert G.relmax(r["verts"][:, :64], c["verts_head_" + tag]) < tol
    assert G.relmax(r["verts"].sum(1), c["verts_sum_" + tag]) < 20 * tol


@pytest.mark.parametrize("name", ["lbfgs_traj_s3", "lbfgs_traj_s0"])
def test_oracle_lbfgs_follows_reference_trajectory(name, syn_model, syn_gmm):
    z = np.load(G.GOLD + "/%s.npz" % name)
    w = z["weights"]
    cfg = O.LossConfig(data_weight=float(w[0]), body_pose_weight=float(w[1]), shape_weight=float(w[2]),
                       bending_prior_weight=float(w[3]))
    om = O.OracleModel.from_numpy(syn_model)
    pri = O.OraclePriors.gmm_from_dict(syn_gmm)
    ct = O.cams_to_torch(S.make_cameras(8), torch.float32)
    f = int(z["frame"])
    losses = []

    def fg(x):
        r = O.closure_eval(om, cfg, pri, ct, x.numpy(), z["gt_uv"][:, f], z["conf"][:, f], np.ones(17, np.float32))
        losses.append(r["loss"])
        return r["loss"], torch.tensor(r["grad"])
    opt = L.LBFGSOracle(torch.tensor(z["x0"]), fg, max_iter=30)
    final, _ = L.run_fitting(opt, 30, 1e-9, 1e-9)
    ref = z["trace"]
    # closures differ in the last ulps, so trajectories are compared with tolerance
    n = min(len(ref), len(losses), 20)
    assert G.relmax(losses[:n], ref[:n]) < 1e-3
    assert abs(len(losses) - len(ref)) <= max(4, len(ref) // 10)
    assert abs(final - float(z["final"])) / float(z["final"]) < 1e-3
