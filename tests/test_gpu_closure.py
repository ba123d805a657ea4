"""GPU parity of the CUDA closure (through the C ABI) with the reference-run golden fixtures
and with the oracle on fresh seeded inputs.  Tolerance: 1e-4 relative (max-norm per tensor),
the north-star fp32 bar."""
import numpy as np
import pytest
import torch

from oracle.lbfgs_oracle import PARAM_SEGMENTS
from tests import golden_util as G

pytestmark = pytest.mark.gpu
TOL = 1e-4


def make_ctx(model, cams, B, gmm=None, model_type="smpllsp"):
    from mvsmplfitting_b200.context import FittingContext
    ctx = FittingContext(0)
    ctx.set_model(model, model_type=model_type)
    if gmm is not None:
        ctx.set_gmm_from_dict(gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"])
    ctx.set_batch(B)
    return ctx


def check_against(c, tag, out, dense):
    assert G.relmax(out["loss"].cpu().numpy(), c["loss_" + tag]) < TOL
    assert G.relmax(out["joints"].cpu().numpy(), c["joints_" + tag]) < TOL
    assert G.relmax(out["proj"].cpu().numpy(), c["proj_" + tag]) < TOL
    g = out["grad"].cpu().numpy().astype(np.float64)
    g_ref = c["grad_" + tag]
    for a, e in PARAM_SEGMENTS:
        if np.abs(g_ref[:, a:e]).max() > 0:
            assert G.relmax(g[:, a:e], g_ref[:, a:e]) < TOL, (a, e)
    if dense:
        v = out["verts"].cpu().numpy().astype(np.float64)
        assert G.relmax(v[:, :64], c["verts_head_" + tag]) < TOL
        assert G.relmax(v.sum(1), c["verts_sum_" + tag]) < TOL


@pytest.mark.parametrize("name", G.closure_cases())
@pytest.mark.parametrize("dense", [False, True])
def test_closure_matches_reference_fixtures(name, dense, syn_model, syn_gmm):
    c = G.load_case(name)
    B = c["X"].shape[0]
    ctx = make_ctx(syn_model, c["cams"], B, syn_gmm, c["meta"]["model_type"])
    ctx.set_keypoints(c["gt_uv"], c["conf"], c["joint_weights"])
    ctx.set_loss(body_prior=c["meta"]["body_prior"], use_joints_conf=c["meta"]["use_joints_conf"],
                 fix_shape=c["meta"]["fix_shape"], frozen=("betas",) if c["meta"]["fix_shape"] else (), **c["w"])
    x = torch.tensor(c["X"], device="cuda")
    out = ctx.closure(x, want_grad=True, want_joints=True, want_proj=True, want_verts=dense)
    torch.cuda.synchronize()
    # fp32 reference output and (tighter arbiter) its fp64 run
    check_against(c, "f32", out, dense)
    check_against(c, "f64", out, dense)
    # forward-only call gives the same loss
    out2 = ctx.closure(x, want_grad=False)
    assert G.relmax(out2["loss"].cpu().numpy(), out["loss"].cpu().numpy().astype(np.float64)) < 1e-5   # sparse fp32 path vs dense TF32 path
    ctx.close()


@pytest.mark.parametrize("B,V", [(1, 8), (5, 4), (70, 8), (130, 3)])
def test_closure_matches_oracle_on_seeded_batches(B, V, syn_model, syn_gmm):
    """ragged batch sizes (not multiples of the 64-frame tile) against the oracle"""
    from mvsmplfitting_b200 import synthetic as S
    from oracle import closure_oracle as O
    cams = S.make_cameras(V)
    fr = S.make_frames(syn_model, cams, B, seed=100 + B)
    w = dict(data_weight=500.0 / 1536, body_pose_weight=4.78, shape_weight=5.0, bending_prior_weight=3.17 * 4.78)
    ctx = make_ctx(syn_model, cams, B, syn_gmm)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_loss(body_prior="gmm", **w)
    X = S.pack_params(fr["init"])
    out = ctx.closure(torch.tensor(X, device="cuda"), want_joints=True, want_proj=True)
    torch.cuda.synchronize()
    om = O.OracleModel.from_numpy(syn_model, dtype=torch.float64)
    pri = O.OraclePriors.gmm_from_dict(syn_gmm, torch.float64)
    sel = sorted(set([0, B - 1, B // 2] + list(range(0, B, max(1, B // 6)))))
    ref = O.closure_eval_batch(om, O.LossConfig(**w), pri, O.cams_to_torch(cams, torch.float64), X[sel],
                               fr["gt_uv"][:, sel], fr["conf"][:, sel], fr["joint_weights"])
    assert G.relmax(out["loss"].cpu().numpy()[sel], ref["loss"]) < TOL
    assert G.relmax(out["joints"].cpu().numpy()[sel], ref["joints"]) < TOL
    assert G.relmax(out["proj"].cpu().numpy()[:, sel], ref["proj"]) < TOL
    g = out["grad"].cpu().numpy()[sel]
    for a, e in PARAM_SEGMENTS:
        assert G.relmax(g[:, a:e], ref["grad"][:, a:e]) < TOL
    # frames are independent: evaluating a frame alone gives bit-identical results
    ctx1 = make_ctx(syn_model, cams, 1, syn_gmm)
    ctx1.set_keypoints(fr["gt_uv"][:, B - 1:B], fr["conf"][:, B - 1:B], fr["joint_weights"])
    ctx1.set_loss(body_prior="gmm", **w)
    o1 = ctx1.closure(torch.tensor(X[B - 1:B], device="cuda"))
    assert torch.equal(o1["loss"][0], out["loss"][B - 1])
    assert torch.equal(o1["grad"][0], out["grad"][B - 1])


def test_closure_is_deterministic(syn_model, syn_gmm):
    from mvsmplfitting_b200 import synthetic as S
    cams = S.make_cameras(8)
    fr = S.make_frames(syn_model, cams, 64, seed=9)
    ctx = make_ctx(syn_model, cams, 64, syn_gmm)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_loss(body_prior="gmm", data_weight=0.3, body_pose_weight=4.78, shape_weight=5.0, bending_prior_weight=15.0)
    x = torch.tensor(S.pack_params(fr["init"]), device="cuda")
    a = ctx.closure(x, want_verts=True)
    b = ctx.closure(x, want_verts=True)
    for k in a:
        assert torch.equal(a[k], b[k])
