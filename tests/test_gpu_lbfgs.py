"""GPU: the device-resident batched L-BFGS (mvs_lbfgs_run) against
  (1) the reference's own optimiser trace (golden fixture written from lbfgs_ls.LBFGS + run_fitting),
  (2) the oracle L-BFGS driving the SAME CUDA closure (isolates the optimiser from the closure).
Line-search branches sit on fp32 scalars, so trajectories are compared with tolerance."""
import numpy as np
import pytest
import torch

from mvsmplfitting_b200 import synthetic as S
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def make_ctx(model, cams, B, gmm):
    from mvsmplfitting_b200.context import FittingContext
    ctx = FittingContext(0)
    ctx.set_model(model)
    ctx.set_gmm_from_dict(gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"])
    ctx.set_batch(B)
    return ctx


@pytest.mark.parametrize("name", ["lbfgs_traj_s3", "lbfgs_traj_s0"])
def test_device_lbfgs_vs_reference_trace(name, syn_model, syn_gmm):
    z = np.load(G.GOLD + "/%s.npz" % name)
    f = int(z["frame"])
    w = z["weights"]
    cams = S.make_cameras(8)
    ctx = make_ctx(syn_model, cams, 1, syn_gmm)
    ctx.set_keypoints(z["gt_uv"][:, f:f + 1], z["conf"][:, f:f + 1], np.ones(17, np.float32))
    ctx.set_loss(body_prior="gmm", data_weight=w[0], body_pose_weight=w[1], shape_weight=w[2], bending_prior_weight=w[3])
    x = torch.tensor(z["x0"][None], device="cuda")
    final, st = ctx.lbfgs_run(x)
    ref_evals, ref_iters = len(z["trace"]), int(z["n_iter"])
    assert abs(float(final[0]) - float(z["final"])) / float(z["final"]) < 2e-3
    assert abs(st["frame_evals"] - ref_evals) <= max(6, ref_evals // 5), (st, ref_evals)
    assert abs(st["frame_iterations"] - ref_iters) <= max(6, ref_iters // 5), (st, ref_iters)
    assert np.abs(x.cpu().numpy()[0] - z["x_final"]).max() < 5e-2


def test_device_lbfgs_vs_oracle_lbfgs_on_the_same_closure(syn_model, syn_gmm):
    from oracle import lbfgs_oracle as L
    cams = S.make_cameras(4)
    B = 6
    fr = S.make_frames(syn_model, cams, B, seed=21)
    w = dict(data_weight=500.0 / 1536, body_pose_weight=4.78, shape_weight=5.0, bending_prior_weight=3.17 * 4.78)
    X0 = S.pack_params(fr["init"])
    ctx = make_ctx(syn_model, cams, B, syn_gmm)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_loss(body_prior="gmm", **w)
    x = torch.tensor(X0, device="cuda")
    cfg = ctx.make_lbfgs_config(max_outer=4)
    final, st = ctx.lbfgs_run(x, cfg)
    xs = x.cpu().numpy()
    tot_evals = tot_iters = 0
    for b in range(B):
        c1 = make_ctx(syn_model, cams, 1, syn_gmm)
        c1.set_keypoints(fr["gt_uv"][:, b:b + 1], fr["conf"][:, b:b + 1], fr["joint_weights"])
        c1.set_loss(body_prior="gmm", **w)

        def fg(xx, c1=c1):
            o = c1.closure(xx.reshape(1, -1).cuda().contiguous())
            return float(o["loss"][0]), o["grad"][0].cpu()
        opt = L.LBFGSOracle(torch.tensor(X0[b]), fg, max_iter=30)
        fin, _ = L.run_fitting(opt, 4, 1e-9, 1e-9)
        tot_evals += opt.evals
        tot_iters += opt.iters
        assert abs(float(final[b]) - fin) / abs(fin) < 1e-3, (b, float(final[b]), fin)
        assert np.abs(xs[b] - opt.x.numpy()).max() < 2e-2
        c1.close()
    assert abs(st["frame_evals"] - tot_evals) <= max(6, tot_evals // 10), (st, tot_evals)
    assert abs(st["frame_iterations"] - tot_iters) <= max(6, tot_iters // 10), (st, tot_iters)


def test_device_lbfgs_batch_equals_single_frame_runs(syn_model, syn_gmm):
    """frames are independent problems: a frame optimised inside a batch of 70 follows exactly the
    trajectory it follows alone (bit-identical parameters), and frozen parameters stay untouched"""
    cams = S.make_cameras(8)
    B = 70
    fr = S.make_frames(syn_model, cams, B, seed=33)
    w = dict(data_weight=500.0 / 1536, body_pose_weight=57.4, shape_weight=10.0, bending_prior_weight=3.17 * 57.4)
    X0 = S.pack_params(fr["init"])
    X0[:, 85] = 1.3
    ctx = make_ctx(syn_model, cams, B, syn_gmm)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_loss(body_prior="gmm", frozen=("scale",), **w)
    x = torch.tensor(X0, device="cuda")
    final, st = ctx.lbfgs_run(x, ctx.make_lbfgs_config(max_outer=3))
    assert st["frames_nan"] == 0
    assert torch.all(x[:, 85] == 1.3)
    l0 = ctx.closure(torch.tensor(X0, device="cuda"), want_grad=False)["loss"]
    assert bool((final < l0).all())
    for b in (0, 37, 69):
        c1 = make_ctx(syn_model, cams, 1, syn_gmm)
        c1.set_keypoints(fr["gt_uv"][:, b:b + 1], fr["conf"][:, b:b + 1], fr["joint_weights"])
        c1.set_loss(body_prior="gmm", frozen=("scale",), **w)
        x1 = torch.tensor(X0[b:b + 1], device="cuda")
        f1, _ = c1.lbfgs_run(x1, c1.make_lbfgs_config(max_outer=3))
        assert torch.equal(x1[0], x[b])
        assert torch.equal(f1[0], final[b])
        c1.close()


def test_fit_host_runs_the_four_stage_schedule(syn_model, syn_gmm):
    cams = S.make_cameras(4)
    B = 9
    fr = S.make_frames(syn_model, cams, B, seed=5)
    ctx = make_ctx(syn_model, cams, B, syn_gmm)
    sw = S.STAGE_WEIGHTS
    stages = [ctx.make_loss_config(data_weight=500.0 / 1536, body_pose_weight=sw["body_pose_prior_weights"][i],
                                   shape_weight=sw["shape_weights"][i],
                                   bending_prior_weight=3.17 * sw["body_pose_prior_weights"][i], body_prior="gmm")
              for i in range(4)]
    X = S.pack_params(fr["init"]).copy()
    final, st = ctx.fit_host(X, fr["gt_uv"], fr["conf"], fr["joint_weights"], stages)
    assert st["frame_evals"] > 4 * B and st["frame_iterations"] > 0
    assert np.isfinite(final).all() and np.isfinite(X).all()
    # every frame ends below the loss of the initial guess under the last stage's weights
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_loss(config=stages[3])
    l0 = ctx.closure(torch.tensor(S.pack_params(fr["init"]), device="cuda"), want_grad=False)["loss"].cpu().numpy()
    l1 = ctx.closure(torch.tensor(X, device="cuda"), want_grad=False)["loss"].cpu().numpy()
    assert (l1 < l0).all()
    # run_fitting returns the loss at the ENTRY of its last step() (fitting.py:100,140), never below the final point's
    assert (l1 <= final * (1 + 1e-3)).all()
