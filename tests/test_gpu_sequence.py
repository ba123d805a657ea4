"""GPU (single rank): the per-frame quadratic anchor of the jointly regularised sequence mode and the
SequenceFitter sweeps.  The term is not in the reference (parity unpinned); its value / gradient are checked
against the closed form on top of the oracle's closure."""
import numpy as np
import pytest
import torch

from mvsmplfitting_b200 import sequence as Q
from mvsmplfitting_b200 import synthetic as S
from oracle import closure_oracle as O
from tests import golden_util as G

pytestmark = pytest.mark.gpu
W = dict(data_weight=500.0 / 1536, body_pose_weight=4.78, shape_weight=5.0, bending_prior_weight=3.17 * 4.78)


@pytest.mark.parametrize("mode", [0, 1])
def test_anchor_term_value_and_gradient(mode, syn_model, syn_gmm):
    from mvsmplfitting_b200.context import FittingContext
    cams = S.make_cameras(4)
    B = 5
    fr = S.make_frames(syn_model, cams, B, seed=12)
    X = S.pack_params(fr["init"])
    rng = np.random.RandomState(0)
    anchor = (X + rng.normal(0, 0.05, X.shape)).astype(np.float32)
    weight = (rng.uniform(0, 3, X.shape) * Q.smooth_mask().numpy()[None]).astype(np.float32)
    ctx = FittingContext(0)
    ctx.set_model(syn_model); ctx.set_gmm_from_dict(syn_gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_exec_mode(mode)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_anchor(torch.tensor(anchor), torch.tensor(weight))
    ctx.set_loss(body_prior="gmm", **W)
    out = ctx.closure(torch.tensor(X, device="cuda"))
    om = O.OracleModel.from_numpy(syn_model, dtype=torch.float64)
    ref = O.closure_eval_batch(om, O.LossConfig(**W), O.OraclePriors.gmm_from_dict(syn_gmm, torch.float64),
                               O.cams_to_torch(cams, torch.float64), X, fr["gt_uv"], fr["conf"], fr["joint_weights"])
    d = X.astype(np.float64) - anchor
    loss_ref = ref["loss"] + (weight * d * d).sum(1)
    grad_ref = ref["grad"] + 2 * weight * d
    assert G.relmax(out["loss"].cpu().numpy(), loss_ref) < 1e-4
    assert G.relmax(out["grad"].cpu().numpy(), grad_ref) < 1e-4
    ctx.set_anchor(None)
    ctx.set_loss(body_prior="gmm", **W)
    out0 = ctx.closure(torch.tensor(X, device="cuda"))
    assert G.relmax(out0["loss"].cpu().numpy(), ref["loss"]) < 1e-4
    ctx.close()


def test_sequence_fitter_smooths_the_track(syn_model, syn_gmm):
    cams = S.make_cameras(4)
    T = 24
    fr = S.make_frames(syn_model, cams, T, seed=3, smooth_walk=True)
    from mvsmplfitting_b200.context import FittingContext
    stages = [FittingContext.make_loss_config(body_prior="gmm", **W)]
    mask = Q.smooth_mask()
    energies = {}
    for lam in (0.0, 200.0):
        fit = Q.SequenceFitter(syn_model, cams, T, gmm=syn_gmm, device=0)
        assert (fit.start, fit.stop) == (0, T)
        x = torch.tensor(S.pack_params(fr["init"]), device="cuda")
        stats = fit.fit(x, fr["gt_uv"], fr["conf"], fr["joint_weights"], stages, smooth_weight=lam, sweeps=3, mask=mask)
        assert all(s["frames_nan"] == 0 for s in stats)
        energies[lam] = float(Q.smoothness_energy(x.cpu(), None, 1.0, mask))
        if lam > 0:
            es = [s["smooth_energy"] for s in stats]
            assert es[-1] <= es[0] * 1.05 and np.isfinite(es).all()
        fit.ctx.close()
    assert energies[200.0] < 0.8 * energies[0.0]


def test_fit_sequences_lockstep_on_device(syn_model, syn_gmm, tmp_path):
    """seqio.fit_sequences (is_seq = True, main.py:76-79 + non_linear_solver.py:157-162) with a real context: two sequences in
    lock-step; first frames = what fit_sequence (cold) gives, later frames warm-started (fewer iterations, finite results)"""
    import pickle
    from mvsmplfitting_b200 import seqio
    from mvsmplfitting_b200.context import FittingContext
    cams = S.make_cameras(4)
    T = 3

    def make_seq(seed, serial):
        fr = S.make_frames(syn_model, cams, T, seed=seed)
        return dict(gt_uv=fr["gt_uv"], conf=fr["conf"], present=np.ones((4, T), bool), cameras=["c%d" % v for v in range(4)],
                    frames=["%05d" % (t + 1) for t in range(T)], serial=serial)

    seqs = [make_seq(31, "0001"), make_seq(32, "0002")]

    def ctx_for(B):
        ctx = FittingContext(0)
        ctx.set_model(syn_model); ctx.set_gmm_from_dict(syn_gmm)
        ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
        return ctx

    ctx = ctx_for(2)
    sw = S.STAGE_WEIGHTS
    stages = [ctx.make_loss_config(body_prior="gmm", data_weight=500.0 / 1536, body_pose_weight=sw["body_pose_prior_weights"][i],
                                   shape_weight=sw["shape_weights"][i], bending_prior_weight=3.17 * sw["body_pose_prior_weights"][i])
              for i in range(4)]
    opt = ctx.make_lbfgs_config(max_outer=10)
    res, tot = seqio.fit_sequences(ctx, seqs, stages, opt, estimate_scale=True, result_folder=str(tmp_path / "res"), pose_format="lsp14")
    ctx.close()
    assert tot["frames_nan"] == 0 and all(np.isfinite(x).all() and np.isfinite(l).all() for x, l in res)
    # first frames: identical to the cold whole-sequence driver on the same frames
    for s_, q in enumerate(seqs):
        c1 = ctx_for(T)
        stages1 = [c1.make_loss_config(body_prior="gmm", data_weight=500.0 / 1536, body_pose_weight=sw["body_pose_prior_weights"][i],
                                       shape_weight=sw["shape_weights"][i], bending_prior_weight=3.17 * sw["body_pose_prior_weights"][i])
                   for i in range(4)]
        xc, lc, stc = seqio.fit_sequence(c1, q, stages1, c1.make_lbfgs_config(max_outer=10), estimate_scale=True, pose_format="lsp14")
        c1.close()
        assert np.array_equal(xc[0], res[s_][0][0]) and lc[0] == res[s_][1][0]
        assert not np.array_equal(xc[1], res[s_][0][1])                     # later frames took the warm path
        got = pickle.load(open(tmp_path / "res" / q["serial"] / q["frames"][2] / "000.pkl", "rb"))
        assert np.array_equal(got["transl"][0], res[s_][0][2][82:85])
