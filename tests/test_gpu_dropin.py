"""GPU: the reference-facing Python surface (create_scale / create_camera / create_prior / create_loss /
create_optimizer / FittingMonitor) driven exactly like code/utils/non_linear_solver.py:156-203 drives the
reference, against the reference-run golden fixtures."""
import pickle

import numpy as np
import pytest
import torch

from mvsmplfitting_b200 import synthetic as S
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def build_scene(syn_model, syn_gmm, tmp_path, cams, B, model_type="smpllsp", body_prior="gmm"):
    from mvsmplfitting_b200 import camera, prior, smplx
    from mvsmplfitting_b200.utils import utils
    dev = torch.device("cuda")
    ds = smplx.Struct(f=syn_model["f"], v_template=syn_model["v_template"], shapedirs=syn_model["shapedirs"],
                      posedirs=syn_model["posedirs"], J_regressor=syn_model["J_regressor"],
                      kintree_table=syn_model["kintree_table"], weights=syn_model["weights"])
    fmt = "lsp14" if model_type == "smpllsp" else "coco17"
    model = smplx.create_scale("unused", model_type=model_type, data_struct=ds,
                               joint_mapper=utils.JointMapper(utils.smpl_to_annotation(model_type, pose_format=fmt)),
                               batch_size=B, dtype=torch.float32).to(dev)
    cam_list = []
    for v in range(cams["R"].shape[0]):
        c = camera.create_camera(focal_length_x=float(cams["f"][v, 0]), focal_length_y=float(cams["f"][v, 1]),
                                 translation=torch.tensor(cams["t"][v]).unsqueeze(0),
                                 rotation=torch.tensor(cams["R"][v]).unsqueeze(0),
                                 center=torch.tensor(cams["c"][v]).unsqueeze(0)).to(dev)
        c.rotation.requires_grad = False
        c.translation.requires_grad = False
        cam_list.append(c)
    with open(tmp_path / "gmm_06.pkl", "wb") as f:
        pickle.dump({k: np.asarray(v) for k, v in syn_gmm.items()}, f)
    bp = prior.create_prior(body_prior, prior_folder=str(tmp_path), num_gaussians=6, dtype=torch.float32)
    if hasattr(bp, "to"):
        bp = bp.to(dev)
    return model, cam_list, bp


@pytest.mark.parametrize("name", ["gmm8_s3", "l2_4_s3", "smpl_coco", "l2_4_fixshape"])
def test_closure_through_the_reference_surface(name, syn_model, syn_gmm, tmp_path):
    from mvsmplfitting_b200 import fitting, prior
    from mvsmplfitting_b200.optimizers import optim_factory
    c = G.load_case(name)
    B = 1
    model, cams, bp = build_scene(syn_model, syn_gmm, tmp_path, c["cams"], B, c["meta"]["model_type"], c["meta"]["body_prior"])
    x = S.unpack_params(c["X"][:B])
    model.reset_params(**{k: torch.tensor(v) for k, v in x.items()})
    if c["meta"]["fix_shape"]:
        model.betas.requires_grad = False
    loss = fitting.create_loss("smplify", rho=100.0, use_joints_conf=c["meta"]["use_joints_conf"], body_pose_prior=bp,
                               shape_prior=prior.create_prior("l2"), angle_prior=prior.create_prior("angle"),
                               interpenetration=False, fix_shape=c["meta"]["fix_shape"]).to("cuda")
    loss.reset_loss_weights({k: torch.tensor(v) for k, v in c["w"].items()})
    params = [p for p in model.parameters() if p.requires_grad]
    opt, cg = optim_factory.create_optimizer(params, optim_type="lbfgsls", lr=1.0, maxiters=30)
    mon = fitting.FittingMonitor(maxiters=30, ftol=1e-9, gtol=1e-9)
    V = c["gt_uv"].shape[0]
    closure = mon.create_fitting_closure(
        opt, model, camera=cams, gt_joints=torch.tensor(c["gt_uv"][:, :B]).cuda(),
        joints_conf=[torch.tensor(c["conf"][v, :B]).cuda() for v in range(V)],
        joint_weights=torch.tensor(c["joint_weights"]).unsqueeze(0).cuda(), loss=loss, create_graph=cg,
        use_vposer=False, vposer=None, pose_embedding=None, return_verts=True, return_full_pose=True, use_3d=False)
    total = closure()
    assert abs(float(total) - c["loss_f32"][0]) / c["loss_f32"][0] < 1e-4
    for k, (a, e) in fitting.PARAM_SLICES.items():
        p = getattr(model, k)
        if not p.requires_grad:
            assert p.grad is None
            continue
        g_ref = c["grad_f32"][0, a:e]
        assert G.relmax(p.grad.cpu().numpy().reshape(-1), g_ref) < 1e-4, k
    # SMPL.forward outside the closure
    out = model(return_verts=True, return_full_pose=True)
    assert G.relmax(out.joints.cpu().numpy(), c["joints_f32"][:B]) < 1e-4
    assert G.relmax(out.vertices.cpu().numpy()[:, :64], c["verts_head_f32"][:B]) < 1e-4
    assert tuple(out.full_pose.shape) == (1, 72)
    # the fused run_fitting moves the caller's parameters and lowers the loss
    before = {k: getattr(model, k).detach().clone() for k in fitting.PARAM_SLICES}
    final = mon.run_fitting(opt, closure, params, model, use_vposer=False)
    assert final < float(total)
    assert float(closure(backward=False)) <= final * (1 + 1e-3)
    if c["meta"]["fix_shape"]:
        assert torch.equal(model.betas, before["betas"])
    assert not torch.equal(model.body_pose, before["body_pose"])
    assert mon.last_stats["frame_evals"] > 5


def test_step_by_step_driving_matches_fused_run(syn_model, syn_gmm, tmp_path):
    """optimizer.step(closure) called from a host loop (what the reference's own run_fitting does) walks the
    same iterates as the fused device loop"""
    from mvsmplfitting_b200 import fitting, prior
    from mvsmplfitting_b200.optimizers import optim_factory
    cams = S.make_cameras(4)
    B = 3
    fr = S.make_frames(syn_model, cams, B, seed=8)
    w = dict(data_weight=500.0 / 1536, body_pose_weight=4.78, shape_weight=5.0, bending_prior_weight=3.17 * 4.78)

    def run(fused):
        model, cam_list, bp = build_scene(syn_model, syn_gmm, tmp_path, cams, B)
        model.reset_params(**{k: torch.tensor(v) for k, v in fr["init"].items()})
        loss = fitting.create_loss("smplify", rho=100.0, body_pose_prior=bp, shape_prior=prior.create_prior("l2"),
                                   angle_prior=prior.create_prior("angle"), interpenetration=False, fix_shape=False).to("cuda")
        loss.reset_loss_weights(w)
        params = [p for p in model.parameters() if p.requires_grad]
        opt, _ = optim_factory.create_optimizer(params, optim_type="lbfgsls", lr=1.0, maxiters=30)
        mon = fitting.FittingMonitor(maxiters=3, ftol=0.0, gtol=0.0)
        closure = mon.create_fitting_closure(opt, model, camera=cam_list, gt_joints=torch.tensor(fr["gt_uv"]).cuda(),
                                             joints_conf=[torch.tensor(fr["conf"][v]).cuda() for v in range(4)],
                                             joint_weights=torch.tensor(fr["joint_weights"]).unsqueeze(0).cuda(), loss=loss)
        closure.ctx.set_exec_mode(1)      # batched kernels on both sides: the step-wise API has no frame-resident form
        if fused:
            mon.run_fitting(opt, closure, params, model, use_vposer=False)
        else:
            for _ in range(3):
                opt.step(closure)
                assert model.body_pose.grad is not None
        return torch.cat([p.detach().reshape(B, -1) for p in model.parameters()], dim=1)
    a, b = run(True), run(False)
    assert torch.equal(a, b)


def test_run_fitting_stages_equals_stage_loop(syn_model, syn_gmm, tmp_path):
    """FittingMonitor.run_fitting_stages (mvs_fit: frames change stage on their own) against the reference-style loop
    'reset_loss_weights -> new optimiser -> run_fitting' over the same stage weights: identical parameters."""
    from mvsmplfitting_b200 import fitting, prior
    from mvsmplfitting_b200.optimizers import optim_factory
    c = G.load_case("gmm8_s3")
    B = 1
    stage_w = [{k: torch.tensor(v) for k, v in c["w"].items()} for _ in range(2)]
    stage_w[1]["body_pose_weight"] = stage_w[1]["body_pose_weight"] * 0.25
    stage_w[1]["shape_weight"] = stage_w[1]["shape_weight"] * 0.5
    results = []
    for merged in (False, True):
        model, cams, bp = build_scene(syn_model, syn_gmm, tmp_path, c["cams"], B, c["meta"]["model_type"], c["meta"]["body_prior"])
        x = S.unpack_params(c["X"][:B])
        model.reset_params(**{k: torch.tensor(v) for k, v in x.items()})
        loss = fitting.create_loss("smplify", rho=100.0, use_joints_conf=c["meta"]["use_joints_conf"], body_pose_prior=bp,
                                   shape_prior=prior.create_prior("l2"), angle_prior=prior.create_prior("angle"),
                                   interpenetration=False, fix_shape=False).to("cuda")
        params = [p for p in model.parameters() if p.requires_grad]
        mon = fitting.FittingMonitor(maxiters=5, ftol=1e-9, gtol=1e-9)
        V = c["gt_uv"].shape[0]

        def make(opt):
            return mon.create_fitting_closure(
                opt, model, camera=cams, gt_joints=torch.tensor(c["gt_uv"][:, :B]).cuda(),
                joints_conf=[torch.tensor(c["conf"][v, :B]).cuda() for v in range(V)],
                joint_weights=torch.tensor(c["joint_weights"]).unsqueeze(0).cuda(), loss=loss, create_graph=False,
                use_vposer=False, vposer=None, pose_embedding=None, return_verts=True, return_full_pose=True, use_3d=False)
        if merged:
            opt, _ = optim_factory.create_optimizer(params, optim_type="lbfgsls", lr=1.0, maxiters=30)
            mon.run_fitting_stages(opt, make(opt), stage_w)
        else:
            for w in stage_w:
                loss.reset_loss_weights(w)
                opt, _ = optim_factory.create_optimizer(params, optim_type="lbfgsls", lr=1.0, maxiters=30)
                mon.run_fitting(opt, make(opt), params, model, use_vposer=False)
        results.append({k: getattr(model, k).detach().cpu().clone() for k in fitting.PARAM_SLICES})
    for k in fitting.PARAM_SLICES:
        assert torch.equal(results[0][k], results[1][k]), k
