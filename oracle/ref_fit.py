"""TEST / MEASUREMENT INFRASTRUCTURE -- one frame fitted by the UNMODIFIED reference: its own
`utils/non_linear_solver.non_linear_solver(setting, data, **args)` (code/utils/non_linear_solver.py:37-288: per stage a
new LBFGSLs optimiser, loss.reset_loss_weights, create_fitting_closure, FittingMonitor.run_fitting) over the synthetic
SMPL-shaped model, cameras and GMM prior.  Nothing of the algorithm is restated here: this file only builds the `setting`
/ `data` dictionaries that code/init.py + code/main.py would build, and counts what the metric counts (L-BFGS iterations
= state['n_iter'] of every optimiser the solver created, lbfgs_ls.py:307; closure evaluations = state['func_evals']).

Used by bench.py --impl reference (CPU arm, and --ref-device cuda) and by the end-to-end fit fixtures / tests.
device="cuda" + interpenetration=True needs the reference's SDF kernel (oracle/_ref/libsdf_refcuda.so, oracle/ref_sdf.py).
"""
from __future__ import annotations

import contextlib
import io
import os

import numpy as np
import torch

from . import ref_harness as RH

_SCENE = {}


def stage_args(stage_weights: dict, interpenetration: bool, use_vposer: bool = False) -> dict:
    """the keyword arguments main.py passes on from cfg_files/fit_smpl.yaml (weights from mvsmplfitting_b200.synthetic);
    interactive=True because the solver only builds its result dict in that branch (non_linear_solver.py:274-288)"""
    return dict(batch_size=1, data_weights=list(stage_weights["data_weights"]),
                body_pose_prior_weights=list(stage_weights["body_pose_prior_weights"]),
                shape_weights=list(stage_weights["shape_weights"]), coll_loss_weights=list(stage_weights["coll_loss_weights"]),
                use_joints_conf=True, use_3d=False, rho=float(stage_weights["rho"]), interpenetration=bool(interpenetration),
                loss_type="smplify", visualize=False, use_vposer=bool(use_vposer), interactive=True, is_seq=False,
                optim_type="lbfgsls", lr=1.0, maxiters=int(stage_weights["maxiters"]), ftol=float(stage_weights["ftol"]),
                gtol=float(stage_weights["gtol"]))


def build_scene(model: dict, gmm: dict, cams: dict, device: str = "cpu", dtype=torch.float32):
    """reference modules for one (model, prior, cameras, device): built once per process"""
    key = (id(model), id(gmm), cams["R"].shape[0], device)
    if key in _SCENE:
        return _SCENE[key]
    ns = RH.import_reference()
    with RH.in_reference_dir():
        from utils import non_linear_solver as nls           # noqa: the reference's own driver of the path
    dev = torch.device(device)
    ref_model = RH.build_reference_model(model, dtype=dtype).to(dev)
    ref_cams = [c.to(dev) for c in RH.build_reference_cameras(cams, dtype=dtype)]
    prior = RH.build_reference_gmm(gmm, dtype=dtype).to(dev)
    sc = dict(ns=ns, nls=nls, model=ref_model, cams=ref_cams, body_pose_prior=prior,
              shape_prior=ns.prior.create_prior("l2"), angle_prior=ns.prior.create_prior("angle", dtype=dtype).to(dev),
              device=dev, dtype=dtype)
    _SCENE[key] = sc
    return sc


def fit_frame(sc: dict, frames: dict, b: int, stage_weights: dict, interpenetration: bool = False, image_height: int = 1536,
              quiet: bool = True, vposer=None) -> dict:
    """non_linear_solver on frame b of `frames` (synthetic.make_frames layout), starting from frames['init'][b].
    vposer: a loaded reference VPoser (RH.load_reference_vposer) -> use_vposer=True as in cfg_files/fit_smpl.yaml: L2 body prior
    (body_prior_type 'l2'), latent code from zero (init_guess.py:96-98)"""
    ns, nls, dev, dtype = sc["ns"], sc["nls"], sc["device"], sc["dtype"]
    init = frames["init"]
    sc["model"].reset_params(**{k: torch.tensor(np.asarray(init[k][b:b + 1]), dtype=dtype, device=dev)
                                for k in ("betas", "global_orient", "body_pose", "transl", "scale")})
    V = frames["gt_uv"].shape[0]
    kp = np.concatenate([frames["gt_uv"][:, b:b + 1], frames["conf"][:, b:b + 1, :, None]], axis=-1)      # [V,1,17,3]
    emb = None
    if vposer is not None:
        vposer = vposer.to(dev)
        emb = torch.zeros([1, 32], dtype=dtype, device=dev, requires_grad=True)
    setting = dict(views=V, device=dev, dtype=dtype, vposer=vposer, model=sc["model"], camera=sc["cams"], pose_embedding=emb,
                   joints_weight=torch.tensor(frames["joint_weights"], dtype=dtype, device=dev).unsqueeze(0), seq_start=True,
                   body_pose_prior=sc["body_pose_prior"] if vposer is None else ns.prior.create_prior("l2"),
                   shape_prior=sc["shape_prior"], angle_prior=sc["angle_prior"], adjustment=False)
    data = {"keypoints": kp.astype(np.float32), "3d_joint": None, "img": [np.zeros((image_height, 2, 3), np.uint8)],
            "img_path": None}
    # count what the metric counts without touching the solver: remember every optimiser it creates
    made = []
    factory = ns.optim_factory.create_optimizer

    def recording_factory(params, **kw):
        out = factory(params, **kw)
        made.append(out[0])
        return out
    ns.optim_factory.create_optimizer = recording_factory
    try:
        with (contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()), \
                (contextlib.redirect_stderr(io.StringIO()) if quiet else contextlib.nullcontext()):
            result = nls.non_linear_solver(setting, data, use_cuda=(dev.type == "cuda"),
                                           **stage_args(stage_weights, interpenetration, vposer is not None))
    finally:
        ns.optim_factory.create_optimizer = factory
    iters = evals = 0
    per_stage = []
    for opt in made:
        st = opt.state[opt._params[0]]
        it, ev = int(st.get("n_iter", 0)), int(st.get("func_evals", 0))
        per_stage.append((it, ev))
        iters += it
        evals += ev
    params = np.concatenate([sc["model"].__getattr__(k).detach().cpu().numpy().reshape(-1)
                             for k in ("betas", "global_orient", "body_pose", "transl", "scale")]).astype(np.float32)
    final = None if result is None else result.get("loss")
    return dict(iterations=iters, evals=evals, per_stage=per_stage, params=params,
                final_loss=float("nan") if final is None else float(final),
                pose_embedding=None if emb is None else emb.detach().cpu().numpy().reshape(-1))
