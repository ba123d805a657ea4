"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference as the ground truth that pins the
oracle.  The tree is /root/reference in the authoring container (used by oracle/make_golden*.py
to write tests/golden/*.npz and by tests marked `needs_reference`) or, on the GPU box, the copy
oracle/stage_reference.py placed under the git-ignored oracle/_ref/reference/ (bench.py --impl
reference and the -m gpu tests that run the reference closure with its own SDF kernel).

Nothing here is imported by the product package.

Recipe follows SURVEY.md section 8(c): chdir into the reference (its SMPL class
loads data/J_regressor_lsp.npz by relative path, body_models_scale.py:284),
register empty stub modules for the GUI / rendering imports the fitting path
never calls, and build the SMPL module through its `data_struct=` hook
(body_models_scale.py:98,169-180) because the licensed pickle is absent.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def _pick_root() -> str:
    env = os.environ.get("MVS_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/code/smplx"):
        return "/root/reference"
    return _STAGED


REF_ROOT = _pick_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "code", "smplx"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_imported = {}


def import_reference():
    """Returns a namespace with the reference modules of the hot path."""
    if _imported:
        return _imported["ns"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _stub("pyrender")
    _stub("pyrender.constants", RenderFlags=object)
    _stub("trimesh")
    _stub("OpenGL")
    _stub("OpenGL.GLUT", __all__=[])
    _stub("torchgeometry")
    _stub("configargparse")
    code = os.path.join(REF_ROOT, "code")
    if code not in sys.path:
        sys.path.insert(0, code)
    sdf_pkg = os.path.join(REF_ROOT, "sdf")                  # the reference's `sdf` package (sdf/sdf/sdf.py), unmodified;
    if sdf_pkg not in sys.path:                              # its compiled half `sdf.csrc` = the reference kernel built by
        sys.path.insert(0, sdf_pkg)                          # oracle/build_ref_sdf.sh (oracle/ref_sdf.py)
    from oracle import ref_sdf
    if ref_sdf.available():
        ref_sdf.install_csrc_stub()
    with in_reference_dir():
        import smplx  # noqa: F401  (reference code/smplx)
        from smplx import body_models_scale, lbs
        import camera
        import prior
        from utils import fitting
        from utils import utils as ref_utils
        from optimizers import optim_factory, lbfgs_ls
    ns = types.SimpleNamespace(
        smplx=sys.modules["smplx"], body_models_scale=body_models_scale, lbs=lbs,
        camera=camera, prior=prior, fitting=fitting, utils=ref_utils,
        optim_factory=optim_factory, lbfgs_ls=lbfgs_ls)
    _imported["ns"] = ns
    return ns


@contextlib.contextmanager
def in_reference_dir():
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(cwd)


def build_reference_model(model: dict, batch_size: int = 1, dtype=torch.float32,
                          model_type: str = "smpllsp", create_body_pose: bool = True):
    """Reference SMPL nn.Module over the synthetic `data_struct`."""
    ns = import_reference()
    from smplx.utils import Struct
    ds = Struct(f=model["f"], v_template=model["v_template"], shapedirs=model["shapedirs"],
                posedirs=model["posedirs"], J_regressor=model["J_regressor"],
                kintree_table=model["kintree_table"], weights=model["weights"])
    maps = ns.utils.smpl_to_annotation(
        model_type=model_type, pose_format="lsp14" if model_type == "smpllsp" else "coco17")
    mapper = ns.utils.JointMapper(maps)
    with in_reference_dir():
        m = ns.body_models_scale.SMPL(
            "unused", data_struct=ds, joint_mapper=mapper, model_type=model_type,
            create_global_orient=True, create_body_pose=create_body_pose, create_betas=True,
            create_transl=True, create_scale=True, dtype=dtype, batch_size=batch_size)
    return m


def build_reference_cameras(cams: dict, dtype=torch.float32):
    """list[PerspectiveCamera] exactly as reference init.py:108-131 builds them."""
    ns = import_reference()
    out = []
    for v in range(cams["R"].shape[0]):
        cam = ns.camera.create_camera(
            focal_length_x=float(cams["f"][v, 0]), focal_length_y=float(cams["f"][v, 1]),
            translation=torch.tensor(cams["t"][v], dtype=dtype).unsqueeze(0),
            rotation=torch.tensor(cams["R"][v], dtype=dtype).unsqueeze(0),
            center=torch.tensor(cams["c"][v], dtype=dtype).unsqueeze(0), dtype=dtype)
        cam.rotation.requires_grad = False
        cam.translation.requires_grad = False
        out.append(cam)
    return out


def build_reference_gmm(gmm: dict, dtype=torch.float32, tmpdir: str = "/tmp/mvs_gmm"):
    """MaxMixturePrior loaded from a pickle in the dict format (prior.py:130-133)."""
    import pickle
    ns = import_reference()
    os.makedirs(tmpdir, exist_ok=True)
    M = gmm["means"].shape[0]
    # written under a private name and renamed: many worker processes (bench.py --impl reference) build the same file at once
    dst = os.path.join(tmpdir, "gmm_%02d.pkl" % M)
    tmp = dst + ".%d.tmp" % os.getpid()
    with open(tmp, "wb") as f:
        pickle.dump({k: np.asarray(v) for k, v in gmm.items()}, f)
    os.replace(tmp, dst)
    return ns.prior.create_prior("gmm", prior_folder=tmpdir, num_gaussians=M, dtype=dtype)


def set_model_params(ref_model, params: dict, frame: int):
    with torch.no_grad():
        for k in ("betas", "global_orient", "body_pose", "transl", "scale"):
            if hasattr(ref_model, k):
                getattr(ref_model, k).copy_(
                    torch.as_tensor(params[k][frame:frame + 1], dtype=getattr(ref_model, k).dtype))
                getattr(ref_model, k).grad = None


def reference_closure_eval(ref_model, ref_cams, frames: dict, frame: int, weights: dict,
                           body_pose_prior, dtype=torch.float32, params: dict | None = None,
                           use_joints_conf=True, rho=100.0, fix_shape=False, device="cpu",
                           interpenetration=False, coll_loss_weight=0.0):
    """One fitting_func() of the reference for frame `frame` (fitting.py:162-203),
    B = 1 as the reference requires.  Returns dict(loss, grads{...}, joints, proj, vertices).
    device="cuda": the reference's shipped mode (cfg_files/fit_smpl.yaml:19); interpenetration=True additionally needs the
    reference's SDF kernel (oracle/_ref/libsdf_refcuda.so) and runs fitting.py:352-393 as written."""
    ns = import_reference()
    dev = torch.device(device)
    if next(ref_model.parameters()).device != dev:
        ref_model.to(dev)
        for cam in ref_cams:
            cam.to(dev)
        if body_pose_prior is not None and hasattr(body_pose_prior, "to"):
            body_pose_prior.to(dev)
    p = params if params is not None else frames["init"]
    set_model_params(ref_model, p, frame)
    V = frames["gt_uv"].shape[0]
    gt = torch.tensor(frames["gt_uv"][:, frame:frame + 1], dtype=dtype, device=dev)          # [V,1,17,2]
    conf = [torch.tensor(frames["conf"][v, frame:frame + 1], dtype=dtype, device=dev) for v in range(V)]
    jw = torch.tensor(frames["joint_weights"], dtype=dtype, device=dev).unsqueeze(0)
    loss = ns.fitting.create_loss(
        "smplify", rho=rho, use_joints_conf=use_joints_conf, dtype=dtype,
        body_pose_prior=body_pose_prior, shape_prior=ns.prior.create_prior("l2"),
        angle_prior=ns.prior.create_prior("angle", dtype=dtype),
        interpenetration=interpenetration, fix_shape=fix_shape).to(dev)
    wd = dict(weights)
    if interpenetration:
        wd["coll_loss_weight"] = coll_loss_weight
    loss.reset_loss_weights({k: torch.tensor(v, dtype=dtype, device=dev) for k, v in wd.items()})
    monitor = ns.fitting.FittingMonitor(maxiters=30, ftol=1e-9, gtol=1e-9)
    plist = [q for q in ref_model.parameters() if q.requires_grad]
    opt = torch.optim.SGD(plist, lr=0.0)
    closure = monitor.create_fitting_closure(
        opt, ref_model, camera=ref_cams, gt_joints=gt, joints_conf=conf, joint_weights=jw,
        loss=loss, create_graph=False, use_vposer=False, vposer=None, pose_embedding=None,
        return_verts=True, return_full_pose=True, use_3d=False)
    total = closure()
    out = ref_model(return_verts=True, return_full_pose=True)
    proj = torch.stack([cam(out.joints) for cam in ref_cams])[:, 0]
    return dict(
        loss=float(total),
        grads={k: getattr(ref_model, k).grad.detach().cpu().numpy().copy().reshape(-1)
               for k in ("betas", "global_orient", "body_pose", "transl", "scale")
               if getattr(ref_model, k).grad is not None},
        joints=out.joints.detach().cpu().numpy()[0].copy(),
        proj=proj.detach().cpu().numpy().copy(),
        vertices=out.vertices.detach().cpu().numpy()[0].copy(),
    )


def load_reference_vposer(device="cpu"):
    """The reference's own VPoser (code/model/VPoser.py) with the snapshot it ships (priors/snapshots/poser_epoch091.pkl),
    through its own loader code/utils/prior.py:23-54.  The pickle is a whole legacy module (torch 0.4 era): loading it needs
    `weights_only=False` and a stand-in for the long-gone torch.nn.backends.thnn."""
    import functools
    ns = import_reference()
    _stub("configer", Configer=object)
    thnn = _stub("torch.nn.backends.thnn", _get_thnn_function_backend=lambda: None)
    if not hasattr(torch.nn, "backends"):
        torch.nn.backends = _stub("torch.nn.backends")
    torch.nn.backends.thnn = thnn
    with in_reference_dir():
        from utils import prior as ref_prior_utils
        orig_load = torch.load
        torch.load = functools.partial(orig_load, weights_only=False, map_location="cpu")      # the snapshot was saved from a CUDA device
        try:
            import contextlib as _cl, io as _io
            with _cl.redirect_stdout(_io.StringIO()):
                vp = ref_prior_utils.load_vposer(os.path.join(REF_ROOT, "priors"))
        finally:
            torch.load = orig_load
    return vp.to(device)


def vposer_weights_numpy(vp) -> dict:
    g = lambda layer, what: getattr(getattr(vp, layer), what).detach().float().cpu().numpy()
    return dict(fc1_w=g("bodyprior_dec_fc1", "weight"), fc1_b=g("bodyprior_dec_fc1", "bias"),
                fc2_w=g("bodyprior_dec_fc2", "weight"), fc2_b=g("bodyprior_dec_fc2", "bias"),
                out_w=g("bodyprior_dec_out", "weight"), out_b=g("bodyprior_dec_out", "bias"))
