"""Python face of one libmvsmpl context (one per GPU / process).

PyTorch is used only as the owner of device memory and streams; every computation
on the path goes through the C ABI in include/mvsmpl.h.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from . import synthetic as S

PRIOR_CODES = {"l2": 0, "gmm": 1, "none": 2, None: 2}
SEGMENT_BITS = {"betas": 0, "global_orient": 1, "body_pose": 2, "transl": 3, "scale": 4}


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    return a.ctypes.data_as(ctypes.c_void_p)


def _stats_dict(st) -> dict:
    return dict(frame_iterations=st.frame_iterations, frame_evals=st.frame_evals, rounds=st.rounds, frames_nan=st.frames_nan,
                dense_frame_evals=st.dense_frame_evals, dense_rounds=st.dense_rounds)


class FittingContext:
    """Owns the device copies of the model / cameras / detections and launches the closure
    and the batched L-BFGS.  All tensors passed in must live on `device` and be contiguous fp32."""

    def __init__(self, device: int | torch.device = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.MvsError("no CUDA device visible: mvsmplfitting_b200 has no CPU fallback")
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        h = ctypes.c_void_p()
        _lib.check(None, self.lib.mvs_create(self.device.index, ctypes.byref(h)), "mvs_create")
        self.h = h
        self.B = 0
        self.V = 0
        self.K = 17
        self.N = 0
        self._loss_cfg = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.mvs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ uploads
    def set_model(self, model: dict, model_type: str = "smpllsp", joint_map=None, extra_vertex_ids=None):
        """`model` holds the reference's data_struct fields (v_template, shapedirs, posedirs [N,3,207],
        J_regressor, kintree_table, weights, f) plus `lsp_regressor` for model_type 'smpllsp'."""
        N = model["v_template"].shape[0]
        posedirs = np.asarray(model["posedirs"])
        if posedirs.ndim == 3:                      # [N,3,207] -> the registered [207, 3N] layout
            posedirs = np.reshape(posedirs, [-1, posedirs.shape[-1]]).T
        parents = np.asarray(model["kintree_table"][0]).astype(np.int64) if "kintree_table" in model \
            else np.asarray(model["parents"]).astype(np.int64)
        parents = parents.copy()
        parents[0] = -1
        if model_type == "smpllsp":
            reg = _f32(model["lsp_regressor"])
            n_reg = reg.shape[0]
            jmap = S.JOINT_MAP_LSP14 if joint_map is None else joint_map
        else:
            reg, n_reg = None, 0
            jmap = S.JOINT_MAP_COCO17_SMPL if joint_map is None else joint_map
        extra = _i32(S.FACE_VERTEX_IDS if extra_vertex_ids is None else extra_vertex_ids)
        faces = _i32(model["f"]) if model.get("f") is not None else None
        keep = dict(vt=_f32(model["v_template"]), sd=_f32(model["shapedirs"]), pd=_f32(posedirs),
                    jr=_f32(model["J_regressor"]), par=_i32(parents), w=_f32(model["weights"]), faces=faces,
                    reg=reg, extra=extra, jmap=_i32(jmap))
        d = _lib.ModelDesc(
            n_verts=N, n_faces=0 if faces is None else faces.shape[0],
            v_template=_ptr(keep["vt"]), shapedirs=_ptr(keep["sd"]), posedirs=_ptr(keep["pd"]),
            J_regressor=_ptr(keep["jr"]), parents=_ptr(keep["par"]), lbs_weights=_ptr(keep["w"]),
            faces=_ptr(faces), n_keypoints=len(keep["jmap"]), n_reg=n_reg, joint_regressor=_ptr(reg),
            n_extra=len(extra), extra_vertex_ids=_ptr(extra), joint_map=_ptr(keep["jmap"]))
        _lib.check(self.h, self.lib.mvs_set_model(self.h, ctypes.byref(d)), "mvs_set_model")
        self.N, self.K = N, len(keep["jmap"])
        self.model_type = model_type

    def set_gmm(self, means, precisions, nll_weights):
        m, p, w = _f32(means), _f32(precisions), _f32(np.reshape(nll_weights, -1))
        _lib.check(self.h, self.lib.mvs_set_gmm_prior(self.h, m.shape[0], _ptr(m), _ptr(p), _ptr(w)), "mvs_set_gmm_prior")

    def set_gmm_from_dict(self, gmm: dict):
        """same float32 buffer path as the reference (prior.py:127-160)"""
        means = gmm["means"].astype(np.float32)
        covs = gmm["covars"].astype(np.float32)
        prec = np.stack([np.linalg.inv(c) for c in covs]).astype(np.float32)
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm["covars"]])
        nllw = np.asarray(gmm["weights"] / ((2 * np.pi) ** (69 / 2.0) * (sqrdets / sqrdets.min())))
        self.set_gmm(means, prec, nllw.astype(np.float32))

    def set_vposer(self, weights: dict):
        """VPoser decoder weights (mvs_set_vposer): dict(fc1_w [512,32], fc1_b, fc2_w [512,512], fc2_b, out_w [138,512],
        out_b) -- e.g. the bodyprior_dec_* tensors of a loaded VPoser, or synthetic.make_vposer().  Enables
        use_vposer=2 loss configurations (pose decoded on the device from the latent code in the pose slot)."""
        arrs = [_f32(weights[k]) for k in ("fc1_w", "fc1_b", "fc2_w", "fc2_b", "out_w", "out_b")]
        assert arrs[0].shape == (512, 32) and arrs[2].shape == (512, 512) and arrs[4].shape == (138, 512)
        _lib.check(self.h, self.lib.mvs_set_vposer(self.h, *[_ptr(a) for a in arrs]), "mvs_set_vposer")

    def vposer_decode(self, params: torch.Tensor) -> torch.Tensor:
        """vposer.decode(z, 'aa') of the latent codes in params[:, 13:45] (mvs_vposer_decode) -> body_pose [B,69]"""
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous() and params.shape == (self.B, 86)
        out = torch.empty(self.B, 69, dtype=torch.float32, device=self.device)
        _lib.check(self.h, self.lib.mvs_vposer_decode(self.h, _ptr(params), _ptr(out), self._stream()), "mvs_vposer_decode")
        return out

    def set_cameras(self, R, t, f, c):
        R, t, f, c = _f32(R), _f32(t), _f32(f), _f32(c)
        self.V = R.shape[0]
        _lib.check(self.h, self.lib.mvs_set_cameras(self.h, self.V, _ptr(R), _ptr(t), _ptr(f), _ptr(c)), "mvs_set_cameras")

    def set_batch(self, B: int):
        _lib.check(self.h, self.lib.mvs_set_batch(self.h, int(B)), "mvs_set_batch")
        self.B = int(B)

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def set_keypoints(self, gt_uv, conf, joint_weights):
        """gt_uv [V,B,K,2], conf [V,B,K], joint_weights [K]; numpy (host) or CUDA tensors."""
        on_dev = isinstance(gt_uv, torch.Tensor) and gt_uv.is_cuda
        if on_dev:
            gt_uv = gt_uv.contiguous().float()
            conf = conf.to(self.device).contiguous().float()
            joint_weights = torch.as_tensor(joint_weights).to(self.device).contiguous().float().reshape(-1)
        else:
            gt_uv, conf, joint_weights = _f32(gt_uv), _f32(conf), _f32(np.reshape(joint_weights, -1))
        assert tuple(gt_uv.shape) == (self.V, self.B, self.K, 2), (tuple(gt_uv.shape), (self.V, self.B, self.K, 2))
        assert tuple(conf.shape) == (self.V, self.B, self.K)
        self._kp_keep = (gt_uv, conf, joint_weights)
        _lib.check(self.h, self.lib.mvs_set_keypoints(self.h, _ptr(gt_uv), _ptr(conf), _ptr(joint_weights),
                                                      1 if on_dev else 0, self._stream()), "mvs_set_keypoints")

    @staticmethod
    def make_loss_config(data_weight=1.0, body_pose_weight=0.0, shape_weight=0.0, bending_prior_weight=0.0,
                         coll_loss_weight=0.0, rho=100.0, body_prior="l2", use_joints_conf=True, use_vposer=False,
                         fix_shape=False, interpenetration=False, sdf_grid=128, sdf_all_faces=False,
                         frozen=()) -> _lib.LossConfig:
        mask = 0
        for name in frozen:
            mask |= 1 << SEGMENT_BITS[name]
        return _lib.LossConfig(float(data_weight), float(body_pose_weight), float(shape_weight),
                               float(bending_prior_weight), float(coll_loss_weight), float(rho),
                               PRIOR_CODES[body_prior], int(use_joints_conf), int(use_vposer), int(fix_shape),
                               int(interpenetration), int(sdf_grid), int(sdf_all_faces), mask)

    def set_loss(self, **kw):
        cfg = kw.pop("config", None) or self.make_loss_config(**kw)
        _lib.check(self.h, self.lib.mvs_set_loss_config(self.h, ctypes.byref(cfg)), "mvs_set_loss_config")
        self._loss_cfg = cfg

    # ------------------------------------------------------------------ compute
    def closure(self, params: torch.Tensor, want_grad=True, want_joints=False, want_proj=False, want_verts=False):
        """One batched fitting_func(): params [B,86] (CUDA fp32) -> dict(loss [B], grad [B,86], ...)."""
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
        assert tuple(params.shape) == (self.B, S.NUM_PARAMS)
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.device)
        out = dict(loss=new(self.B))
        if want_grad:
            out["grad"] = new(self.B, S.NUM_PARAMS)
        if want_joints:
            out["joints"] = new(self.B, self.K, 3)
        if want_proj:
            out["proj"] = new(self.V, self.B, self.K, 2)
        if want_verts:
            out["verts"] = new(self.B, self.N, 3)
        _lib.check(self.h, self.lib.mvs_closure(self.h, _ptr(params), _ptr(out["loss"]), _ptr(out.get("grad")),
                                                _ptr(out.get("joints")), _ptr(out.get("proj")), _ptr(out.get("verts")),
                                                self._stream()), "mvs_closure")
        return out

    def forward_only(self, params: torch.Tensor, want_verts: bool = True):
        """SMPL.forward outside the closure: dict(joints [B,K,3], verts [B,N,3])"""
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
        out = dict(joints=torch.empty(self.B, self.K, 3, dtype=torch.float32, device=self.device))
        if want_verts:
            out["verts"] = torch.empty(self.B, self.N, 3, dtype=torch.float32, device=self.device)
        _lib.check(self.h, self.lib.mvs_forward(self.h, _ptr(params), _ptr(out["joints"]), _ptr(out.get("verts")),
                                                self._stream()), "mvs_forward")
        return out

    @staticmethod
    def make_lbfgs_config(max_outer=30, max_iter=30, max_eval=None, history_size=100, lr=1.0, tolerance_grad=1e-5,
                          tolerance_change=1e-9, ftol=1e-9, gtol=1e-9) -> _lib.LbfgsConfig:
        return _lib.LbfgsConfig(int(max_outer), int(max_iter), int(max_eval if max_eval is not None else max_iter * 5 // 4),
                                int(history_size), float(lr), float(tolerance_grad), float(tolerance_change),
                                float(ftol), float(gtol))

    def lbfgs_run(self, params: torch.Tensor, config: _lib.LbfgsConfig | None = None):
        """Runs FittingMonitor.run_fitting for every frame on the device; params updated in place.
        Returns (final_loss [B] tensor, stats dict)."""
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
        cfg = config or self.make_lbfgs_config()
        final = torch.empty(self.B, dtype=torch.float32, device=self.device)
        st = _lib.LbfgsStats()
        _lib.check(self.h, self.lib.mvs_lbfgs_run(self.h, _ptr(params), _ptr(final), ctypes.byref(cfg), ctypes.byref(st),
                                                  self._stream()), "mvs_lbfgs_run")
        return final, _stats_dict(st)

    def lbfgs_step(self, params: torch.Tensor, config=None, reset: bool = False):
        """one LBFGS.step() for every frame; returns (loss at entry [B], grads of the last closure call [B,86], stats)"""
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
        cfg = config or self.make_lbfgs_config()
        loss = torch.empty(self.B, dtype=torch.float32, device=self.device)
        grad = torch.empty(self.B, S.NUM_PARAMS, dtype=torch.float32, device=self.device)
        st = _lib.LbfgsStats()
        _lib.check(self.h, self.lib.mvs_lbfgs_step(self.h, _ptr(params), _ptr(loss), _ptr(grad), ctypes.byref(cfg),
                                                   1 if reset else 0, ctypes.byref(st), self._stream()), "mvs_lbfgs_step")
        return loss, grad, _stats_dict(st)

    def init_guess(self, estimate_scale: bool = True, fixed_scale: float = 1.0, use_torso: bool = True,
                   hip_seed: float = 1.0, want_joints3d: bool = True, umeyama_as_written: bool = False):
        """Initial parameter block of every frame from the uploaded detections (mvs_init_guess: triangulation -- or the
        single-view depth guess -- + similarity alignment, init_guess.py:18-107 + fix_params :190-212).
        umeyama_as_written: evaluate code/utils/umeyama.py as written (see include/mvsmpl.h) instead of the published
        algorithm.  Returns (params [B,86], joints3d [B,K,3])."""
        params = torch.empty(self.B, 86, dtype=torch.float32, device=self.device)
        j3 = torch.empty(self.B, self.K, 3, dtype=torch.float32, device=self.device) if want_joints3d else None
        cfg = _lib.InitConfig(int(bool(estimate_scale)), float(fixed_scale), int(bool(use_torso)), float(hip_seed),
                              int(bool(umeyama_as_written)))
        _lib.check(self.h, self.lib.mvs_init_guess(self.h, _ptr(params), _ptr(j3), ctypes.byref(cfg), self._stream()),
                   "mvs_init_guess")
        return params, j3

    def fit(self, params: torch.Tensor, stage_cfgs, opt_cfg=None, warm=None):
        """All stages of a fit on device buffers (mvs_fit): params [B,86] CUDA float32, updated in place.  Frames move
        from stage to stage on their own; returns (final_loss [B], stats).  warm [B] bool (mvs_fit_seq): frames that
        continue a video sequence from the previous frame's result skip stages 0, 1 and damp stage 2's pose prior
        (non_linear_solver.py:157-162)."""
        assert params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
        arr = (_lib.LossConfig * len(stage_cfgs))(*stage_cfgs)
        cfg = opt_cfg or self.make_lbfgs_config()
        final = torch.empty(self.B, dtype=torch.float32, device=self.device)
        st = _lib.LbfgsStats()
        if warm is not None:
            wm = np.ascontiguousarray(np.asarray(warm).astype(np.uint8).reshape(self.B))
            _lib.check(self.h, self.lib.mvs_fit_seq(self.h, _ptr(params), len(stage_cfgs), arr, ctypes.byref(cfg),
                                                    wm.ctypes.data_as(ctypes.c_void_p), _ptr(final), ctypes.byref(st),
                                                    self._stream()), "mvs_fit_seq")
        else:
            _lib.check(self.h, self.lib.mvs_fit(self.h, _ptr(params), len(stage_cfgs), arr, ctypes.byref(cfg), _ptr(final),
                                                ctypes.byref(st), self._stream()), "mvs_fit")
        return final, _stats_dict(st)

    def fit_host(self, params_host: np.ndarray, gt_uv: np.ndarray, conf: np.ndarray, joint_weights: np.ndarray,
                 stage_cfgs, opt_cfg=None):
        """Host-buffer entry point (mvs_fit_host): params_host [B,86] float32 updated in place."""
        assert params_host.dtype == np.float32 and params_host.flags["C_CONTIGUOUS"]
        gt_uv, conf, jw = _f32(gt_uv), _f32(conf), _f32(np.reshape(joint_weights, -1))
        arr = (_lib.LossConfig * len(stage_cfgs))(*stage_cfgs)
        cfg = opt_cfg or self.make_lbfgs_config()
        final = np.zeros(self.B, dtype=np.float32)
        st = _lib.LbfgsStats()
        _lib.check(self.h, self.lib.mvs_fit_host(self.h, _ptr(params_host), _ptr(gt_uv), _ptr(conf), _ptr(jw),
                                                 len(stage_cfgs), arr, ctypes.byref(cfg), _ptr(final), ctypes.byref(st),
                                                 self._stream()), "mvs_fit_host")
        return final, _stats_dict(st)

    def sdf_grid(self, faces: torch.Tensor, verts: torch.Tensor, grid_size: int, num_faces: int | None = None):
        """The reference's `sdf.csrc.sdf` op: phi [B,G,G,G] for verts [B,N,3] normalised to [-1,1]."""
        assert verts.is_cuda and faces.is_cuda
        faces = faces.to(torch.int32).contiguous()
        verts = verts.float().contiguous()
        Bv, Nv = verts.shape[0], verts.shape[1]
        phi = torch.zeros(Bv, grid_size, grid_size, grid_size, dtype=torch.float32, device=self.device)
        nf = int(num_faces if num_faces is not None else faces.reshape(-1, 3).shape[0])
        _lib.check(self.h, self.lib.mvs_sdf_grid(self.h, _ptr(phi), _ptr(faces), nf, _ptr(verts), Bv, Nv, grid_size,
                                                 self._stream()), "mvs_sdf_grid")
        return phi

    def set_anchor(self, anchor: torch.Tensor | None, weight: torch.Tensor | None = None):
        """per-frame quadratic anchor sum_i w_i (x_i - a_i)^2 (sequence mode); anchor=None switches it off.
        Takes effect with the next set_loss()."""
        if anchor is None:
            _lib.check(self.h, self.lib.mvs_set_anchor(self.h, None, None, 0, self._stream()), "mvs_set_anchor")
            return
        anchor = anchor.to(self.device, torch.float32).contiguous()
        weight = weight.to(self.device, torch.float32).contiguous()
        if weight.dim() == 1:
            weight = weight[None].expand(self.B, -1).contiguous()
        assert tuple(anchor.shape) == (self.B, S.NUM_PARAMS) and tuple(weight.shape) == (self.B, S.NUM_PARAMS)
        _lib.check(self.h, self.lib.mvs_set_anchor(self.h, _ptr(anchor), _ptr(weight), 1, self._stream()), "mvs_set_anchor")

    def set_exec_mode(self, mode: int):
        """0 = frame-resident kernels where they apply (default), 1 = batched kernels only"""
        _lib.check(self.h, self.lib.mvs_set_exec_mode(self.h, int(mode)), "mvs_set_exec_mode")

    def profile(self, mask: int = 0xFFFFFFFF):
        """start (mask != 0) / stop (0) per-kernel CUDA-event timing (see mvs_profile in mvsmpl.h)"""
        _lib.check(self.h, self.lib.mvs_profile(self.h, ctypes.c_uint(mask & 0xFFFFFFFF)), "mvs_profile")

    def profile_read(self) -> dict:
        """{kernel name: (total ms, launches)} since the last profile() call"""
        n = _lib.NUM_KERNEL_IDS
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_longlong * n)()
        _lib.check(self.h, self.lib.mvs_profile_read(self.h, ms, cnt), "mvs_profile_read")
        return {self.lib.mvs_kernel_name(k).decode(): (ms[k], int(cnt[k])) for k in range(n) if cnt[k]}

    def launch_count(self) -> int:
        return int(self.lib.mvs_launch_count(self.h))
