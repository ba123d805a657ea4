"""CPU: the mirror of the reference's utils/init_guess.py (init_guess / load_init / fix_params) drives the context and
the model exactly as the reference drives numpy and the model; the device call is replaced by a recording stub that
answers with the oracle's numbers."""
import numpy as np
import pytest
import torch

from mvsmplfitting_b200 import synthetic as S
from mvsmplfitting_b200.utils import init_guess as IG
from oracle import init_oracle as IO


def make_model(syn_model, B):
    from mvsmplfitting_b200 import smplx
    from mvsmplfitting_b200.utils import utils
    ds = smplx.Struct(f=syn_model["f"], v_template=syn_model["v_template"], shapedirs=syn_model["shapedirs"],
                      posedirs=syn_model["posedirs"], J_regressor=syn_model["J_regressor"],
                      kintree_table=syn_model["kintree_table"], weights=syn_model["weights"])
    return smplx.create_scale("unused", model_type="smpllsp", data_struct=ds, joint_mapper=utils.JointMapper(
        utils.smpl_to_annotation("smpllsp", pose_format="lsp14")), create_global_orient=True, create_body_pose=True,
        create_betas=True, create_transl=True, create_scale=True, dtype=torch.float32, batch_size=B,
        lsp_regressor_path="/nonexistent")


class StubCtx:
    """stands in for FittingContext: checks what it is given, answers like mvs_init_guess (from the oracle)"""

    def __init__(self, B, rest, ext, intr):
        self.B, self.rest, self.ext, self.intr, self.calls = B, rest, ext, intr, []

    def set_cameras(self, R, t, f, c):
        self.calls.append("cams")
        V = R.shape[0]
        assert np.allclose(R, self.ext[:V, :3, :3]) and np.allclose(t, self.ext[:V, :3, 3])
        assert np.allclose(f[:, 0], self.intr[:V, 0, 0]) and np.allclose(c[:, 1], self.intr[:V, 1, 2])

    def set_keypoints(self, gt_uv, conf, jw):
        self.calls.append("kp")
        self.gt_uv, self.conf = np.asarray(gt_uv), np.asarray(conf)
        assert self.gt_uv.shape[1] == self.B and self.gt_uv.shape[2:] == (17, 2) and jw.sum() == 17

    def init_guess(self, estimate_scale, fixed_scale, use_torso, hip_seed, umeyama_as_written=False):
        self.calls.append(("init", estimate_scale, fixed_scale, use_torso, hip_seed))
        self.as_written = umeyama_as_written
        V = self.gt_uv.shape[0]
        x = np.zeros((self.B, 86), np.float32)
        for b in range(self.B):
            kps = [np.concatenate([self.gt_uv[v, b], self.conf[v, b][:, None]], 1) for v in range(V)]
            o = IO.init_guess(self.ext[:V], self.intr[:V], kps, self.rest, estimate_scale, fixed_scale, use_torso,
                              as_written=umeyama_as_written, svd=IO.svd_sign_normalised)
            x[b, 10:13], x[b, 82:85], x[b, 85] = o["global_orient"], o["transl"], o["scale"]
        return torch.tensor(x), None


@pytest.mark.parametrize("B,fix_scale", [(1, False), (3, True)])
def test_init_guess_then_fix_params(B, fix_scale, syn_model, monkeypatch):
    V = 4
    cams = S.make_cameras(V)
    fr = S.make_frames(syn_model, cams, B, seed=40)
    ext = np.tile(np.eye(4), (V, 1, 1))
    ext[:, :3, :3], ext[:, :3, 3] = cams["R"], cams["t"]
    intr = np.tile(np.eye(3), (V, 1, 1))
    intr[:, 0, 0], intr[:, 1, 1], intr[:, 0, 2], intr[:, 1, 2] = cams["f"][:, 0], cams["f"][:, 1], cams["c"][:, 0], cams["c"][:, 1]
    z = lambda n: np.zeros((1, n))
    rest = S.model_keypoints_np(syn_model, z(10), z(3), z(69), z(3), np.ones((1, 1)), "smpllsp")[0]
    model = make_model(syn_model, B)
    with torch.no_grad():
        model.betas.fill_(0.3)
        model.body_pose.fill_(0.2)
    stub = StubCtx(B, rest, ext, intr)
    monkeypatch.setattr(IG, "model_context", lambda m: stub)
    emb = torch.ones(1, 32)
    setting = dict(model=model, dtype=torch.float32, extris=ext, intris=intr, fix_scale=fix_scale,
                   fixed_scale=None if not fix_scale else 1.1, pose_embedding=emb)
    data = dict(keypoints=[np.concatenate([fr["gt_uv"][v], fr["conf"][v][..., None]], -1) for v in range(V)])
    IG.init_guess(setting, data, use_torso=True, use_vposer=True)
    assert stub.calls[:2] == ["cams", "kp"] and stub.calls[2] == ("init", not fix_scale, 1.1 if fix_scale else 1.0, True, 0.0)
    assert float(emb.abs().max()) == 0.0                                    # init_guess.py:96-98
    assert float(model.betas.detach().abs().max()) == 0 and float(model.body_pose.detach().abs().max()) == 0     # zero-filled by reset_params
    assert np.abs(model.transl.detach().numpy() - fr["gt"]["transl"]).max() < 0.5
    if fix_scale:
        assert np.allclose(model.scale.detach().numpy(), 1.1)
    else:
        assert np.abs(model.scale.detach().numpy() - 1.0).max() < 0.25
    t0, r0, s0 = model.transl.detach().clone(), model.global_orient.detach().clone(), model.scale.detach().clone()
    IG.fix_params(setting, scale=1.1 if fix_scale else None, shape=[0.5] * 10 if fix_scale else None)
    bp = model.body_pose.detach().numpy()
    assert (bp[:, :6] == 1).all() and (bp[:, 6:] == 0).all()                 # init_guess.py:198-201
    assert torch.equal(model.transl.detach(), t0) and torch.equal(model.global_orient.detach(), r0)
    if fix_scale:
        assert not model.scale.requires_grad and not model.betas.requires_grad
        assert np.allclose(model.betas.detach().numpy(), 0.5) and np.allclose(model.scale.detach().numpy(), 1.1)
    else:
        assert model.scale.requires_grad and torch.equal(model.scale.detach(), s0)
    # one view (init_guess.py:54-78) and the as-written alignment reach the context as options of the same call
    setting1 = dict(setting, extris=setting["extris"][:1], intris=setting["intris"][:1])
    IG.init_guess(setting1, dict(keypoints=data["keypoints"][:1]), use_torso=True, umeyama_as_written=True)
    assert stub.as_written is True and stub.gt_uv.shape[0] == 1


def test_load_init_warm_start_and_fallback(syn_model, monkeypatch):
    model = make_model(syn_model, 1)
    setting = dict(model=model, dtype=torch.float32, device=None, seq_start=False)
    res = dict(loss=12.0, transl=np.full((1, 3), 0.4, np.float32), global_orient=np.full((1, 3), 0.1, np.float32),
               scale=np.array([[1.2]], np.float32), betas=np.full((1, 10), 0.3, np.float32),
               pose_embedding=np.full((1, 32), 0.7, np.float32))
    IG.load_init(setting, {}, res, use_vposer=True)
    assert np.allclose(model.transl.detach().numpy(), 0.4) and np.allclose(model.betas.detach().numpy(), 0.3)
    assert np.allclose(model.scale.detach().numpy(), 1.2) and float(model.body_pose.abs().max()) == 0
    assert setting["pose_embedding"].requires_grad and np.allclose(setting["pose_embedding"].detach().numpy(), 0.7)
    called = []
    monkeypatch.setattr(IG, "init_guess", lambda s, d, use_torso=False, **kw: called.append(use_torso))
    res["loss"] = 6000.0
    IG.load_init(setting, {}, res, use_torso=True)
    assert called == [True] and setting["seq_start"] is True                # init_guess.py:144-148
