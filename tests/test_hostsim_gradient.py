"""CPU: the hand-derived adjoint (mvs_math.cuh building blocks, executed serially on the host)
against the reference-run golden fixtures.  Catches derivation mistakes before any GPU time."""
import numpy as np
import pytest

from oracle.lbfgs_oracle import PARAM_SEGMENTS
from tests import golden_util as G
from tests.hostsim import HostSim


@pytest.mark.parametrize("name", G.closure_cases())
@pytest.mark.parametrize("use_double", [False, True])
def test_hostsim_matches_reference(name, use_double, syn_model, syn_gmm):
    c = G.load_case(name)
    hs = HostSim(syn_model, c["cams"], c["meta"]["model_type"])
    tag = "f64" if use_double else "f32"
    tol = 1e-9 if use_double else 1e-4
    B = c["X"].shape[0]
    for b in range(B):
        r = hs.eval(c["X"][b], c["gt_uv"][:, b], c["conf"][:, b], c["joint_weights"], c["w"],
                    body_prior=c["meta"]["body_prior"], gmm=syn_gmm, use_conf=c["meta"]["use_joints_conf"],
                    fix_shape=c["meta"]["fix_shape"], use_double=use_double, want_verts=True)
        assert abs(r["loss"] - c["loss_" + tag][b]) / abs(c["loss_" + tag][b]) < tol
        assert G.relmax(r["joints"], c["joints_" + tag][b]) < tol
        assert G.relmax(r["verts"][:64], c["verts_head_" + tag][b]) < tol
        g_ref = c["grad_" + tag][b]
        g = r["grad"].copy()
        if c["meta"]["fix_shape"]:
            g[:10] = 0
        for a, e in PARAM_SEGMENTS:
            if np.abs(g_ref[a:e]).max() > 0:
                assert G.relmax(g[a:e], g_ref[a:e]) < tol, (name, a, e)
