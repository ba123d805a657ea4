"""TEST INFRASTRUCTURE -- writes tests/golden/fit_e2e_ref.npz: complete 4-stage fits of a few frames by the UNMODIFIED
reference (oracle/ref_fit.py -> utils/non_linear_solver.non_linear_solver, torch CPU fp32, SDF term off), with what the
benchmark metric counts: L-BFGS iterations and closure evaluations per stage, final loss, final parameters.

    python -m oracle.make_golden_fit            (authoring container: needs /root/reference)

L-BFGS with ftol = 1e-9 on an fp32 loss is chaotic in its last bits: two implementations whose closures
agree to 1e-7 stop after different iteration counts, so the fixture pins DISTRIBUTIONS (mean iterations / evaluations per
frame and stage, final losses), not single trajectories; tests/test_gpu_fit_e2e.py states the tolerances.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from oracle import ref_fit as RF  # noqa: E402

SEED, B, V = 4100, 16, 8


def main():
    torch.set_num_threads(1)
    model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(V)
    fr = S.make_frames(model, cams, B, seed=SEED)
    sc = RF.build_scene(model, gmm, cams)
    its, evs, fin, par = [], [], [], []
    for b in range(B):
        r = RF.fit_frame(sc, fr, b, S.STAGE_WEIGHTS, interpenetration=False)
        its.append([p[0] for p in r["per_stage"]]); evs.append([p[1] for p in r["per_stage"]])
        fin.append(r["final_loss"]); par.append(r["params"])
        print(b, r["per_stage"], r["final_loss"])
    out = os.path.join(ROOT, "tests", "golden", "fit_e2e_ref.npz")
    np.savez_compressed(out, seed=SEED, B=B, V=V, X0=S.pack_params(fr["init"]), iterations=np.array(its), evals=np.array(evs),
                        final_loss=np.array(fin, np.float64), params=np.stack(par))
    print("wrote", out, "mean iterations/frame", np.array(its).sum(1).mean(), "mean evals/frame", np.array(evs).sum(1).mean())


if __name__ == "__main__":
    main()
