"""Throughput of the 256 x 8 SDF workload with N batches in flight (one context + stream + host thread each).
Usage: python scripts/inflight_probe.py [max_inflight=4] [steps_per_lane=3]"""
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mvsmplfitting_b200 import synthetic as S  # noqa: E402
from mvsmplfitting_b200.context import FittingContext  # noqa: E402


def lane(i, B=256, V=8):
    model, gmm, cams = S.make_model(0), S.make_gmm(7), S.make_cameras(V)
    fr = S.make_frames(model, cams, B, seed=1000 + 17 * i)
    ctx = FittingContext(0)
    ctx.set_model(model)
    ctx.set_gmm_from_dict(gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"])
    ctx.set_batch(B)
    stages = [ctx.make_loss_config(body_prior="gmm", interpenetration=True, sdf_grid=128, **st) for st in bench.stage_table()]
    x0 = torch.tensor(S.pack_params(fr["init"]), device="cuda")
    ctx.set_keypoints(torch.tensor(fr["gt_uv"], device="cuda"), torch.tensor(fr["conf"], device="cuda"),
                      torch.tensor(fr["joint_weights"], device="cuda"))
    return dict(ctx=ctx, stages=stages, x0=x0, x=x0.clone(), opt=ctx.make_lbfgs_config(), stream=torch.cuda.Stream(), it=0)


def work(L, steps):
    torch.cuda.set_device(0)
    with torch.cuda.stream(L["stream"]):
        for _ in range(steps):
            L["x"].copy_(L["x0"])
            _, st = L["ctx"].fit(L["x"], L["stages"], L["opt"])
            L["it"] += st["frame_iterations"]
        L["stream"].synchronize()


def main():
    nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lanes = [lane(i) for i in range(nmax)]
    for L in lanes:
        work(L, 1)                                            # warm-up
    out = {}
    n = 1
    while n <= nmax:
        for L in lanes:
            L["it"] = 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        th = [threading.Thread(target=work, args=(lanes[i], steps)) for i in range(n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        it = sum(L["it"] for L in lanes[:n])
        out[n] = dict(ms_total=ms, wall_ms=(time.time() - t0) * 1e3, batches=n * steps, ms_per_batch=ms / (n * steps),
                      frame_iterations_per_s=it / (ms * 1e-3))
        print(n, out[n], flush=True)
        n *= 2
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_inflight_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
