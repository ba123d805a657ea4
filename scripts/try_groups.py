"""experiment: split the 256 frames over G contexts / streams / host threads on ONE GPU"""
import sys, os, threading, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mvsmplfitting_b200 import synthetic as S
from mvsmplfitting_b200.context import FittingContext

model = S.make_model(0); gmm = S.make_gmm(7); cams = S.make_cameras(8)
B = 256
fr = S.make_frames(model, cams, B, seed=1000)
X0 = S.pack_params(fr["init"])
for G in (1, 2, 4):
    per = (B + G - 1) // G
    groups = []
    for g in range(G):
        sl = slice(g * per, (g + 1) * per)
        ctx = FittingContext(0)
        ctx.set_model(model); ctx.set_gmm_from_dict(gmm)
        ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(per)
        stream = torch.cuda.Stream()
        with torch.cuda.stream(stream):
            ctx.set_keypoints(torch.tensor(fr["gt_uv"][:, sl]).cuda(), torch.tensor(fr["conf"][:, sl]).cuda(), torch.tensor(fr["joint_weights"]).cuda())
        stages = [ctx.make_loss_config(body_prior="gmm", interpenetration=True, sdf_grid=128, **st) for st in bench.stage_table()]
        groups.append(dict(ctx=ctx, stream=stream, x0=torch.tensor(X0[sl]).cuda(), x=torch.tensor(X0[sl]).cuda(), stages=stages, it=0))
    torch.cuda.synchronize()
    def work(gr):
        with torch.cuda.stream(gr["stream"]):
            gr["x"].copy_(gr["x0"])
            _, st = gr["ctx"].fit(gr["x"], gr["stages"])
            gr["it"] = st["frame_iterations"]
    def step():
        ths = [threading.Thread(target=work, args=(gr,)) for gr in groups]
        for t in ths: t.start()
        for t in ths: t.join()
    step(); step()
    torch.cuda.synchronize()
    t0 = time.time()
    n = 3
    for _ in range(n): step()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / n
    it = sum(gr["it"] for gr in groups)
    print(json.dumps(dict(G=G, ms_per_step=dt * 1e3, frame_it_per_s=it / dt)))
    for gr in groups: gr["ctx"].close()
