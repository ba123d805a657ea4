"""Times the batched closure (CUDA events) for a few (B, V, dense) points; prints JSON lines."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvsmplfitting_b200 import synthetic as S
from mvsmplfitting_b200.context import FittingContext

model = S.make_model(0)
gmm = S.make_gmm(7)
p = torch.cuda.get_device_properties(0)
print(json.dumps(dict(gpu=p.name, sms=p.multi_processor_count, l2=p.L2_cache_size, smem_optin=p.shared_memory_per_block_optin,
                      mem_gb=p.total_memory / 2**30, ref_present=os.path.isdir("/root/reference"))))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for B, V in [(1, 8), (64, 4), (256, 8), (1024, 8)]:
    cams = S.make_cameras(V)
    fr = S.make_frames(model, cams, B, seed=1)
    ctx = FittingContext(0)
    ctx.set_model(model); ctx.set_gmm_from_dict(gmm)
    ctx.set_cameras(cams["R"], cams["t"], cams["f"], cams["c"]); ctx.set_batch(B)
    ctx.set_keypoints(fr["gt_uv"], fr["conf"], fr["joint_weights"])
    ctx.set_loss(body_prior="gmm", data_weight=500 / 1536, body_pose_weight=4.78, shape_weight=5.0, bending_prior_weight=15.15)
    x = torch.tensor(S.pack_params(fr["init"]), device="cuda")
    for dense in (False, True):
        for _ in range(3):
            ctx.closure(x, want_verts=dense)
        torch.cuda.synchronize()
        ts = []
        for it in range(10):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ctx.closure(x, want_verts=dense); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        # back-to-back (warm L2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.closure(x, want_verts=dense)
        e1.record(); torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 20
        print(json.dumps(dict(B=B, V=V, dense=dense, ms_cold_median=ts[len(ts) // 2], ms_warm=warm,
                              frame_closures_per_s=B / (warm * 1e-3))))
    ctx.close()
