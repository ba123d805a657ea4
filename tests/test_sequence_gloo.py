"""CPU, world_size 2 and 3 over gloo: the host logic of the multi-GPU sequence mode -- frame sharding, the
344-byte halo exchange, the block-Jacobi anchors and the single scalar all-reduce of the smoothness energy."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvsmplfitting_b200 import sequence as Q


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_the_sequence():
    for T in (1, 7, 256, 1024, 1025):
        for W in (1, 2, 3, 8):
            spans = [Q.shard_bounds(T, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == T
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, T, lam, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    X = torch.randn(T, 86)
    a, b = Q.shard_bounds(T, world, rank)
    x = X[a:b].clone()
    mask = Q.smooth_mask()
    left, right = Q.exchange_halo(x)
    anchor, wf = Q.neighbour_anchor(x, left, right, lam, mask)
    e = Q.smoothness_energy(x, left, lam, mask)
    torch.save(dict(a=a, b=b, left=left, right=right, anchor=anchor, wf=wf, e=float(e)), os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 11), (3, 10), (2, 2)])
def test_halo_anchor_and_energy_match_single_process(world, T, tmp_path):
    from oracle import smooth_oracle as SO
    lam = 0.7
    mp.spawn(_worker, args=(world, _free_port(), T, lam, str(tmp_path)), nprocs=world, join=True)
    torch.manual_seed(0)
    X = torch.randn(T, 86)
    mask = Q.smooth_mask()
    full_anchor, full_wf = Q.neighbour_anchor(X, None, None, lam, mask)
    e_ref = float(SO.smoothness_energy(X.double(), lam, mask.double()))
    g_ref = SO.smoothness_grad(X, lam, mask)
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, "r%d.pt" % r), weights_only=False)
        a, b = d["a"], d["b"]
        if a > 0:
            assert torch.equal(d["left"], X[a - 1])
        else:
            assert d["left"] is None
        if b < T:
            assert torch.equal(d["right"], X[b])
        else:
            assert d["right"] is None
        assert torch.allclose(d["anchor"], full_anchor[a:b]) and torch.allclose(d["wf"], full_wf[a:b])
        assert abs(d["e"] - e_ref) / max(e_ref, 1e-12) < 1e-6          # every rank holds the all-reduced scalar
        # block-Jacobi identity: d/dx_t of  wf_t * ||(x_t - anchor_t) mask||^2  ==  d E_s / d x_t
        g = 2.0 * d["wf"][:, None] * mask[None] * (X[a:b] - d["anchor"])
        assert torch.allclose(g.double(), g_ref[a:b], atol=1e-5)
