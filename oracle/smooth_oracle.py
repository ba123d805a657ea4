"""TEST INFRASTRUCTURE -- autograd restatement of the temporal smoothness energy of the jointly
regularised sequence mode.  This term does NOT exist in the reference: parity is
unpinned by construction; this file only pins the algebra the product uses (block-Jacobi anchors)."""
import torch


def smoothness_energy(x: torch.Tensor, lam: float, mask: torch.Tensor) -> torch.Tensor:
    """E_s = lam * sum_t ||(x_t - x_{t-1}) * mask||^2 over a whole sequence x [T,86]"""
    return lam * (((x[1:] - x[:-1]) * mask) ** 2).sum()


def smoothness_grad(x: torch.Tensor, lam: float, mask: torch.Tensor) -> torch.Tensor:
    xx = x.clone().double().requires_grad_(True)
    smoothness_energy(xx, lam, mask.double()).backward()
    return xx.grad
