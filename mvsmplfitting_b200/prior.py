"""Pose / shape priors: drop-in for reference code/prior.py (create_prior :36-50, SMPLifyAnglePrior :53-89,
L2Prior :92-97, MaxMixturePrior :100-231).  Inside the closure the priors, their data-dependent guards
(fitting.py:334,349) and their gradients are evaluated by the CUDA frame kernel from the buffers these
modules hold; forward() is kept for standalone use."""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np
import torch
import torch.nn as nn

DEFAULT_DTYPE = torch.float32


def create_prior(prior_type, **kwargs):
    if prior_type == "gmm":
        return MaxMixturePrior(**kwargs)
    if prior_type == "l2":
        return L2Prior(**kwargs)
    if prior_type == "angle":
        return SMPLifyAnglePrior(**kwargs)
    if prior_type == "none" or prior_type is None:
        def no_prior(*args, **kwargs):
            return 0.0
        return no_prior
    raise ValueError("Prior {}".format(prior_type) + " is not implemented")


class SMPLifyAnglePrior(nn.Module):
    """exp(sign * angle)^2 on the elbow / knee bending components (prior.py:57-89)"""

    def __init__(self, dtype=torch.float32, **kwargs):
        super().__init__()
        self.register_buffer("angle_prior_idxs", torch.tensor([55, 58, 12, 15], dtype=torch.long))
        self.register_buffer("angle_prior_signs", torch.tensor([1, -1, -1, -1], dtype=dtype))

    def forward(self, pose, with_global_pose=False):
        idx = self.angle_prior_idxs - (not with_global_pose) * 3
        return torch.exp(pose[:, idx] * self.angle_prior_signs).pow(2)


class L2Prior(nn.Module):
    def __init__(self, dtype=DEFAULT_DTYPE, reduction="sum", **kwargs):
        super().__init__()

    def forward(self, module_input, *args):
        return torch.sum(module_input.pow(2))


class MaxMixturePrior(nn.Module):
    """GMM pose prior, "merged" min-over-components negative log-likelihood (prior.py:100-196).
    Loads `gmm_{N:02d}.pkl` (dict with means / covars / weights, or a legacy sklearn GMM)."""

    def __init__(self, prior_folder="prior", num_gaussians=6, dtype=DEFAULT_DTYPE, epsilon=1e-16, use_merged=True,
                 **kwargs):
        super().__init__()
        if dtype == DEFAULT_DTYPE:
            np_dtype = np.float32
        elif dtype == torch.float64:
            np_dtype = np.float64
        else:
            print("Unknown float type {}, exiting!".format(dtype))
            sys.exit(-1)
        self.num_gaussians = num_gaussians
        self.epsilon = epsilon
        self.use_merged = use_merged
        full = os.path.join(prior_folder, "gmm_{:02d}.pkl".format(num_gaussians))
        if not os.path.exists(full):
            print('The path to the mixture prior "{}"'.format(full) + " does not exist, exiting!")
            sys.exit(-1)
        with open(full, "rb") as f:
            gmm = pickle.load(f, encoding="latin1")
        if type(gmm) == dict:
            means, covs, weights = gmm["means"], gmm["covars"], gmm["weights"]
        elif "sklearn.mixture.gmm.GMM" in str(type(gmm)):
            means, covs, weights = gmm.means_, gmm.covars_, gmm.weights_
        else:
            print("Unknown type for the prior: {}, exiting!".format(type(gmm)))
            sys.exit(-1)
        covs64 = np.asarray(covs, dtype=np.float64)
        means = np.asarray(means).astype(np_dtype)
        covs = np.asarray(covs).astype(np_dtype)
        self.register_buffer("means", torch.tensor(means, dtype=dtype))
        self.register_buffer("covs", torch.tensor(covs, dtype=dtype))
        self.register_buffer("precisions", torch.tensor(np.stack([np.linalg.inv(c) for c in covs]).astype(np_dtype), dtype=dtype))
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in covs64])
        nll = np.asarray(np.asarray(weights, dtype=np.float64) / ((2 * np.pi) ** (69 / 2.0) * (sqrdets / sqrdets.min())))
        self.register_buffer("nll_weights", torch.tensor(nll, dtype=dtype).unsqueeze(0))
        self.register_buffer("weights", torch.tensor(np.asarray(weights), dtype=dtype).unsqueeze(0))
        self.random_var_dim = self.means.shape[1]

    def get_mean(self):
        return torch.matmul(self.weights, self.means)

    def forward(self, pose, betas=None):
        diff = pose.unsqueeze(1) - self.means
        quad = (torch.einsum("mij,bmj->bmi", self.precisions, diff) * diff).sum(-1)
        return torch.min(0.5 * quad - torch.log(self.nll_weights), dim=1)[0]
