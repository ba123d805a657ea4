"""create_optimizer with the reference's signature (code/optimizers/optim_factory.py:27-65)."""
from __future__ import annotations

import torch.optim as optim

from .lbfgs_ls import LBFGS as LBFGSLs


def create_optimizer(parameters, optim_type="lbfgs", lr=1e-3, momentum=0.9, use_nesterov=True, beta1=0.9, beta2=0.999,
                     epsilon=1e-8, use_locking=False, weight_decay=0.0, centered=False, rmsprop_alpha=0.99,
                     maxiters=20, gtol=1e-6, ftol=1e-9, **kwargs):
    if optim_type == "adam":
        return optim.Adam(parameters, lr=lr, betas=(beta1, beta2), weight_decay=weight_decay), False
    if optim_type == "lbfgs":
        return optim.LBFGS(parameters, lr=lr, max_iter=maxiters), False
    if optim_type == "lbfgsls":
        return LBFGSLs(parameters, lr=lr, max_iter=maxiters, line_search_fn="strong_Wolfe"), False
    if optim_type == "rmsprop":
        return optim.RMSprop(parameters, lr=lr, eps=epsilon, alpha=rmsprop_alpha, weight_decay=weight_decay,
                             momentum=momentum, centered=centered), False
    if optim_type == "sgd":
        return optim.SGD(parameters, lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=use_nesterov), False
    raise ValueError("Optimizer {} not supported!".format(optim_type))
