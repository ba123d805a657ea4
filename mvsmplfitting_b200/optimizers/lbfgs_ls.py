"""L-BFGS with strong-Wolfe line search behind the reference's optimiser object
(code/optimizers/lbfgs_ls.py:172-445): `LBFGS(params, lr, max_iter, line_search_fn='strong_Wolfe')`,
`step(closure) -> loss`, `zero_grad()`.  The iteration itself runs on the GPU (mvs_lbfgs_step /
mvs_lbfgs_run in libmvsmpl): one independent problem per frame, state kept across step() calls."""
from __future__ import annotations

from torch.optim import Optimizer


class LBFGS(Optimizer):
    def __init__(self, params, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-5, tolerance_change=1e-9,
                 history_size=100, line_search_fn=None):
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size, line_search_fn=line_search_fn)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        if line_search_fn not in (None, "strong_Wolfe"):
            raise RuntimeError("only 'strong_Wolfe' is supported")
        if line_search_fn is None:
            raise RuntimeError("the B200 L-BFGS implements the strong_Wolfe variant the fitting path uses "
                               "(optim_factory.py:50-52)")
        self._params = self.param_groups[0]["params"]
        self._fresh = True
        self.stats = dict(frame_iterations=0, frame_evals=0, rounds=0)

    def lbfgs_config(self, ctx, max_outer=1, ftol=0.0, gtol=0.0):
        g = self.param_groups[0]
        return ctx.make_lbfgs_config(max_outer=max_outer, max_iter=g["max_iter"], max_eval=g["max_eval"],
                                     history_size=g["history_size"], lr=g["lr"], tolerance_grad=g["tolerance_grad"],
                                     tolerance_change=g["tolerance_change"], ftol=ftol, gtol=gtol)

    def step(self, closure):
        """One optimisation step for every frame; returns the (summed) loss at entry like the reference."""
        from ..fitting import FittingClosure
        if not isinstance(closure, FittingClosure):
            raise TypeError("mvsmplfitting_b200 LBFGS.step needs the closure made by "
                            "FittingMonitor.create_fitting_closure (the optimiser runs on the GPU)")
        if closure.use_vposer and not closure.vposer_native:
            # the iteration runs on the GPU, so the decoder has to as well (mvs_set_vposer, use_vposer = 2): that needs the
            # reference's decoder layers (model/VPoser.py:190-197) and a 32-d latent code.  No host-driven fallback.
            raise NotImplementedError("LBFGS.step with use_vposer=True needs a VPoser module with the reference's decoder "
                                      "layers (bodyprior_dec_fc1 / fc2 / out) and a [B,32] pose_embedding")
        x = closure.gather_params()
        closure.sync_loss_config()
        loss, grad, st = closure.ctx.lbfgs_step(x, self.lbfgs_config(closure.ctx), reset=self._fresh)
        self._fresh = False
        for k in self.stats:
            self.stats[k] += st[k]
        closure.scatter_params(x, grad)
        return loss.sum()
