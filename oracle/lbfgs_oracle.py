"""TEST INFRASTRUCTURE -- CPU restatement of the reference optimiser loop for ONE
frame: L-BFGS with strong-Wolfe cubic line search (code/optimizers/lbfgs_ls.py)
driven by FittingMonitor.run_fitting (code/utils/fitting.py:71-142).

The restatement works on a flat parameter vector and a callable
`fg(x) -> (loss: float, grad: tensor)` instead of nn.Parameters + closure, but it
keeps the reference's arithmetic types (fp32 0-dim tensors for dot products,
Python floats for losses) and every branch, so that iterates match the reference
when both are given the same closure.  Pinned by tests/test_oracle_vs_reference.py.

Counters follow BASELINE.md section 3: an *iteration* is one pass of the
`while n_iter < max_iter` body (lbfgs_ls.py:304-434), an *eval* is one closure call.
"""
from __future__ import annotations

import numpy as np
import torch


def cubic_min(x1, f1, g1, x2, f2, g2, bounds=None):
    """lbfgs_ls.py:11-36."""
    if bounds is not None:
        lo, hi = bounds
    else:
        lo, hi = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    sq = d1 ** 2 - g1 * g2
    if sq >= 0:
        d2 = sq.sqrt() if torch.is_tensor(sq) else sq ** 0.5
        if x1 <= x2:
            pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(pos, lo), hi)
    return (lo + hi) / 2.


def strong_wolfe(phi, t, d, f, g, gtd, c1=1e-4, c2=0.9, tol_change=1e-9, max_iter=20, max_ls=25):
    """lbfgs_ls.py:39-167.  `phi(t) -> (f_new: float, g_new: tensor)` evaluates the
    closure at x + t*d (lbfgs_ls.py:249-254)."""
    d_norm = d.abs().max()
    g = g.clone()
    f_new, g_new = phi(t)
    evals = 1
    gtd_new = g_new.dot(d)
    t_prev, f_prev, g_prev, gtd_prev = 0, f, g, gtd
    done = False
    it = 0
    br = None
    while it < max_ls:
        if f_new > (f + c1 * t * gtd) or (it > 1 and f_new >= f_prev):
            br = ([t_prev, t], [f_prev, f_new], [g_prev, g_new.clone()], [gtd_prev, gtd_new])
            break
        if abs(gtd_new) <= -c2 * gtd:
            br = ([t], [f_new], [g_new], [gtd_new])
            done = True
            break
        if gtd_new >= 0:
            br = ([t_prev, t], [f_prev, f_new], [g_prev, g_new.clone()], [gtd_prev, gtd_new])
            break
        lo_step = t + 0.01 * (t - t_prev)
        hi_step = t * 10
        keep = t
        t = cubic_min(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(lo_step, hi_step))
        t_prev, f_prev, g_prev, gtd_prev = keep, f_new, g_new.clone(), gtd_new
        f_new, g_new = phi(t)
        evals += 1
        gtd_new = g_new.dot(d)
        it += 1
    if it == max_ls:
        br = ([0, t], [f, f_new], [g, g_new], [gtd, gtd_new])
    bt, bf, bg, bgtd = br
    stalled = False
    lo, hi = (0, 1) if bf[0] <= bf[-1] else (1, 0)
    while not done and it < max_iter:
        t = cubic_min(bt[0], bf[0], bgtd[0], bt[1], bf[1], bgtd[1])
        eps = 0.1 * (max(bt) - min(bt))
        if min(max(bt) - t, t - min(bt)) < eps:
            if stalled or t >= max(bt) or t <= min(bt):
                if abs(t - max(bt)) < abs(t - min(bt)):
                    t = max(bt) - eps
                else:
                    t = min(bt) + eps
                stalled = False
            else:
                stalled = True
        else:
            stalled = False
        f_new, g_new = phi(t)
        evals += 1
        gtd_new = g_new.dot(d)
        it += 1
        if f_new > (f + c1 * t * gtd) or f_new >= bf[lo]:
            bt[hi], bf[hi], bg[hi], bgtd[hi] = t, f_new, g_new.clone(), gtd_new
            lo, hi = (0, 1) if bf[0] <= bf[1] else (1, 0)
        else:
            if abs(gtd_new) <= -c2 * gtd:
                done = True
            elif gtd_new * (bt[hi] - bt[lo]) >= 0:
                bt[hi], bf[hi], bg[hi], bgtd[hi] = bt[lo], bf[lo], bg[lo], bgtd[lo]
            bt[lo], bf[lo], bg[lo], bgtd[lo] = t, f_new, g_new.clone(), gtd_new
        if abs(bt[1] - bt[0]) * d_norm < tol_change:
            break
    return bf[lo], bg[lo], bt[lo], evals


class LBFGSOracle:
    """lbfgs_ls.py:172-445 with line_search_fn='strong_Wolfe' (optim_factory.py:50-52)."""

    def __init__(self, x0: torch.Tensor, fg, lr=1.0, max_iter=30, max_eval=None,
                 tolerance_grad=1e-5, tolerance_change=1e-9, history_size=100):
        self.x = x0.clone()
        self.fg = fg
        self.lr, self.max_iter = lr, max_iter
        self.max_eval = max_eval if max_eval is not None else max_iter * 5 // 4
        self.tol_g, self.tol_c, self.hist = tolerance_grad, tolerance_change, history_size
        self.st = dict(func_evals=0, n_iter=0)
        self.last_grad = None            # what p.grad holds after step() (the last closure call)
        self.iters = 0
        self.evals = 0

    def _eval(self, x):
        f, g = self.fg(x)
        self.evals += 1
        self.last_grad = g.clone()
        return float(f), g

    def step(self):
        st = self.st
        loss, flat_grad = self._eval(self.x)
        orig_loss = loss
        current_evals = 1
        st["func_evals"] += 1
        if flat_grad.abs().max() <= self.tol_g:
            return orig_loss
        d, t = st.get("d"), st.get("t")
        old_dirs, old_stps, ro = st.get("old_dirs"), st.get("old_stps"), st.get("ro")
        H_diag, prev_flat_grad, prev_loss = st.get("H_diag"), st.get("prev_flat_grad"), st.get("prev_loss")
        n_iter = 0
        while n_iter < self.max_iter:
            n_iter += 1
            st["n_iter"] += 1
            self.iters += 1
            if st["n_iter"] == 1:
                d = flat_grad.neg()
                old_dirs, old_stps, ro = [], [], []
                H_diag = 1
            else:
                y = flat_grad.sub(prev_flat_grad)
                s = d.mul(t)
                ys = y.dot(s)
                if ys > 1e-10:
                    if len(old_dirs) == self.hist:
                        old_dirs.pop(0); old_stps.pop(0); ro.pop(0)
                    old_dirs.append(y); old_stps.append(s); ro.append(1. / ys)
                    H_diag = (ys / y.dot(y)) * 1.
                k = len(old_dirs)
                al = [None] * k
                q = flat_grad.neg()
                for i in range(k - 1, -1, -1):
                    al[i] = old_stps[i].dot(q) * ro[i]
                    q.add_(old_dirs[i], alpha=-float(al[i]))
                d = r = torch.mul(q, H_diag)
                for i in range(k):
                    be = old_dirs[i].dot(r) * ro[i]
                    r.add_(old_stps[i], alpha=float(al[i] - be))
            prev_flat_grad = flat_grad.clone() if prev_flat_grad is None else prev_flat_grad.copy_(flat_grad)
            prev_loss = loss
            if st["n_iter"] == 1:
                t = min(1., 1. / flat_grad.abs().sum()) * self.lr
            else:
                t = self.lr
            gtd = flat_grad.dot(d)
            if gtd > -self.tol_c:
                break
            x0 = self.x.clone()

            def phi(tt, x0=x0, d=d):
                return self._eval(torch.add(x0, d, alpha=float(tt)))
            loss, flat_grad, t, ls_evals = strong_wolfe(phi, t, d, loss, flat_grad, gtd, max_iter=self.max_iter)
            self.x = torch.add(x0, d, alpha=float(t))
            opt_cond = flat_grad.abs().max() <= self.tol_g
            current_evals += ls_evals
            st["func_evals"] += ls_evals
            if n_iter == self.max_iter:
                break
            if current_evals >= self.max_eval:
                break
            if opt_cond:
                break
            if d.mul(t).abs().max() <= self.tol_c:
                break
            if abs(loss - prev_loss) < self.tol_c:
                break
        st.update(d=d, t=t, old_dirs=old_dirs, old_stps=old_stps, ro=ro, H_diag=H_diag,
                  prev_flat_grad=prev_flat_grad, prev_loss=prev_loss)
        return orig_loss


PARAM_SEGMENTS = ((0, 10), (10, 13), (13, 82), (82, 85), (85, 86))


def run_fitting(opt: LBFGSOracle, maxiters=30, ftol=1e-9, gtol=1e-9, segments=PARAM_SEGMENTS):
    """fitting.py:71-142 (visualisation branch removed).  Returns (final prev_loss, outer steps).
    Note the gtol test is abs(max(grad)) per parameter tensor on the grads left by
    the LAST closure call (fitting.py:115-117)."""
    prev_loss = None
    n_done = 0
    for n in range(maxiters):
        loss = opt.step()
        n_done = n + 1
        if np.isnan(loss) or np.isinf(loss):
            break
        if n > 0 and prev_loss is not None and ftol > 0:
            rel = (prev_loss - loss) / max([abs(prev_loss), abs(loss), 1])      # utils/utils.py:348-349
            if rel <= ftol:
                break
        g = opt.last_grad
        if all(abs(float(g[a:b].max())) < gtol for a, b in segments):
            break
        prev_loss = loss
    return prev_loss, n_done
