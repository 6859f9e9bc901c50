"""Oracle: optimizer + LR schedule of the FM update.  TEST INFRASTRUCTURE ONLY.

Restates optax 0.1.9 (third-party, pinned by /root/reference/environment.yaml:177) as configured
in /root/reference/exe_flow_matching.py:129-137,184 and :189-198:
    tx = apply_if_finite(chain(adamw(lr_fn, b1, b2, eps, weight_decay, mask=no-bias), clip(1.0)), 10)
adamw = scale_by_adam -> add_decayed_weights(mask) -> scale_by_schedule(-lr); ``clip`` acts on the
UPDATES, elementwise to [-1, 1].  apply_if_finite inspects the incoming gradients: when any is
non-finite the update is zeroed and the inner state (moments, schedule count) is not advanced,
unless more than ``max_consecutive_errors`` consecutive failures occurred.
**parity unpinned** (no optax here).
"""
from __future__ import annotations

import numpy as np


def learning_rate_fn(num_train_steps, num_warmup_steps, learning_rate):
    """create_learning_rate_fn (exe_flow_matching.py:189-198): join_schedules([warmup, decay],
    [num_warmup_steps]); optax.linear_schedule with transition_steps<=0 is the constant
    init_value, so warmup_steps=0 yields pure linear decay lr*(1-step/num_train_steps)."""
    def lin(init, end, steps):
        if steps <= 0:
            return lambda c: init
        return lambda c: (init - end) * (1 - np.clip(c, 0, steps) / steps) + end
    warm = lin(0.0, learning_rate, num_warmup_steps)
    decay = lin(learning_rate, 0.0, num_train_steps - num_warmup_steps)
    return lambda step: warm(step) if step < num_warmup_steps else decay(step - num_warmup_steps)


class AdamWClipIfFinite:
    def __init__(self, params, lr_fn, b1=0.9, b2=0.999, eps=1e-8, weight_decay=1e-4, clip=1.0,
                 max_consecutive_errors=10):
        self.lr_fn, self.b1, self.b2, self.eps, self.wd, self.clip = lr_fn, b1, b2, eps, weight_decay, clip
        self.max_err = max_consecutive_errors
        self.mu = {k: {n: np.zeros_like(a) for n, a in v.items()} for k, v in params["params"].items()}
        self.nu = {k: {n: np.zeros_like(a) for n, a in v.items()} for k, v in params["params"].items()}
        self.count = 0            # scale_by_adam.count == scale_by_schedule.count
        self.notfinite_count = 0
        self.total_notfinite = 0
        self.last_finite = True

    def update(self, grads, params):
        """Returns new params (optax.apply_updates(params, updates))."""
        G = grads["params"]; P = params["params"]
        finite = all(np.isfinite(a).all() for v in G.values() for a in v.values())
        self.notfinite_count = 0 if finite else self.notfinite_count + 1
        self.last_finite = finite
        if not finite:
            self.total_notfinite += 1
        if not (finite or self.notfinite_count > self.max_err):
            return params
        f32 = np.float32
        c = self.count + 1
        bc1 = f32(1) - f32(self.b1) ** f32(c)
        bc2 = f32(1) - f32(self.b2) ** f32(c)
        lr = f32(self.lr_fn(self.count))
        new = {}
        for k in P:
            new[k] = {}
            for n in P[k]:
                g = G[k][n].astype(f32)
                self.mu[k][n] = (f32(1 - self.b1) * g + f32(self.b1) * self.mu[k][n]).astype(f32)
                self.nu[k][n] = (f32(1 - self.b2) * g * g + f32(self.b2) * self.nu[k][n]).astype(f32)
                u = (self.mu[k][n] / bc1) / (np.sqrt(self.nu[k][n] / bc2) + f32(self.eps))
                if n != "bias":
                    u = u + f32(self.wd) * P[k][n]
                u = np.clip(-lr * u, -f32(self.clip), f32(self.clip))
                new[k][n] = (P[k][n] + u).astype(f32)
        self.count = c
        return {"params": new}


def tempering_beta(prev_beta, logliks, alpha, maxiter=30, tol=1e-5, dtype=np.float32):
    """beta_fn (exe_flow_matching.py:391-402) with jaxopt 0.8.3 Bisection restated (parity unpinned):
    bracket sign from f(lower), f(upper) (0 when not bracketing, check_bracket=False); each iteration
    evaluates the midpoint, moves the bracket, stops when |f(mid)| <= tol; returns the last midpoint."""
    dt = np.dtype(dtype).type
    ll = np.asarray(logliks, dtype)
    n = ll.shape[0]

    def ess_zero(beta):
        logw = ll * dt(beta - prev_beta)
        w = np.exp(logw - logw.max())
        w = w / w.sum()
        return dt(1.0) / (w * w).sum() - dt(alpha * n)

    lo, hi = dt(prev_beta), dt(1.0)
    flo, fhi = ess_zero(lo), ess_zero(hi)
    sign = 1.0 if (flo < 0 and fhi >= 0) else (-1.0 if (flo > 0 and fhi <= 0) else 0.0)
    mid, err, it = lo, np.inf, 0
    while it < maxiter and err > tol:
        mid = dt(0.5) * (hi + lo)
        v = ess_zero(mid)
        too_large = sign * v > 0
        hi = mid if too_large else hi
        lo = lo if too_large else mid
        err = abs(v)
        it += 1
    return float(mid)
