"""Oracle: jax.experimental.ode.odeint (adaptive Dormand-Prince 5(4)), batched over chains.
TEST INFRASTRUCTURE ONLY.

Restates jax 0.4.26 jax/experimental/ode.py (third-party, NOT under /root/reference; pinned by
environment.yaml:101) as it is called from /root/reference/exe_flow_matching.py:345-349:
``odeint(func, x0, linspace(0,1,2 or 5), rtol, atol, mxstep)``; the augmented state is
ravel((x, ldj)) of length d+1 (:208-221, :225-242).  Under ``jax.vmap`` every chain runs its own
controller inside one batched while_loop; lanes whose condition is false keep their state.
That is what the per-chain masks below reproduce.  **parity unpinned** (no JAX here); checked on
linear fields with known solution/log-det in tests/test_oracle_ode.py.
"""
from __future__ import annotations

import numpy as np

ALPHA = np.array([1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0])
BETA = np.array([
    [1 / 5, 0, 0, 0, 0, 0],
    [3 / 40, 9 / 40, 0, 0, 0, 0],
    [44 / 45, -56 / 15, 32 / 9, 0, 0, 0],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0, 0],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656, 0],
    [35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
])
C_SOL = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0])
C_ERR = np.array([35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
                  -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0])
C_MID = np.array([6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2,
                  -2691868925 / 45128329728 / 2, 187940372067 / 1594534317056 / 2,
                  -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2])


def _norm(a):
    return np.sqrt((a * a).sum(1))


def initial_step_size(fun, t0, y0, order, rtol, atol, f0):
    dt = y0.dtype
    scale = dt.type(atol) + np.abs(y0) * dt.type(rtol)
    d0 = _norm(y0 / scale)
    d1 = _norm(f0 / scale)
    with np.errstate(divide="ignore", invalid="ignore"):
        h0 = np.where((d0 < 1e-5) | (d1 < 1e-5), dt.type(1e-6), dt.type(0.01) * d0 / d1).astype(dt)
    y1 = y0 + h0[:, None] * f0
    f1 = fun(y1, t0 + h0)
    d2 = _norm((f1 - f0) / scale) / h0
    with np.errstate(divide="ignore", invalid="ignore"):
        h1 = np.where((d1 <= 1e-15) & (d2 <= 1e-15),
                      np.maximum(dt.type(1e-6), h0 * dt.type(1e-3)),
                      (dt.type(0.01) / np.maximum(d1, d2)) ** dt.type(1.0 / (order + 1.0))).astype(dt)
    return np.minimum(dt.type(100.0) * h0, h1)


def odeint_final(fun, y0, ts, rtol, atol, mxstep, stats=None):
    """Returns y(ts[-1]) for a batch y0 [N, D].  ``fun(y [N,D], t [N]) -> [N,D]``.
    The step sequence does not depend on intermediate output times (steps are never clamped to
    them) apart from the per-segment ``mxstep`` budget, which is reproduced."""
    dt_ = y0.dtype
    N = y0.shape[0]
    c = lambda a: a.astype(dt_)
    t = np.full(N, ts[0], dt_)
    f = fun(y0, t)
    h = np.clip(initial_step_size(fun, t, y0, 4, rtol, atol, f), 0, np.inf).astype(dt_)
    y = y0.copy()
    last_t = t.copy()
    coeff = np.stack([y0] * 5)          # a,b,c,d,e
    n_eval = 2
    n_acc = np.zeros(N, np.int64)
    n_try = np.zeros(N, np.int64)
    y_target = y0
    for target in ts[1:]:
        target = dt_.type(target)
        i = np.zeros(N, np.int64)
        while True:
            active = (t < target) & (i < mxstep) & (h > 0)
            if not active.any():
                break
            k = np.zeros((7,) + y.shape, dt_)
            k[0] = f
            for s in range(1, 7):
                ti = t + h * dt_.type(ALPHA[s - 1])
                yi = y + h[:, None] * np.tensordot(c(BETA[s - 1, :s]), k[:s], axes=(0, 0))
                k[s] = fun(yi, ti)
            n_eval += 6
            y1 = h[:, None] * np.tensordot(c(C_SOL), k, axes=(0, 0)) + y
            err = h[:, None] * np.tensordot(c(C_ERR), k, axes=(0, 0))
            tol = dt_.type(atol) + dt_.type(rtol) * np.maximum(np.abs(y), np.abs(y1))
            ratio = np.sqrt(np.mean((err / tol) ** 2, axis=1)).astype(dt_)
            y_mid = y + h[:, None] * np.tensordot(c(C_MID), k, axes=(0, 0))
            hh = h[:, None]
            dy0, dy1 = k[0], k[6]
            ca = -2. * hh * dy0 + 2. * hh * dy1 - 8. * y - 8. * y1 + 16. * y_mid
            cb = 5. * hh * dy0 - 3. * hh * dy1 + 18. * y + 14. * y1 - 32. * y_mid
            cc = -4. * hh * dy0 + hh * dy1 - 11. * y - 5. * y1 + 16. * y_mid
            cd = hh * dy0
            new_coeff = np.stack([ca, cb, cc, cd, y]).astype(dt_)
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                dfac = np.where(ratio < 1, dt_.type(1.0), dt_.type(0.2))
                fac = np.minimum(dt_.type(10.0),
                                 np.maximum(ratio ** dt_.type(-1.0 / 5.0) * dt_.type(0.9), dfac))
                new_h = np.where(ratio == 0, h * dt_.type(10.0), h * fac).astype(dt_)
            new_h = np.clip(new_h, 0, np.inf)
            acc = active & (ratio <= 1.0)
            # jnp.where(error_ratio <= 1, new, old); i and dt advance either way (when active)
            coeff = np.where(acc[None, :, None], new_coeff, coeff)
            last_t = np.where(acc, t, last_t)
            t = np.where(acc, t + h, t)
            y = np.where(acc[:, None], y1, y)
            f = np.where(acc[:, None], k[6], f)
            h = np.where(active, new_h, h)
            i = i + active
            n_acc += acc
            n_try += active
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = ((target - last_t) / (t - last_t)).astype(dt_)[:, None]
        y_target = coeff[0]
        for j in range(1, 5):
            y_target = y_target * rel + coeff[j]
    if stats is not None:
        stats.update(n_eval=n_eval, n_acc=n_acc, n_try=n_try)
    return y_target
