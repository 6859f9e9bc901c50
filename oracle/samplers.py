"""Oracle: MALA step, flow-MH steps, train_data_generator dispatch.  TEST INFRASTRUCTURE ONLY.

MALA follows /root/reference/bblackjax/mcmc/mala.py:51-54,57-120, diffusions.py:22-33,
util.py:57-82, proposal.py:80-122,152-159,169-186 *as coded* (the accept ratio is the
negative of the textbook log-ratio; do not "fix").
Flow steps follow exe_flow_matching.py:206-242 (push/pull + log-det), :246-260 (independent MH),
:264-278 (random-walk MH in latent space, the default), :300-314 (dispatch).
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

from . import ode, threefry as tf, vector_field as vf


class MALAState(NamedTuple):
    position: np.ndarray
    logdensity: np.ndarray
    logdensity_grad: np.ndarray


class MALAInfo(NamedTuple):
    acceptance_rate: np.ndarray
    is_accepted: np.ndarray
    proposed_position: np.ndarray
    proposed_weight: np.ndarray


def mala_init(x, target, beta=1.0):
    l, g = target.value_and_grad(x, beta)
    return MALAState(x, l, g)


def mala_step(keys, state, target, step_size, beta=1.0, noise=None, rng_dtype=None):
    """vmap(kernel)(keys, states).  keys: uint32 [N,2].  rng_dtype: dtype of the random draws
    (float32 = x64 off); arithmetic runs in state.position.dtype."""
    x, l, g = state
    dt = x.dtype
    N, d = x.shape
    h = dt.type(step_size)
    ki = np.empty((N, 2), np.uint32); kr = np.empty((N, 2), np.uint32)
    for n in range(N):
        ki[n], kr[n] = tf.split(keys[n])
    rdt = np.dtype(rng_dtype or dt)
    if noise is None:
        noise = tf.vmap_normal(ki, d, rdt).astype(dt)
    xn = x + h * g + np.sqrt(dt.type(2) * h) * noise
    ln, gn = target.value_and_grad(xn, beta)
    quarter = dt.type(0.25 * (1.0 / float(step_size)))   # python-float arithmetic, mala.py:79
    th_new = xn - x - h * g
    e_new = -l + quarter * (th_new * th_new).sum(1)          # transition_energy(state, new_state)
    th_prev = x - xn - h * gn
    e_prev = -ln + quarter * (th_prev * th_prev).sum(1)      # transition_energy(new_state, state)
    with np.errstate(invalid="ignore", over="ignore"):
        delta = e_prev - e_new
        delta = np.where(np.isnan(delta), -np.inf, delta).astype(dt)
        p_accept = np.minimum(np.exp(delta), dt.type(1))
        u = np.array([tf.uniform(kr[n], (), rdt) for n in range(N)], dt)
        acc = u < p_accept
        weight = np.exp(ln + quarter * (th_prev * th_prev).sum(1))
    new = MALAState(np.where(acc[:, None], xn, x), np.where(acc, ln, l), np.where(acc[:, None], gn, g))
    return new, MALAInfo(p_accept, acc, xn, weight), dict(delta=delta, u=u, noise=noise)


class Flow:
    """Bundles (params, omega, target, config) for the CNF push/pull."""

    def __init__(self, params, omega, target, hutch, rtol=1e-5, atol=1e-5, mxstep=1000,
                 grad_clip=None, ts=(0.0, 1.0), rng_dtype=None):
        self.rng_dtype = rng_dtype
        self.params, self.omega, self.target = params, omega, target
        self.hutch, self.rtol, self.atol, self.mxstep = hutch, rtol, atol, mxstep
        self.grad_clip, self.ts = grad_clip, ts

    def _probe(self, keys, d, dt):
        return tf.vmap_normal(keys, d, np.dtype(self.rng_dtype or dt)).astype(dt) if self.hutch else None

    def transform_and_logdet(self, keys, u, stats=None, z=None):
        """forward ODE of (v(u,t), -div) from t=0 to 1 (exe_flow_matching.py:206-221)."""
        N, d = u.shape
        if z is None:
            z = self._probe(keys, d, u.dtype)

        def aug(y, t):
            v, div = vf.field_and_div(self.params, self.omega, np.ascontiguousarray(y[:, :d]), t, self.target, z, self.grad_clip)
            return np.concatenate([v, -div[:, None]], 1)

        y0 = np.concatenate([u, np.zeros((N, 1), u.dtype)], 1)
        y1 = ode.odeint_final(aug, y0, self.ts, self.rtol, self.atol, self.mxstep, stats)
        return y1[:, :d], y1[:, d]

    def inverse_and_logdet(self, keys, x, stats=None, z=None):
        """ODE of (-v(x,1-s), +div v(x,1-s)) (exe_flow_matching.py:223-242)."""
        N, d = x.shape
        if z is None:
            z = self._probe(keys, d, x.dtype)

        def aug(y, s):
            t = x.dtype.type(1.0) - s
            v, div = vf.field_and_div(self.params, self.omega, np.ascontiguousarray(y[:, :d]), t, self.target, z, self.grad_clip)
            return np.concatenate([-v, div[:, None]], 1)

        y0 = np.concatenate([x, np.zeros((N, 1), x.dtype)], 1)
        y1 = ode.odeint_final(aug, y0, self.ts, self.rtol, self.atol, self.mxstep, stats)
        return y1[:, :d], y1[:, d]


def _split4(keys):
    N = keys.shape[0]
    out = np.empty((4, N, 2), np.uint32)
    for n in range(N):
        out[:, n] = tf.split(keys[n], 4)
    return out


def rw_flow_mh_step(keys, state, target, flow: Flow, beta=1.0, stats=None):
    """random_walk_metropolis_hastings (exe_flow_matching.py:262-278)."""
    x, l, g = state
    dt = x.dtype
    N, d = x.shape
    key_gen, key_acc, key_h1, key_h2 = _split4(keys)
    s_inv, s_fwd = {}, {}
    u0, V0 = flow.inverse_and_logdet(key_h2, x, s_inv)
    rdt = np.dtype(flow.rng_dtype or dt)
    scale = dt.type(2.38) / np.sqrt(dt.type(d))
    up = u0 + scale * tf.vmap_normal(key_gen, d, rdt).astype(dt)
    xp, Vp = flow.transform_and_logdet(key_h1, up, s_fwd)
    lp, gp = target.value_and_grad(xp, beta)
    with np.errstate(over="ignore", invalid="ignore"):
        log_acc = lp - Vp - l - V0
        acc_prob = np.exp(log_acc)
        u = np.array([tf.uniform(key_acc[n], (), rdt) for n in range(N)], dt)
        acc = u <= acc_prob
    new = MALAState(np.where(acc[:, None], xp, x), np.where(acc, lp, l), np.where(acc[:, None], gp, g))
    if stats is not None:
        stats.update(inv=s_inv, fwd=s_fwd, log_acc=log_acc, u=u, u0=u0, V0=V0, Vp=Vp, lp=lp)
    return new, MALAInfo(acc_prob, acc, xp, np.zeros(N, dt))


def indep_flow_mh_step(keys, state, target, flow: Flow, ref, beta=1.0, stats=None):
    """indep_metropolis_hastings (exe_flow_matching.py:246-260)."""
    x, l, g = state
    dt = x.dtype
    N, d = x.shape
    key_gen, key_acc, key_h1, key_h2 = _split4(keys)
    rdt = np.dtype(flow.rng_dtype or dt)
    up = ref.sample(key_gen, rdt).astype(dt)
    xp, Vp = flow.transform_and_logdet(key_h1, up)
    u0, V0 = flow.inverse_and_logdet(key_h2, x)
    lp, gp = target.value_and_grad(xp, beta)
    with np.errstate(over="ignore", invalid="ignore"):
        log_acc = lp - ref.loglik(up) - Vp + ref.loglik(u0) - V0 - l
        acc_prob = np.exp(log_acc)
        u = np.array([tf.uniform(key_acc[n], (), rdt) for n in range(N)], dt)
        acc = u <= acc_prob
    new = MALAState(np.where(acc[:, None], xp, x), np.where(acc, lp, l), np.where(acc[:, None], gp, g))
    if stats is not None:
        stats.update(log_acc=log_acc, u=u, lp=lp)
    return new, MALAInfo(acc_prob, acc, xp, np.zeros(N, dt))


def cis_flow_step(keys, state, target, flow: Flow, ref, n_is, beta=1.0, stats=None):
    """conditional_importance_sampling (exe_flow_matching.py:280-296), vmapped over chains.  Per chain: pull the current
    state back (its weight), push n_is fresh reference samples (each with its own probe key), draw ONE index from the
    n_is + 1 normalised weights with jax.random.choice(key_choice, n_is + 1, p=norm_weights) - a scalar draw: one uniform,
    r = p_cuml[-1] * (1 - u), searchsorted.  The gradient of an accepted sample is NOT recomputed (as coded: the new state
    keeps prev_state.logdensity_grad).  Not yet built on the device: this is the round-2 oracle."""
    from . import resample as R
    x, l, g = state
    dt = x.dtype
    N, d = x.shape
    rdt = np.dtype(flow.rng_dtype or dt)
    key_sample, key_hutch_prev, key_hutch, key_choice = _split4(keys)
    u_prev, vol_prev = flow.inverse_and_logdet(key_hutch_prev, x)
    with np.errstate(over="ignore", invalid="ignore"):
        prev_w = np.exp(l - ref.logprob(u_prev) - vol_prev)
    new_x, new_l = x.copy(), l.copy()
    acc = np.zeros(N, bool); rate = np.zeros(N, dt); prop = x.copy(); wsel = np.zeros(N, dt)
    for n in range(N):                                   # the reference vmaps this body over chains
        ks = tf.split(key_sample[n], n_is)
        refs = ref.sample(ks, rdt).astype(dt)
        kh = tf.split(key_hutch[n], n_is)
        samples, vols = flow.transform_and_logdet(kh, refs)
        ld = target.logprob(samples, beta)
        with np.errstate(over="ignore", invalid="ignore"):
            w = np.exp(ld - ref.logprob(refs) - vols)
        tot = prev_w[n] + w.sum()
        norm = np.concatenate([[prev_w[n]], w]) / tot
        idx, _ = R.choice_indices(key_choice[n], n_is + 1, 1, norm.astype(rdt), rdt)
        c = int(min(idx[0], n_is))
        rate[n] = wsel[n] = norm[c]
        if c > 0:
            acc[n] = True
            new_x[n], new_l[n], prop[n] = samples[c - 1], ld[c - 1], samples[c - 1]
    if stats is not None:
        stats.update(prev_weight=prev_w)
    return MALAState(new_x, new_l, g.copy()), MALAInfo(rate, acc, prop, wsel)


def train_data_generator(rng_key, states, count, target, flow, step_size, mcmc_per_flow_steps,
                         beta=1.0, num_importance_samples=0, ref=None):
    """exe_flow_matching.py:300-314 (integer mcmc_per_flow_steps >= 1 branch)."""
    N = states.position.shape[0]
    keys = tf.split(rng_key, N)
    if count % (int(mcmc_per_flow_steps) + 1) == 0:
        if num_importance_samples > 0:
            return cis_flow_step(keys, states, target, flow, ref, num_importance_samples, beta)
        if num_importance_samples < 0:
            return indep_flow_mh_step(keys, states, target, flow, ref, beta)
        return rw_flow_mh_step(keys, states, target, flow, beta)
    new, info, _ = mala_step(keys, states, target, step_size, beta)
    return new, info
