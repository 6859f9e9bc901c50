"""Oracle: final sampling + importance resampling.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/exe_flow_matching.py:389 (sample_reference) and :453-459 (flow samples, importance weights,
multinomial resampling).  jax.random.choice (jax 0.4.26 jax/_src/random.py, third-party, pinned by environment.yaml:101)
is restated for the call the reference makes - replace=True, p given, axis 0:
    p_cuml = cumsum(p);  r = p_cuml[-1] * (1 - uniform(key, shape, p.dtype));  ind = searchsorted(p_cuml, r)
**parity unpinned**: no JAX here; XLA's summation order inside cumsum is not part of any contract, NumPy's sequential
float32 order is used (and reproduced bit for bit by the CUDA kernel).
"""
from __future__ import annotations

import numpy as np

from . import threefry as tf


def choice_indices(key, n_pop, n_draw, p, dtype=np.float32):
    dt = np.dtype(dtype)
    p = np.asarray(p, dt)
    assert p.shape == (n_pop,)
    p_cuml = np.cumsum(p, dtype=dt)
    u = tf.uniform(key, (n_draw,), dt)
    r = p_cuml[-1] * (dt.type(1) - u)
    return np.searchsorted(p_cuml, r, side="left").astype(np.int64), dict(p_cuml=p_cuml, r=r)


def choice(key, a, n_draw, p, dtype=np.float32):
    idx, _ = choice_indices(key, a.shape[0], n_draw, p, dtype)
    return np.take(a, np.minimum(idx, a.shape[0] - 1), axis=0), idx


def sample_flow(key_gen, target, ref, flow, n_samples, rng_dtype=np.float32, dtype=np.float64):
    """exe_flow_matching.py:453-459.  `flow` is oracle.samplers.Flow; `ref` an oracle IndepGaussian.
    key_gen is used twice, as coded: for the reference samples and for split(key_gen) -> (key_hutch, key_choice);
    the SAME key_hutch drives every row's Hutchinson probe (the vmap closes over it)."""
    dt = np.dtype(dtype)
    u = ref.sample(tf.split(key_gen, n_samples), rng_dtype).astype(dt)
    key_hutch, key_choice = tf.split(key_gen)
    keys = np.broadcast_to(key_hutch, (n_samples, 2)).copy()
    flow_samples, vols = flow.transform_and_logdet(keys, u)
    logd = target.logprob(flow_samples)
    log_w = logd - ref.logprob(u) - vols
    w = np.exp(log_w - log_w.max())
    exact, idx = choice(key_choice, flow_samples, n_samples, w.astype(rng_dtype), rng_dtype)
    return dict(u=u, flow_samples=flow_samples, vols=vols, samples_logdensity=logd, log_weights=log_w, weights=w,
                exact_samples=exact, indices=idx)
