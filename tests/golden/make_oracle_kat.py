"""Freeze known-answer vectors of the ORACLE (oracle/*.py) into tests/golden/oracle_kat.json.

The reference itself cannot run here (no JAX), so these are not reference outputs: they pin the oracle - the anchor of every
parity test - against accidental edits, and give a later session with JAX a ready-made list of values to confirm.
Every entry names the reference expression it restates.  Run from the repo root: python tests/golden/make_oracle_kat.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ode as OO, optim as OP, resample as OR, samplers as OS, targets as OT, threefry as tf, vector_field as VF  # noqa: E402

out = {}
k = tf.PRNGKey(59049)
out["jax.random.split(PRNGKey(59049), 6)"] = tf.split(k, 6).tolist()
out["jax.random.uniform(PRNGKey(59049), (5,), float32)"] = [float(v) for v in tf.uniform(k, (5,), np.float32)]
out["jax.random.normal(PRNGKey(59049), (5,), float32)"] = [float(v) for v in tf.normal(k, (5,), np.float32)]
out["jax.random.normal(PRNGKey(59049), (2,3), float32)"] = tf.normal(k, (2, 3), np.float32).astype(float).tolist()

x2 = np.array([[7.5, 8.25], [-0.5, 0.125]])
for name, t, x in [("four_mode", OT.four_mode(), x2), ("gmm16", OT.gmm16(), x2),
                   ("phi_four_d8", OT.PhiFour(8), np.linspace(-0.9, 0.8, 16).reshape(2, 8)),
                   ("indep_gauss_d3", OT.IndepGaussian(3), np.array([[0.5, -1.0, 2.0]]))]:
    out[f"{name}.logprob"] = t.logprob(x).tolist()
    out[f"{name}.grad"] = t.grad(x).tolist()
pines = OT.LogGaussianCoxPines(1600)
xp = pines.mu + 0.1 * np.sin(np.arange(1600.0))[None]
out["pines.logprob(mu + 0.1 sin(i))"] = pines.logprob(xp).tolist()
out["pines.loglik"] = pines.loglik(xp).tolist()
out["pines.grad[:4]"] = pines.grad(xp)[0, :4].tolist()
out["pines.logprob tempered 0.25"] = pines.logprob(xp, 0.25).tolist()

rng = np.random.default_rng(2024)
params = VF.init_params(rng, 2, 8, 4, head_scale=0.5, dtype=np.float64)
omega = rng.standard_normal(4)
t4 = OT.four_mode()
xx = np.array([[7.0, 9.0], [-8.5, 7.5], [0.25, -0.5]])
tt = np.array([0.1, 0.5, 0.9])
v, div = VF.field_and_div(params, omega, xx, tt, t4, None, None)
out["VectorFieldNet v"] = v.tolist(); out["exact divergence"] = div.tolist()
flow = OS.Flow(params, omega, t4, hutch=False)
y, ldj = flow.transform_and_logdet(None, xx)
out["transform_and_logdet x"] = y.tolist(); out["transform_and_logdet ldj"] = ldj.tolist()
times, xt, target = VF.fm_batch(tf.PRNGKey(7), xx, OT.IndepGaussian(2).sample, 1e-4, rng_dtype=np.float32)
loss, G = VF.fm_loss_and_grad(params, omega, xt, times, target, t4.grad, None)
out["flow_matching_loss"] = float(loss); out["dloss/dDense_7.bias"] = G["params"]["Dense_7"]["bias"].tolist()
st = OS.mala_init(xx, t4)
new, info, _ = OS.mala_step(tf.split(tf.PRNGKey(11), 3), st, t4, 0.2, rng_dtype=np.float32)
out["mala acceptance_rate"] = info.acceptance_rate.tolist(); out["mala is_accepted"] = [bool(b) for b in info.is_accepted]
out["mala proposed_position"] = info.proposed_position.tolist()
idx, _ = OR.choice_indices(tf.PRNGKey(3), 6, 12, np.array([0.1, 0.0, 0.3, 0.2, 0.0, 0.4], np.float32))
out["jax.random.choice(PRNGKey(3), 6, (12,), p=[.1,0,.3,.2,0,.4])"] = idx.tolist()
out["tempering beta_fn(0, loglik=40*sin(i), alpha=0.5)"] = OP.tempering_beta(0.0, 40.0 * np.sin(np.arange(256.0)), 0.5, dtype=np.float64)
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_kat.json"), "w"), indent=1)
print(len(out), "entries")
