"""Freeze the 16-mode GMM constants (multi_modal.py:39-45) as a fixture.

modes/covs follow the threefry oracle in float32 mode (split(PRNGKey(0),3); uniform(-12.8,12.8);
exp(0.5*normal)).  jax.random.dirichlet (gamma rejection sampler) is NOT restated: weights come
from numpy.random.default_rng(0).dirichlet(4*ones(16)) and are therefore NOT the reference's
weights (stated in DESIGN.md).  Columns: mode_x mode_y cov_x cov_y weight.
"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import threefry as tf

km, kc, kw = tf.split(tf.PRNGKey(0), 3)
modes = tf.uniform(km, (16, 2), np.float32, 16 * -0.8, 16 * 0.8)
covs = np.exp(np.float32(0.5) * tf.normal(kc, (16, 2), np.float32))
w = np.random.default_rng(0).dirichlet(4.0 * np.ones(16)).astype(np.float32)
out = np.concatenate([modes, covs, w[:, None]], 1).astype(np.float64)
np.savetxt("mfm_b200/data/gmm16.txt", out, fmt="%.9g")
print(out)
