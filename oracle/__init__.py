"""CPU oracle for the Markovian-Flow-Matching hot path (TEST INFRASTRUCTURE ONLY).

This package is a NumPy restatement of the reference algorithm (albcab/mfm) and of the
third-party numerics it calls (jax.random threefry2x32, jax.experimental.ode.odeint,
flax Dense, optax adamw/clip/apply_if_finite).  It exists to CHECK the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``mfm_b200/`` imports it; the product
path fails loudly when the CUDA library is missing.

PARITY STATUS: **parity unpinned** except for the RNG.  JAX/flax/optax are not installable in
this image (no wheels, no network) and the reference ships no tests or golden vectors, so the
restatement cannot be executed against the reference itself.  Pins that do exist:
  * threefry2x32: Random123 known-answer vectors (tests/test_oracle_rng.py);
  * jax.random layout: ``split(PRNGKey(0))``, ``uniform(PRNGKey(0))`` and the
    ``random.normal`` values printed in JAX's public "Sharp Bits" notebook;
  * pines constants: bin-count sums / Cholesky log-det cross-checked against SURVEY.md §8(c);
  * analytic gradients/Hessians checked by finite differences; ODE checked on linear fields.
"""
