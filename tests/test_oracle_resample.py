"""jax.random.choice restatement (oracle/resample.py): definition, edge cases, statistics; and the C-ABI's host-visible parts."""
import numpy as np

from oracle import resample as OR, threefry as tf


def test_choice_is_inverse_cdf_of_one_uniform_draw():
    key = tf.PRNGKey(4)
    p = np.array([0.1, 0.0, 0.4, 0.5, 0.0], np.float32)
    idx, aux = OR.choice_indices(key, 5, 1000, p)
    u = tf.uniform(key, (1000,), np.float32)
    cum = np.array([0.1, 0.1, 0.5, 1.0, 1.0], np.float32)
    for i in range(1000):
        r = cum[-1] * (np.float32(1) - u[i])
        j = 0
        while cum[j] < r:
            j += 1
        assert idx[i] == j
    assert not np.isin(idx, [1, 4]).any()               # zero-probability entries are never drawn
    assert abs((idx == 3).mean() - 0.5) < 0.05 and abs((idx == 0).mean() - 0.1) < 0.03


def test_choice_accepts_unnormalised_weights_and_takes_rows():
    key = tf.PRNGKey(11)
    a = np.arange(12, dtype=np.float64).reshape(4, 3)
    w = np.array([2.0, 6.0, 0.0, 2.0], np.float32)       # the reference passes exp(lw - max), not probabilities
    rows, idx = OR.choice(key, a, 4000, w)
    assert rows.shape == (4000, 3) and np.array_equal(rows, a[idx])
    assert abs((idx == 1).mean() - 0.6) < 0.03 and (idx != 2).all()
    idx2, _ = OR.choice_indices(key, 4, 4000, w / w.sum())
    assert (idx == idx2).mean() > 0.999                   # same draws; scaling only moves boundaries by rounding


def test_choice_single_population_and_degenerate_weights():
    idx, _ = OR.choice_indices(tf.PRNGKey(0), 1, 7, np.array([3.0], np.float32))
    assert (idx == 0).all()
    idx, _ = OR.choice_indices(tf.PRNGKey(0), 3, 50, np.array([0.0, 0.0, 1.0], np.float32))
    assert (idx == 2).all()


def test_sample_flow_identity_flow_reproduces_reference_samples():
    """Zero-init heads (identity flow): flow samples = reference samples, vols = 0, weights = pi/q ratio."""
    from oracle import samplers as OS, targets as OT, vector_field as VF
    rng = np.random.default_rng(0)
    params = VF.init_params(rng, 2, 8, 4, head_scale=0.0, dtype=np.float64)
    for i in (4, 7):
        params["params"][f"Dense_{i}"]["bias"][:] = 0
    t, ref = OT.four_mode(), OT.IndepGaussian(2)
    flow = OS.Flow(params, rng.standard_normal(4), t, hutch=True, rng_dtype=np.float32)
    key = tf.PRNGKey(2)
    out = OR.sample_flow(key, t, ref, flow, 64)
    assert np.array_equal(out["flow_samples"], out["u"]) and (out["vols"] == 0).all()
    for i, k in enumerate(tf.split(key, 64)):            # :389: row i = sample_model(split(key_gen, n)[i])
        assert np.array_equal(out["u"][i], tf.normal(k, (2,), np.float32).astype(np.float64))
    lw = t.logprob(out["u"]) - ref.logprob(out["u"])
    assert np.allclose(out["log_weights"], lw) and out["weights"].max() == 1.0
    k_h, k_c = tf.split(key)
    idx, _ = OR.choice_indices(k_c, 64, 64, out["weights"].astype(np.float32))
    assert np.array_equal(out["indices"], idx) and np.array_equal(out["exact_samples"], out["flow_samples"][idx])
