"""Oracle vector field / FM loss / MALA / flow-MH (oracle/vector_field.py, oracle/samplers.py, oracle/optim.py) checked
against finite differences, closed forms written straight from the reference's formulas, and structural identities
(zero-init heads = identity flow, push o pull = identity)."""
import numpy as np
import pytest

from oracle import optim as OO, samplers as OS, targets as OT, threefry as tf, vector_field as VF


def _small(seed=0, d=3, H=8, F=4, head=0.3):
    rng = np.random.default_rng(seed)
    params = VF.init_params(rng, d, H, F, head_scale=head, dtype=np.float64)
    for v in params["params"].values():          # move the pre-activations away from the relu kink
        v["bias"] += 0.05 * rng.standard_normal(v["bias"].shape)
    omega = rng.standard_normal(F)
    return params, omega, rng


def test_zero_heads_give_zero_field_and_identity_flow():
    """exe_flow_matching.py:81,86: Dense_4 / Dense_7 kernels zero-initialised -> v = bias terms only; with zero biases v = 0."""
    params, omega, rng = _small(head=0.0)
    for i in (4, 7):
        params["params"][f"Dense_{i}"]["bias"][:] = 0
    t = OT.PhiFour(3)
    x = rng.uniform(-1, 1, (5, 3))
    v, div = VF.field_and_div(params, omega, x, rng.uniform(0, 1, 5), t, None, None)
    assert (v == 0).all() and (div == 0).all()
    flow = OS.Flow(params, omega, t, hutch=False)
    y, ldj = flow.transform_and_logdet(None, x)
    assert np.array_equal(y, x) and (ldj == 0).all()


def test_jvp_and_divergence_match_finite_differences():
    params, omega, rng = _small()
    t = OT.four_mode()
    t3 = OT.PhiFour(3)
    x = rng.uniform(-0.9, 0.9, (4, 3)); tt = rng.uniform(0, 1, 4); z = rng.standard_normal((4, 3))
    for clip in (None, 1.0):
        v, c = VF.forward(params, omega, x, tt, t3.grad, clip, want_cache=True)
        jz = VF.jvp_x(params, c, z, lambda zz: t3.hvp(x, zz), clip)
        h = 1e-6
        fd = (VF.forward(params, omega, x + h * z, tt, t3.grad, clip) - VF.forward(params, omega, x - h * z, tt, t3.grad, clip)) / (2 * h)
        assert np.abs(fd - jz).max() < 1e-6 * max(1.0, np.abs(jz).max())
        # exact divergence = trace of the finite-difference Jacobian; Hutchinson form z.(Jz)
        _, div = VF.field_and_div(params, omega, x, tt, t3, None, clip)
        tr = np.zeros(4)
        for j in range(3):
            e = np.zeros_like(x); e[:, j] = h
            tr += ((VF.forward(params, omega, x + e, tt, t3.grad, clip) - VF.forward(params, omega, x - e, tt, t3.grad, clip)) / (2 * h))[:, j]
        assert np.abs(div - tr).max() < 1e-6 * max(1.0, np.abs(tr).max())
        _, dh = VF.field_and_div(params, omega, x, tt, t3, z, clip)
        assert np.allclose(dh, (z * jz).sum(1), rtol=1e-13)


def test_fm_batch_draws_and_loss_gradient():
    """cond_flow_fn (exe_flow_matching.py:151-169): key split 4-way, per-row reference keys, one [N,d] noise draw;
    loss = SUM over chains and dims of (v(x_t, t) - (x - ref))^2; gradient vs finite differences of every parameter block."""
    params, omega, rng = _small(seed=1)
    t3 = OT.PhiFour(3)
    ref = OT.IndepGaussian(3)
    x = rng.uniform(-1, 1, (6, 3))
    key = tf.PRNGKey(3)
    times, xt, target = VF.fm_batch(key, x, ref.sample, 1e-4, rng_dtype=np.float32)
    k_time, k_ref, k_gauss, _ = tf.split(key, 4)
    tm = tf.uniform(k_time, (6, 1), np.float32).astype(np.float64)
    rf = np.stack([tf.normal(k, (3,), np.float32) for k in tf.split(k_ref, 6)]).astype(np.float64)
    eps = tf.normal(k_gauss, (6, 3), np.float32).astype(np.float64)
    assert np.array_equal(times, tm[:, 0]) and np.array_equal(target, x - rf)
    assert np.allclose(xt, 1e-4 * eps + tm * x + (1 - tm) * rf, rtol=1e-15)
    loss, G = VF.fm_loss_and_grad(params, omega, xt, times, target, t3.grad, None)
    v = VF.forward(params, omega, xt, times, t3.grad, None)
    assert np.isclose(loss, ((v - target) ** 2).sum(), rtol=1e-14)
    h = 1e-6
    for i in range(8):
        for name in ("kernel", "bias"):
            a = params["params"][f"Dense_{i}"][name]
            idx = tuple(rng.integers(0, s) for s in a.shape)
            old = a[idx]
            a[idx] = old + h; lp, _ = VF.fm_loss_and_grad(params, omega, xt, times, target, t3.grad, None)
            a[idx] = old - h; lm, _ = VF.fm_loss_and_grad(params, omega, xt, times, target, t3.grad, None)
            a[idx] = old
            g = G["params"][f"Dense_{i}"][name][idx]
            assert abs((lp - lm) / (2 * h) - g) <= 2e-5 * max(1.0, abs(g)), (i, name)


def test_mala_step_as_coded():
    """mala.py:68-118, diffusions.py:22-33, proposal.py:104-186 written out for one chain."""
    t = OT.PhiFour(8)
    x = np.random.default_rng(4).uniform(-1, 1, (5, 8))
    st = OS.mala_init(x, t, 0.7)
    keys = tf.split(tf.PRNGKey(9), 5)
    h = 1e-3
    new, info, aux = OS.mala_step(keys, st, t, h, beta=0.7, rng_dtype=np.float32)
    for n in range(5):
        k_i, k_r = tf.split(keys[n])
        noise = tf.normal(k_i, (8,), np.float32).astype(np.float64)
        xn = x[n] + h * st.logdensity_grad[n] + np.sqrt(2 * h) * noise
        ln, gn = t.value_and_grad(xn[None], 0.7)
        e_new = -st.logdensity[n] + ((xn - x[n] - h * st.logdensity_grad[n]) ** 2).sum() / (4 * h)
        e_prev = -ln[0] + ((x[n] - xn - h * gn[0]) ** 2).sum() / (4 * h)
        p = min(1.0, np.exp(e_prev - e_new))
        u = float(tf.uniform(k_r, (), np.float32))
        assert np.allclose(info.proposed_position[n], xn, rtol=1e-14)
        assert np.isclose(info.acceptance_rate[n], p, rtol=1e-10)
        assert bool(info.is_accepted[n]) == (u < p)                                    # strict <, jax.random.bernoulli
        assert np.isclose(info.proposed_weight[n], np.exp(ln[0] + ((x[n] - xn - h * gn[0]) ** 2).sum() / (4 * h)), rtol=1e-10)
        exp_pos = xn if u < p else x[n]
        assert np.allclose(new.position[n], exp_pos, rtol=1e-14)
    # NaN energy difference -> -inf -> never accepted (proposal.py:105)
    bad = OS.MALAState(np.full((1, 8), np.nan), np.array([np.nan]), np.full((1, 8), np.nan))
    _, info, _ = OS.mala_step(keys[:1], bad, t, h)
    assert not info.is_accepted[0] and info.acceptance_rate[0] == 0


@pytest.mark.parametrize("hutch", [False, True])
def test_push_after_pull_is_identity_and_logdets_cancel(hutch):
    params, omega, rng = _small(seed=2, d=2, head=0.5)
    t = OT.four_mode()
    flow = OS.Flow(params, omega, t, hutch=hutch, rtol=1e-8, atol=1e-8)
    x = 8.0 + rng.standard_normal((4, 2))
    keys = tf.split(tf.PRNGKey(1), 4)
    z = rng.standard_normal((4, 2)) if hutch else None
    u, V0 = flow.inverse_and_logdet(keys, x, z=z)
    y, V1 = flow.transform_and_logdet(keys, u, z=z)
    assert np.abs(y - x).max() < 5e-5 and np.abs(V0 + V1).max() < 5e-5      # a relu field is only C0: kinks cost accuracy
    assert np.abs(u - x).max() > 1e-3                  # the flow actually moves points


def test_flow_mh_steps_as_coded():
    """random_walk_metropolis_hastings (:264-278) and indep_metropolis_hastings (:246-260): key roles, latent random walk of
    scale 2.38/sqrt(d), unclipped ratio, non-strict <=, proposed_weight = 0."""
    params, omega, rng = _small(seed=5, d=2, head=0.2)
    t = OT.four_mode()
    flow = OS.Flow(params, omega, t, hutch=False)
    x = 8.0 + rng.standard_normal((3, 2))
    st = OS.mala_init(x, t)
    keys = tf.split(tf.PRNGKey(2), 3)
    stats = {}
    new, info = OS.rw_flow_mh_step(keys, st, t, flow, 1.0, stats)
    for n in range(3):
        k_gen, k_acc, _, _ = tf.split(keys[n], 4)
        up = stats["u0"][n] + 2.38 / np.sqrt(2) * tf.normal(k_gen, (2,), np.float64)
        xp, Vp = flow.transform_and_logdet(None, up[None])
        lp = t.logprob(xp)[0]
        assert np.allclose(info.proposed_position[n], xp[0], atol=1e-9)
        ratio = np.exp(lp - Vp[0] - st.logdensity[n] - stats["V0"][n])
        assert np.isclose(info.acceptance_rate[n], ratio, rtol=1e-6)
        assert bool(info.is_accepted[n]) == (float(tf.uniform(k_acc, (), np.float64)) <= ratio)
    assert (info.proposed_weight == 0).all()
    new2, info2 = OS.indep_flow_mh_step(keys, st, t, flow, OT.IndepGaussian(2))
    assert (info2.proposed_weight == 0).all() and np.isfinite(info2.acceptance_rate).all()
    # dispatch (exe_flow_matching.py:304-313): count % (m+1) == 0 -> flow step
    _, i_m = OS.train_data_generator(tf.PRNGKey(3), st, 1, t, flow, 0.2, 2)
    _, i_f = OS.train_data_generator(tf.PRNGKey(3), st, 3, t, flow, 0.2, 2)
    assert (i_m.proposed_weight != 0).any() and (i_f.proposed_weight == 0).all()


def test_adamw_clip_apply_if_finite():
    """optax chain as configured (exe_flow_matching.py:129-137): first step u = -lr (sign(g) + wd p) clipped to [-1, 1];
    biases are not decayed; non-finite gradients skip the update and do not advance the inner state until more than 10 in a row."""
    p = {"params": {"Dense_0": {"kernel": np.array([[1.0, -2.0]], np.float32), "bias": np.array([0.5, 0.5], np.float32)}}}
    g = {"params": {"Dense_0": {"kernel": np.array([[0.3, -4.0]], np.float32), "bias": np.array([2.0, -1e-3], np.float32)}}}
    lr = OO.learning_rate_fn(10, 0, 0.1)
    assert [round(lr(s), 6) for s in (0, 1, 5, 10)] == [0.1, 0.09, 0.05, 0.0]
    opt = OO.AdamWClipIfFinite(p, lr, weight_decay=0.01)
    p1 = opt.update(g, p)
    k = p1["params"]["Dense_0"]["kernel"]; b = p1["params"]["Dense_0"]["bias"]
    assert np.allclose(k, [[1.0 - 0.1 * (1 + 0.01 * 1.0), -2.0 - 0.1 * (-1 + 0.01 * -2.0)]], rtol=1e-6)
    assert np.allclose(b, [0.5 - 0.1, 0.5 + 0.1], rtol=1e-5)                           # Adam's first step is lr * sign(g)
    assert opt.count == 1
    bad = {"params": {"Dense_0": {"kernel": np.array([[np.nan, 0.0]], np.float32), "bias": np.zeros(2, np.float32)}}}
    q = p1
    for i in range(10):
        q2 = opt.update(bad, q)
        assert q2 is q and opt.count == 1 and opt.notfinite_count == i + 1
    q3 = opt.update(bad, q)                         # the 11th consecutive failure is applied regardless
    assert q3 is not q and opt.count == 2 and opt.total_notfinite == 11
    big = OO.AdamWClipIfFinite(p, lambda s: 50.0)   # update clip: elementwise to [-1, 1]
    pb = big.update(g, p)
    assert np.allclose(np.abs(pb["params"]["Dense_0"]["bias"] - p["params"]["Dense_0"]["bias"]), 1.0)


def test_tempering_beta_bisection():
    """beta_fn (exe_flow_matching.py:391-402): ESS(beta) = alpha N by bisection on [prev, 1]."""
    rng = np.random.default_rng(0)
    ll = rng.standard_normal(512) * 40.0
    n, alpha = ll.size, 0.5

    def ess(beta, prev):
        w = np.exp((beta - prev) * ll - ((beta - prev) * ll).max()); w /= w.sum()
        return 1.0 / (w * w).sum()

    b1 = OO.tempering_beta(0.0, ll, alpha, dtype=np.float64)
    assert 0.0 < b1 < 1.0 and abs(ess(b1, 0.0) - alpha * n) < 0.5
    b2 = OO.tempering_beta(b1, ll, alpha, dtype=np.float64)
    assert b1 < b2 <= 1.0
    # ESS at beta = 1 already above alpha N: no bracket (sign 0) -> the lower end climbs to 1 - 2^-30 (1 - prev)
    flat = rng.standard_normal(512) * 1e-3
    b = OO.tempering_beta(0.25, flat, alpha, dtype=np.float64)
    assert abs(b - (1.0 - 2.0 ** -30 * 0.75)) < 1e-12


def test_cis_flow_step_as_coded():
    """conditional_importance_sampling (:280-296): weights, one categorical draw per chain, gradient carried over unchanged."""
    params, omega, rng = _small(seed=6, d=2, head=0.2)
    t, ref = OT.four_mode(), OT.IndepGaussian(2)
    flow = OS.Flow(params, omega, t, hutch=False)
    x = 8.0 + rng.standard_normal((4, 2))
    st = OS.mala_init(x, t)
    keys = tf.split(tf.PRNGKey(3), 4)
    K = 5
    stats = {}
    new, info = OS.cis_flow_step(keys, st, t, flow, ref, K, stats=stats)
    for n in range(4):
        k_sample, k_hp, k_h, k_choice = tf.split(keys[n], 4)
        u_prev, vol_prev = flow.inverse_and_logdet(None, x[n:n + 1])
        pw = np.exp(st.logdensity[n] - ref.logprob(u_prev)[0] - vol_prev[0])
        assert np.isclose(stats["prev_weight"][n], pw, rtol=1e-6)        # batched vs single-chain solve: BLAS summation order
        refs = np.stack([tf.normal(k, (2,), np.float64) for k in tf.split(k_sample, K)])
        samples, vols = flow.transform_and_logdet(None, refs)
        w = np.exp(t.logprob(samples) - ref.logprob(refs) - vols)
        norm = np.concatenate([[pw], w]) / (pw + w.sum())
        u = float(tf.uniform(k_choice, (1,), np.float64)[0])
        c = int(np.searchsorted(np.cumsum(norm), np.cumsum(norm)[-1] * (1 - u)))
        assert bool(info.is_accepted[n]) == (c > 0)
        assert np.isclose(info.acceptance_rate[n], norm[c], rtol=1e-6) and info.proposed_weight[n] == info.acceptance_rate[n]
        expect = samples[c - 1] if c > 0 else x[n]
        assert np.allclose(new.position[n], expect, atol=1e-9) and np.allclose(info.proposed_position[n], expect, atol=1e-9)
    assert np.array_equal(new.logdensity_grad, st.logdensity_grad)        # as coded: not recomputed
    # dispatch: num_importance_samples > 0 selects it on flow iterations
    _, i_f = OS.train_data_generator(tf.PRNGKey(3), st, 3, t, flow, 0.2, 2, num_importance_samples=K, ref=ref)
    assert (i_f.proposed_weight == i_f.acceptance_rate).all()


def test_adamw_restatement_matches_torch_adamw():
    """optax.adamw (decoupled decay, bias-corrected moments, eps outside the root) against torch.optim.AdamW, an independent
    implementation of the same update, for several steps with the schedule's learning rates (clip inactive: |update| << 1)."""
    import torch
    rng = np.random.default_rng(0)
    p0 = {"params": {"Dense_0": {"kernel": rng.standard_normal((6, 5)).astype(np.float32), "bias": rng.standard_normal(5).astype(np.float32)}}}
    lr_fn = OO.learning_rate_fn(100, 0, 1e-3)
    opt = OO.AdamWClipIfFinite(p0, lr_fn, weight_decay=1e-2)
    k = torch.nn.Parameter(torch.tensor(p0["params"]["Dense_0"]["kernel"]))
    b = torch.nn.Parameter(torch.tensor(p0["params"]["Dense_0"]["bias"]))
    topt = torch.optim.AdamW([{"params": [k], "weight_decay": 1e-2}, {"params": [b], "weight_decay": 0.0}], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    p = p0
    for step in range(6):
        g = {"params": {"Dense_0": {"kernel": rng.standard_normal((6, 5)).astype(np.float32), "bias": rng.standard_normal(5).astype(np.float32)}}}
        p = opt.update(g, p)
        for grp in topt.param_groups:
            grp["lr"] = float(lr_fn(step))
        k.grad = torch.tensor(g["params"]["Dense_0"]["kernel"]); b.grad = torch.tensor(g["params"]["Dense_0"]["bias"])
        topt.step()
        assert np.allclose(p["params"]["Dense_0"]["kernel"], k.detach().numpy(), rtol=2e-6, atol=2e-7)
        assert np.allclose(p["params"]["Dense_0"]["bias"], b.detach().numpy(), rtol=2e-6, atol=2e-7)


def test_activation_derivatives_by_finite_differences():
    """oracle.vector_field.activation: (act, act') pairs of jax.nn.{relu, tanh, elu, gelu (tanh approximation), swish}."""
    v = np.linspace(-4.0, 4.0, 401) + 1e-3
    h = 1e-6
    for name in ("relu", "tanh", "elu", "gelu", "swish"):
        f, df = VF.activation(name, v)
        fd = (VF.activation(name, v + h)[0] - VF.activation(name, v - h)[0]) / (2 * h)
        assert np.abs(df - fd).max() < 1e-6, name
    # definitions at known points: gelu(1) with the tanh approximation, swish(1) = sigmoid(1), elu(-1) = e^-1 - 1
    one = np.array([1.0])
    assert VF.activation("gelu", one)[0][0] == np.float64(0.5 * (1 + np.tanh(np.sqrt(2 / np.pi) * 1.044715)))
    assert abs(VF.activation("swish", one)[0][0] - 1 / (1 + np.exp(-1.0))) < 1e-15
    assert abs(VF.activation("elu", -one)[0][0] - (np.exp(-1.0) - 1)) < 1e-15
