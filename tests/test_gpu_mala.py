"""Fused MALA transition vs the oracle restatement of bblackjax/mcmc/mala.py (as coded).

Contract (north_star): log-densities within 1e-4 relative; accept decisions identical except
where |log alpha - log u| is inside the tolerance band."""
import numpy as np
import pytest
import torch

from oracle import samplers as OS, threefry as tf
from tests.helpers import key_dev, make_targets, rel_err, to_dev

pytestmark = pytest.mark.gpu

STEP = {"4-mode": 0.2, "gmm16": 0.2, "phi-four": 1e-4, "pines": 0.01}


@pytest.fixture(scope="module")
def targets(cuda, lib):
    return make_targets(cuda)


def _check(name, ot, dd, cuda, n, beta, per_chain, steps=3, rng_dtype=np.float32):
    from mfm_b200.bblackjax.mcmc import mala as M
    h = STEP[name]
    x0 = ot.init_positions(tf.PRNGKey(1), n, np.float32)
    fn = dd.tempered(beta)
    st_d = M.init(to_dev(x0, cuda), fn)
    st_o = OS.mala_init(x0.astype(np.float64), ot, beta)
    assert rel_err(st_d.logdensity.cpu().numpy(), st_o.logdensity) < 1e-4
    key = tf.PRNGKey(1024)
    kernel = M.build_kernel()
    n_flip = 0
    for it in range(steps):
        key, sub = tf.split(key)
        keys = tf.split(sub, n)
        # feed the oracle the device state (float64 copy) so each transition is compared in isolation
        st_in = OS.MALAState(st_d.position.cpu().numpy().astype(np.float64),
                             st_d.logdensity.cpu().numpy().astype(np.float64),
                             st_d.logdensity_grad.cpu().numpy().astype(np.float64))
        noise = tf.vmap_normal(np.stack([tf.split(k)[0] for k in keys]), ot.dim, rng_dtype).astype(np.float64)
        new_o, info_o, dbg = OS.mala_step(keys, st_in, ot, h, beta, noise=noise, rng_dtype=rng_dtype)
        if per_chain:
            st_d, info_d = kernel(key_dev(keys, cuda), st_d, fn, h)
        else:
            st_d, info_d = M.mala_step(fn, key_dev(sub, cuda), st_d, h, per_chain_keys=False)
        prop = info_d.proposed_position.cpu().numpy()
        assert rel_err(prop, info_o.proposed_position) < 1e-5, (name, it)
        p_d = info_d.acceptance_rate.cpu().numpy(); p_o = info_o.acceptance_rate
        acc_d = info_d.is_accepted.cpu().numpy(); acc_o = info_o.is_accepted
        # decisions may only differ inside the tolerance band around the threshold
        # (1e-5 - north_star's figure - where float32 can deliver it: the log-densities of the mixtures are O(10); for phi-four and
        # pines |l| ~ 2e3, whose float32 ulp alone is 1.2e-4)
        fac = 1e-5 if np.abs(st_in.logdensity).max() < 100.0 else 1e-4
        band = np.abs(p_o - dbg["u"]) < fac * np.maximum(1.0, np.abs(dbg["delta"]))
        assert ((acc_d == acc_o) | band).all(), (name, it)
        n_flip += int((acc_d != acc_o).sum())
        ok = np.isfinite(dbg["delta"])
        np.testing.assert_allclose(p_d[ok], p_o[ok], rtol=2e-3, atol=1e-5)
        same = acc_d == acc_o
        assert rel_err(st_d.logdensity.cpu().numpy()[same], new_o.logdensity[same]) < 1e-4
        assert rel_err(st_d.position.cpu().numpy()[same], new_o.position[same]) < 1e-5
        assert rel_err(st_d.logdensity_grad.cpu().numpy()[same], new_o.logdensity_grad[same]) < 1e-4
        w_d = info_d.proposed_weight.cpu().numpy(); w_o = info_o.proposed_weight
        fin = np.isfinite(w_o) & (w_o > 1e-30) & (w_o < 1e30)
        if fin.any():
            np.testing.assert_allclose(np.log(w_d[fin]), np.log(w_o[fin]), rtol=1e-4, atol=1e-3)
    return n_flip


@pytest.mark.parametrize("per_chain", [True, False])
def test_mala_small_targets(cuda, targets, per_chain):
    for name, ot, dd in targets[:3]:
        _check(name, ot, dd, cuda, 131, 1.0, per_chain)


def test_mala_x64_draws(cuda, lib, targets):
    """The reference as shipped runs with jax_enable_x64 (multi_modal.py:14): noise and accept uniforms are float64 draws
    (64 bits each).  With mfm_set_rng_x64(1) the fused MALA kernels consume exactly those draws (rounded to float32)."""
    lib.mfm_set_rng_x64(1)
    try:
        for name, ot, dd in targets[:3]:
            _check(name, ot, dd, cuda, 67, 1.0, False, rng_dtype=np.float64)
        name, ot, dd = targets[3]
        _check(name, ot, dd, cuda, 5, 1.0, True, steps=2, rng_dtype=np.float64)
    finally:
        lib.mfm_set_rng_x64(0)


def test_mala_tempered(cuda, targets):
    for name, ot, dd in targets[:3]:
        _check(name, ot, dd, cuda, 64, 0.25, False)


def test_mala_pines(cuda, targets):
    name, ot, dd = targets[3]
    _check(name, ot, dd, cuda, 130, 1.0, False, steps=2)
    _check(name, ot, dd, cuda, 16, 0.5, True, steps=1)


def test_mala_sharded_keys_match_single(cuda, targets):
    """chain_offset/n_total: two half-ensembles reproduce the full ensemble bit for bit."""
    from mfm_b200.bblackjax.mcmc import mala as M
    name, ot, dd = targets[2]
    n = 64
    x0 = to_dev(ot.init_positions(tf.PRNGKey(3), n, np.float32), cuda)
    fn = dd.tempered(1.0)
    st = M.init(x0, fn)
    key = key_dev(tf.PRNGKey(9), cuda)
    full, info = M.mala_step(fn, key, st, STEP[name], per_chain_keys=False)
    for lo, hi in ((0, 32), (32, 64)):
        part = M.MALAState(st.position[lo:hi].contiguous(), st.logdensity[lo:hi].contiguous(),
                           st.logdensity_grad[lo:hi].contiguous())
        new, inf = M.mala_step(fn, key, part, STEP[name], per_chain_keys=False, chain_offset=lo, n_total=n)
        assert torch.equal(new.position, full.position[lo:hi])
        assert torch.equal(inf.is_accepted, info.is_accepted[lo:hi])


def test_unknown_callable_raises(cuda, lib):
    from mfm_b200.bblackjax.mcmc import mala as M
    with pytest.raises(TypeError, match="no fallback"):
        M.init(torch.zeros(4, 2, device=cuda), lambda x: -(x * x).sum())
