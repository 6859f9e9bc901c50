"""Target log-density + gradient (mala.init) vs the float64 oracle: 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

from oracle import threefry as tf
from tests.helpers import make_targets, rel_err, to_dev

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def targets(cuda, lib):
    return make_targets(cuda)


@pytest.mark.parametrize("beta", [1.0, 0.37])
def test_value_and_grad_all_targets(cuda, targets, beta):
    for name, ot, dd in targets:
        n = 257 if ot.dim < 1000 else 130
        x64 = ot.init_positions(tf.PRNGKey(1), n, np.float32).astype(np.float64)
        if name in ("4-mode", "gmm16"):
            x64 = x64 * 4
        l_ref, g_ref = ot.value_and_grad(x64, beta)
        l, g, ll = dd.tempered(beta).value_and_grad(to_dev(x64, cuda), want_loglik=True)
        assert rel_err(l.cpu().numpy(), l_ref) < RTOL, name
        assert rel_err(g.cpu().numpy(), g_ref) < RTOL, name
        assert rel_err(ll.cpu().numpy(), ot.loglik(x64)) < RTOL, name


def test_reference_method_names(cuda, targets):
    for name, ot, dd in targets:
        x64 = ot.init_positions(tf.PRNGKey(2), 5, np.float32).astype(np.float64)
        xd = to_dev(x64, cuda)
        assert rel_err(dd.logprob(xd).cpu().numpy(), ot.logprob(x64)) < RTOL
        assert rel_err(dd.loglik(xd).cpu().numpy(), ot.loglik(x64)) < RTOL
        assert dd.logprob(xd[0]).dim() == 0


def test_gmm_underflow_as_coded(cuda, targets):
    """Far from every mode the probability-domain sum underflows: -inf value, NaN gradient
    (distributions.py:59-61 has no log-sum-exp)."""
    _, ot, dd = targets[0]
    x = np.array([[200.0, 200.0]], np.float32)
    l, g = dd.tempered(1.0).value_and_grad(to_dev(x, cuda))
    assert np.isneginf(l.cpu().numpy()).all() and np.isnan(g.cpu().numpy()).all()
    with np.errstate(all="ignore"):
        assert np.isneginf(ot.loglik(x)).all()


def test_initialize_model_matches_oracle(cuda, targets):
    from mfm_b200 import random as mr
    for name, ot, dd in targets:
        n = 33 if ot.dim < 1000 else 9
        dd.initialize_model(mr.PRNGKey(59049, cuda), n)
        exp = ot.init_positions(tf.PRNGKey(59049), n, np.float32)
        got = dd.init_params.cpu().numpy()
        assert got.shape == exp.shape
        assert np.abs(got - exp).max() <= 2e-5 * max(1.0, np.abs(exp).max()), name


@pytest.mark.parametrize("n", [9, 300])
def test_whitened_pines(cuda, lib, n):
    """LogGaussianCoxPines(use_whitened=True) (distributions.py:276-297): state = white noise e, latents L e + mu.  Value,
    gradient and log-likelihood (tempered), one MALA transition, and the vector field with its Hutchinson divergence (the
    Hessian-vector product goes through two more GEMMs against the Cholesky factor) against the oracle; n = 300 takes the
    tcgen05 GEMM path, n = 9 the warp-level one."""
    from types import SimpleNamespace
    from mfm_b200 import distributions as D, exe_flow_matching as E
    from mfm_b200.bblackjax.mcmc import mala as M
    from oracle import samplers as OS, targets as OT, vector_field as VF
    from tests.helpers import key_dev
    ot = OT.LogGaussianCoxPinesWhitened(1600)
    dd = D.LogGaussianCoxPines(1600, use_whitened=True, device=cuda)
    rng = np.random.default_rng(n)
    e = rng.standard_normal((n, 1600))
    for beta in (1.0, 0.4):
        l_ref, g_ref = ot.value_and_grad(e, beta)
        l, g, ll = dd.tempered(beta).value_and_grad(to_dev(e, cuda), want_loglik=True)
        assert rel_err(l.cpu().numpy(), l_ref) < RTOL and rel_err(g.cpu().numpy(), g_ref) < RTOL
        assert rel_err(ll.cpu().numpy(), ot.loglik(e)) < RTOL
    # MALA (mala.py:86-118) on the whitened target
    fn = dd.tempered(1.0)
    st_d = M.init(to_dev(e, cuda), fn)
    st_o = OS.mala_init(e, ot)
    keys = tf.split(tf.PRNGKey(5), n)
    new_o, info_o, dbg = OS.mala_step(keys, st_o, ot, 0.01, rng_dtype=np.float32)
    new_d, info_d = M.build_kernel()(key_dev(keys, cuda), st_d, fn, 0.01)
    assert rel_err(info_d.proposed_position.cpu().numpy(), info_o.proposed_position) < 1e-5
    acc_d, acc_o = info_d.is_accepted.cpu().numpy(), info_o.is_accepted
    band = np.abs(info_o.acceptance_rate - dbg["u"]) < 1e-4 * np.maximum(1.0, np.abs(dbg["delta"]))
    assert ((acc_d == acc_o) | band).all()
    same = acc_d == acc_o
    assert rel_err(new_d.logdensity.cpu().numpy()[same], new_o.logdensity[same]) < RTOL
    assert rel_err(new_d.logdensity_grad.cpu().numpy()[same], new_o.logdensity_grad[same]) < RTOL
    # vector field + Hutchinson divergence
    H = 128
    prm = VF.init_params(np.random.default_rng(3), 1600, H, 128, head_scale=0.1)
    omega = np.random.default_rng(4).standard_normal(128).astype(np.float32)
    model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", 1.0)
    P = E.VectorFieldParams(1600, H, 128, cuda).load_dict(prm)
    t = np.linspace(0.0, 1.0, n); z = rng.standard_normal(e.shape)
    v_ref, div_ref = VF.field_and_div(prm, omega, e, t, ot, z, 1.0)
    v, div = model.apply(P, to_dev(e, cuda), to_dev(t, cuda), to_dev(z, cuda), hutch=True, want_div=True)
    assert rel_err(v.cpu().numpy(), v_ref) < RTOL
    assert np.abs(div.cpu().numpy() - div_ref).max() < RTOL * max(np.abs(div_ref).max(), 1.0)
