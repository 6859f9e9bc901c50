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
