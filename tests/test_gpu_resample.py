"""Final sampling (exe_flow_matching.py:453-459) on the device vs the oracle: jax.random.choice restatement bit-exact on
identical float32 weights, importance weights to float32 rounding, the whole sample_flow path within the ODE tolerance."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import resample as OR, samplers as OS, targets as OT, threefry as tf, vector_field as VF
from tests.helpers import key_dev, to_dev, ulp_diff_f32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_pop,n_draw", [(1, 5), (5, 1), (33, 1000), (12800, 12800), (1027, 4097), (100000, 777)])
def test_choice_indices_bit_exact(cuda, lib, n_pop, n_draw):
    from mfm_b200 import random as mr
    rng = np.random.default_rng(n_pop + n_draw)
    w = np.exp(rng.standard_normal(n_pop) * 3).astype(np.float32)
    w[rng.random(n_pop) < 0.2] = 0.0                     # underflowed weights
    if w.sum() == 0:
        w[0] = 1.0
    key = tf.PRNGKey(n_pop)
    ref, aux = OR.choice_indices(key, n_pop, n_draw, w)
    got = mr.choice(key_dev(key, cuda), n_pop, (n_draw,), to_dev(w, cuda)).cpu().numpy()
    assert np.array_equal(got, ref)


def test_choice_gathers_rows(cuda, lib):
    from mfm_b200 import random as mr
    rng = np.random.default_rng(0)
    a = rng.standard_normal((300, 7)).astype(np.float32)
    w = rng.random(300).astype(np.float32)
    key = tf.PRNGKey(5)
    rows_ref, idx_ref = OR.choice(key, a, 450, w)
    rows, idx = mr.choice(key_dev(key, cuda), to_dev(a, cuda), (450,), to_dev(w, cuda), return_index=True)
    assert np.array_equal(idx.cpu().numpy(), idx_ref) and np.array_equal(rows.cpu().numpy(), rows_ref)


def test_importance_weights(cuda, lib):
    from mfm_b200 import _lib
    rng = np.random.default_rng(1)
    n = 5000
    l, r, v = (rng.standard_normal(n).astype(np.float32) * s for s in (30.0, 3.0, 5.0))
    ld, rd, vd = to_dev(l, cuda), to_dev(r, cuda), to_dev(v, cuda)
    lw, w = torch.empty(n, device=cuda), torch.empty(n, device=cuda)
    _lib.check(lib.mfm_importance_weights(_lib.ptr(ld), _lib.ptr(rd), _lib.ptr(vd), n, _lib.ptr(lw), _lib.ptr(w), _lib.stream()))
    lw_ref = l - r - v                                    # float32, same association as the reference expression
    assert np.array_equal(lw.cpu().numpy(), lw_ref)
    w_ref = np.exp((lw_ref - lw_ref.max()).astype(np.float64))
    assert np.abs(w.cpu().numpy() - w_ref).max() <= 2e-6 and w.max().item() == 1.0


@pytest.mark.parametrize("hutch", [False, True])
def test_sample_flow_matches_oracle(cuda, lib, hutch):
    from mfm_b200 import distributions as D, exe_flow_matching as E
    ot = OT.four_mode()
    dd = D.GaussianMixture(ot.modes, ot.covs, ot.weights, device=cuda)
    H, n = 128, 96
    rng = np.random.default_rng(7)
    params = VF.init_params(rng, 2, H, 128, head_scale=0.2)
    omega = rng.standard_normal(128).astype(np.float32)
    model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", None)
    P = E.VectorFieldParams(2, H, 128, cuda).load_dict(params)
    args = SimpleNamespace(hutchs=hutch, num_importance_samples=0, mcmc_per_flow_steps=10, step_size=0.2)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=5)
    _, _, transform_and_logdet = E.create_train_data_gn(dd, model, opts, args)
    ref_d = E.ref_dists["stdgauss"](2, device=cuda)
    key = tf.PRNGKey(3)
    got = E.sample_flow(key_dev(key, cuda), dd, ref_d, transform_and_logdet, P, n)
    flow = OS.Flow(params, omega, ot, hutch, 1e-5, 1e-5, 1000, None, tuple(np.linspace(0, 1, 5)), rng_dtype=np.float32)
    ref = OR.sample_flow(key, ot, OT.IndepGaussian(2), flow, n)
    assert ulp_diff_f32(got["u"].cpu().numpy(), ref["u"].astype(np.float32)).max() <= 4      # normals: erf_inv to 4 ulp
    fs = got["flow_samples"].cpu().numpy()
    # adaptive solve of a relu field whose score switches sharply between the four modes: float32 and float64 solutions
    # differ by ~1e-3 once their step sequences diverge (tests/test_gpu_flow.py states the same limit)
    assert np.abs(fs - ref["flow_samples"]).max() <= 5e-3 * np.abs(ref["flow_samples"]).max()
    lw, lw_ref = got["log_weights"].cpu().numpy(), ref["log_weights"]
    fin = np.isfinite(lw_ref)
    assert (np.isfinite(lw) == fin).all()
    assert np.abs(lw[fin] - lw_ref[fin]).max() <= 0.05 + 2e-3 * np.abs(lw_ref[fin]).max()     # |grad log pi| ~ 10 x the position error
    # resampling of the DEVICE weights is the oracle's choice() bit for bit; against the oracle's own weights it may
    # differ where a draw falls within the weights' rounding of a CDF boundary
    k_c = tf.split(key)[1]
    idx_same_w, _ = OR.choice_indices(k_c, n, n, got["weights"].cpu().numpy())
    idx = got["indices"].cpu().numpy()
    assert np.array_equal(idx, idx_same_w)
    assert (idx == ref["indices"]).mean() >= 0.9
    assert np.array_equal(got["exact_samples"].cpu().numpy(), fs[idx])
