"""Oracle targets (oracle/targets.py) pinned against what can be pinned without JAX:
closed forms restated straight from the reference formulas (scalar loops), the pines constants of SURVEY.md 8(c),
scipy's normal log-pdf, and finite differences of the analytic gradient / Hessian-vector product / Hessian diagonal
the CUDA kernels are later compared with.  float64 throughout (the reference's dtype with jax_enable_x64)."""
import numpy as np
import pytest
import scipy.stats

from oracle import targets as OT
from oracle import threefry as tf


def _fd_grad(f, x, h=1e-6):
    g = np.zeros_like(x)
    for j in range(x.shape[1]):
        e = np.zeros_like(x); e[:, j] = h
        g[:, j] = (f(x + e) - f(x - e)) / (2 * h)
    return g


def _targets():
    rng = np.random.default_rng(0)
    counts = rng.poisson(0.08, 64).astype(np.float64)
    return {"4-mode": (OT.four_mode(), 8.0 + rng.standard_normal((5, 2))),
            "gmm16": (OT.gmm16(), OT.gmm16().modes[:6] + 0.3 * rng.standard_normal((6, 2))),
            "phi-four": (OT.PhiFour(16), rng.uniform(-1, 1, (4, 16))),
            "pines8x8": (OT.LogGaussianCoxPines(64, counts), 3.0 + 0.5 * rng.standard_normal((3, 64)))}


@pytest.mark.parametrize("name", ["4-mode", "gmm16", "phi-four", "pines8x8"])
@pytest.mark.parametrize("beta", [1.0, 0.37])
def test_gradient_hvp_hdiag_match_finite_differences(name, beta):
    t, x = _targets()[name]
    val, grad = t.value_and_grad(x, beta)
    assert np.allclose(val, t.logprob(x, beta), rtol=0, atol=0)
    fd = _fd_grad(lambda y: t.logprob(y, beta), x)
    assert np.abs(fd - grad).max() <= 2e-6 * max(1.0, np.abs(grad).max())
    if beta == 1.0:
        rng = np.random.default_rng(1)
        z = rng.standard_normal(x.shape)
        h = 1e-6
        hv_fd = (t.grad(x + h * z) - t.grad(x - h * z)) / (2 * h)
        assert np.abs(hv_fd - t.hvp(x, z)).max() <= 5e-6 * max(1.0, np.abs(hv_fd).max())
        hd = np.stack([t.hvp(x, np.eye(x.shape[1])[j][None].repeat(x.shape[0], 0))[:, j] for j in range(x.shape[1])], 1)
        assert np.abs(hd - t.hdiag(x)).max() <= 1e-9 * max(1.0, np.abs(hd).max())


def test_phi_four_as_coded():
    """distributions.py:131-160: U = a d / 2 * sum_{i=0..d} (xp[i+1]-xp[i])^2 with zero padding, V = sum (1-x^2)^2 / (4 a d)."""
    d, a, beta = 64, 0.1, 20.0
    x = np.random.default_rng(2).uniform(-1, 1, (3, d))
    t = OT.PhiFour(d, a, beta)
    for n in range(3):
        xp = [0.0] + list(x[n]) + [0.0]
        U = a * d / 2 * sum((xp[i + 1] - xp[i]) ** 2 for i in range(d + 1))
        V = sum((1 - v * v) ** 2 for v in x[n]) / (4 * a * d)
        assert abs(t.loglik(x)[n] + beta * (U + V)) <= 1e-12 * beta * (U + V)
    assert (t.logprior(x) == 0).all()
    assert abs(a * d - 6.4) < 1e-12                     # coef = 6.4 at d = 64


def test_gaussian_mixture_as_coded_and_underflow():
    """distributions.py:58-61: log sum_k w_k prod_j N(x_j; m_kj, sqrt(cov_kj)) in the probability domain."""
    t = OT.four_mode()
    x = np.array([[8.0, 8.0], [0.3, -0.2], [7.0, -9.0]])
    ref = np.log(sum(0.25 * scipy.stats.norm.pdf(x[:, 0], m[0], 1.0) * scipy.stats.norm.pdf(x[:, 1], m[1], 1.0) for m in t.modes))
    assert np.allclose(t.logprob(x), ref, rtol=1e-13)
    g = OT.gmm16()
    assert g.modes.shape == (16, 2) and np.isclose(g.weights.sum(), 1.0) and (np.abs(g.modes) <= 12.8).all()
    # no log-sum-exp: float32 underflows to -inf ~14 sigma out, float64 ~38 sigma out
    far32 = np.array([[8.0 + 15.0, 8.0]], np.float32)
    assert np.isneginf(t.logprob(far32))[0] and np.isfinite(t.logprob(far32.astype(np.float64)))[0]
    assert np.isneginf(t.logprob(np.array([[8.0 + 40.0, 8.0 + 40.0]])))[0]


def test_indep_gaussian_matches_scipy_and_sampling_uses_row_keys():
    t = OT.IndepGaussian(5)
    x = np.random.default_rng(3).standard_normal((4, 5))
    assert np.allclose(t.logprob(x), scipy.stats.norm.logpdf(x).sum(1), rtol=1e-13)
    assert np.allclose(t.grad(x), -x)
    keys = tf.split(tf.PRNGKey(7), 3)
    s = t.sample(keys, np.float32)
    for i in range(3):                                  # vmap(sample_model)(keys): row i = normal(keys[i], (d,))
        assert np.array_equal(s[i], tf.normal(keys[i], (5,), np.float32))


def test_pines_constants_match_survey():
    """SURVEY.md 8(c) golden constants of the 40 x 40 Finnish-pines LGCP target."""
    t = OT.LogGaussianCoxPines(1600)
    assert t.counts.sum() == 126 and t.counts.max() == 3 and (t.counts > 0).sum() == 111
    assert abs(t.mu - 3.881281906951478) < 1e-14
    assert abs(t.half_log_det - 225.70547889445027) < 1e-8
    assert abs(t.log_norm - (-1696.0071320219267)) < 1e-8
    assert abs(np.linalg.cond(t.K) - 27.62) < 0.01
    # K_ij = 1.91 exp(-|p_i - p_j| / (40/33)), points in itertools.product order
    assert abs(t.K[0, 1] - 1.91 * np.exp(-1 / (40 / 33))) < 1e-14 and abs(t.K[0, 40] - t.K[0, 1]) < 1e-14
    assert abs(t.K[0, 41] - 1.91 * np.exp(-np.sqrt(2) / (40 / 33))) < 1e-14


def test_pines_prior_solve_form_equals_precision_form():
    """The CUDA path evaluates the prior as x K^-1 (one GEMM); the reference does two triangular solves."""
    t = OT.LogGaussianCoxPines(1600)
    x = t.init_positions(tf.PRNGKey(1), 3, np.float64)
    r = x - t.mu
    q = -0.5 * np.einsum("ni,ij,nj->n", r, t.Kinv, r) + t.log_norm
    assert np.allclose(t.logprior(x), q, rtol=1e-10)
    assert np.abs(t.grad_logprior(x) + r @ t.Kinv).max() < 1e-9
    lik = (x * t.counts - np.exp(x) / 1600).sum(1)
    assert np.allclose(t.loglik(x), lik, rtol=1e-13)
    # tempered density (exe_flow_matching.py:301): beta * loglik + logprior
    assert np.allclose(t.logprob(x, 0.25), 0.25 * lik + t.logprior(x), rtol=1e-13)


def test_initial_positions_follow_the_reference_recipes():
    key = tf.PRNGKey(5)
    p4 = OT.PhiFour(8).init_positions(key, 4, np.float32)
    for i, k in enumerate(tf.split(key, 4)):            # distributions.py:162-164
        assert np.array_equal(p4[i], tf.uniform(k, (8,), np.float32) * np.float32(2) - np.float32(1))
    g = OT.four_mode().init_positions(key, 4, np.float32)
    for i, k in enumerate(tf.split(key, 4)):            # distributions.py:69-71
        assert np.array_equal(g[i], tf.normal(k, (2,), np.float32))


def test_pines_prior_is_the_multivariate_normal_density():
    """unwhitened_prior_log_density (distributions.py:299-303) = log N(x; mu 1, K), checked against scipy's MVN on the real K."""
    t = OT.LogGaussianCoxPines(1600)
    x = t.init_positions(tf.PRNGKey(2), 2, np.float64)
    ref = scipy.stats.multivariate_normal(mean=np.full(1600, t.mu), cov=t.K, allow_singular=False).logpdf(x)
    assert np.allclose(t.logprior(x), ref, rtol=1e-10)


def test_whitened_pines_oracle_by_finite_differences():
    """LogGaussianCoxPinesWhitened: analytic gradient / Hessian-vector product / Hessian diagonal of log-likelihood + log-prior in
    the white-noise parameterisation (distributions.py:276-297), and its link to the unwhitened target: the two log-densities
    differ by the constant log |det L| at corresponding points f = L e + mu."""
    from oracle import targets as OT
    t = OT.LogGaussianCoxPinesWhitened(1600)
    u = OT.LogGaussianCoxPines(1600)
    rng = np.random.default_rng(0)
    e = rng.standard_normal((2, 1600)); z = rng.standard_normal((2, 1600))
    h = 1e-5
    g = t.grad(e)
    assert abs(((t.logprob(e + h * z) - t.logprob(e - h * z)) / (2 * h) - (g * z).sum(1)) / (g * z).sum(1)).max() < 1e-6
    hv = t.hvp(e, z)
    assert np.abs(hv - (t.grad(e + h * z) - t.grad(e - h * z)) / (2 * h)).max() < 1e-6 * np.abs(hv).max()
    ej = np.zeros((1, 1600)); ej[0, 7] = 1
    assert abs(t.hdiag(e)[0, 7] - ((t.grad(e[:1] + h * ej) - t.grad(e[:1] - h * ej)) / (2 * h))[0, 7]) < 1e-6
    f = e @ t.L.T + t.mu
    assert np.allclose(t.logprob(e) - u.logprob(f), t.half_log_det, rtol=0, atol=1e-8)


def test_phi_four_base_reference_distribution():
    """PhiFourBase (distributions.py:168-226, 'coupled'): the Gaussian whose precision is the quadratic part of the phi-four
    action - checked against the action itself, against scipy's multivariate normal, and the sampler against its covariance."""
    d = 12
    ref = OT.PhiFourBase(d, 0.1, 20.0)
    c = 0.1 * d
    P = 20.0 * (np.diag(np.full(d, 2 * c + 1 / c)) - c * (np.eye(d, k=1) + np.eye(d, k=-1)))
    assert np.allclose(ref.prec, P)
    # the quadratic part of PhiFour's energy (Dirichlet boundary, value 0): c/2 sum (x_{i+1} - x_i)^2 with zero padding + x^2 / (2 c)
    x = np.random.default_rng(0).standard_normal((5, d))
    xp = np.pad(x, ((0, 0), (1, 1)))
    quad = 20.0 * (0.5 * c * (np.diff(xp, axis=1) ** 2).sum(1) + (x ** 2).sum(1) / (2 * c))
    assert np.allclose(0.5 * np.einsum("ni,ij,nj->n", x, ref.prec, x), quad)
    mvn = scipy.stats.multivariate_normal(np.zeros(d), np.linalg.inv(P))
    assert np.allclose(ref.loglik(x), mvn.logpdf(x), rtol=1e-12, atol=1e-10)
    eps = 1e-6
    g_fd = np.stack([(ref.loglik(x + eps * np.eye(d)[i]) - ref.loglik(x - eps * np.eye(d)[i])) / (2 * eps) for i in range(d)], 1)
    assert np.allclose(ref.grad_loglik(x), g_fd, rtol=1e-6, atol=1e-6)
    assert np.allclose(ref.chol_cov @ ref.chol_cov.T, np.linalg.inv(P))
    keys = tf.split(tf.PRNGKey(3), 4)
    s = ref.sample(keys, np.float64)
    z = tf.vmap_normal(keys, d, np.float64)
    assert np.allclose(s, z @ ref.chol_cov.T)                  # sample_model(key) = chol_cov @ normal(key, (d,)), row by row
