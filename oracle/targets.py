"""Oracle: target log-densities (batched NumPy restatement).  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/distributions.py:42-77 (GaussianMixture), :80-97 (IndepGaussian),
:114-165 (PhiFour), :231-314 (LogGaussianCoxPines) and cox_process_utils.py:28-165.
The reference differentiates these with jax.grad / jax.jvp / jax.jacfwd; here the gradient,
Hessian-vector product and Hessian diagonal are written out analytically and verified by finite
differences in tests/test_oracle_targets.py.

All functions are batched: x is [N, d]; values are [N]; grads [N, d].  dtype follows x
(float64 = the reference as shipped with jax_enable_x64, multi_modal.py:14; float32 = the
parity mode the CUDA path computes in).
"""
from __future__ import annotations

import os

import numpy as np

from . import threefry as tf

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "mfm_b200", "data", "pines_counts_40x40.txt")


class Target:
    """loglik/logprior split as the reference uses it (exe_flow_matching.py:301):
    logprob_beta(x) = beta * loglik(x) + logprior(x)."""
    dim: int

    def loglik(self, x):
        raise NotImplementedError

    def logprior(self, x):
        return np.zeros(x.shape[0], x.dtype)

    def grad_loglik(self, x):
        raise NotImplementedError

    def grad_logprior(self, x):
        return np.zeros_like(x)

    def hvp_loglik(self, x, z):
        raise NotImplementedError

    def hvp_logprior(self, x, z):
        return np.zeros_like(x)

    def hdiag_loglik(self, x):
        raise NotImplementedError

    def hdiag_logprior(self, x):
        return np.zeros_like(x)

    # -- derived -----------------------------------------------------------------------------
    def logprob(self, x, beta=1.0):
        b = x.dtype.type(beta)
        return b * self.loglik(x) + self.logprior(x)

    def value_and_grad(self, x, beta=1.0):
        b = x.dtype.type(beta)
        return (b * self.loglik(x) + self.logprior(x),
                b * self.grad_loglik(x) + self.grad_logprior(x))

    def grad(self, x):
        return self.grad_loglik(x) + self.grad_logprior(x)

    def hvp(self, x, z):
        return self.hvp_loglik(x, z) + self.hvp_logprior(x, z)

    def hdiag(self, x):
        return self.hdiag_loglik(x) + self.hdiag_logprior(x)


class GaussianMixture(Target):
    """distributions.py:42-67.  logprob = log sum_k w_k prod_j N(x_j; m_kj, s_kj), s = sqrt(covs).
    Computed in the probability domain exactly as coded (no log-sum-exp): underflows to -inf."""

    def __init__(self, modes, covs, weights):
        self.modes = np.asarray(modes, np.float64)      # [K, d]
        self.covs = np.asarray(covs, np.float64)        # [K, d] (diagonal entries)
        self.chol = np.sqrt(self.covs)                  # distributions.py:51
        self.weights = np.asarray(weights, np.float64)  # [K]
        self.dim = self.modes.shape[1]

    def _parts(self, x):
        dt = x.dtype
        m, s, w = self.modes.astype(dt), self.chol.astype(dt), self.weights.astype(dt)
        s2 = s * s
        diff = x[:, None, :] - m[None]                                  # [N,K,d]
        # jax.scipy.stats.norm.logpdf: -(log(2 pi s^2) + (x-m)^2/s^2)/2 ; pdf = exp(logpdf)
        logpdf = -(np.log(dt.type(2 * np.pi) * s2)[None] + diff * diff / s2[None]) / dt.type(2)
        with np.errstate(under="ignore"):
            pk = w[None] * np.prod(np.exp(logpdf), axis=2)              # [N,K]
        return pk, diff, s2

    def loglik(self, x):
        pk, _, _ = self._parts(x)
        with np.errstate(divide="ignore"):
            return np.log(pk.sum(1))

    def grad_loglik(self, x):
        pk, diff, s2 = self._parts(x)
        S = pk.sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            return (pk[:, :, None] * (-diff / s2[None])).sum(1) / S[:, None]

    def _hess(self, x):
        pk, diff, s2 = self._parts(x)
        S = pk.sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            r = pk / S[:, None]                                          # [N,K]
            gk = -diff / s2[None]                                        # [N,K,d]
            g = (r[:, :, None] * gk).sum(1)                              # [N,d]
            H = np.einsum("nk,nki,nkj->nij", r, gk, gk)
            H -= np.einsum("ni,nj->nij", g, g)
            idx = np.arange(self.dim)
            H[:, idx, idx] += (r[:, :, None] * (-1.0 / s2[None])).sum(1)
        return H

    def hvp_loglik(self, x, z):
        return np.einsum("nij,nj->ni", self._hess(x), z)

    def hdiag_loglik(self, x):
        H = self._hess(x)
        idx = np.arange(self.dim)
        return H[:, idx, idx]

    def init_positions(self, key, n, dtype=np.float32):
        """initialize_model (distributions.py:69-71)."""
        return tf.vmap_normal(tf.split(key, n), self.dim, dtype)


def four_mode():
    """multi_modal.py:78-85."""
    modes = 8.0 * np.array([[1, 1], [1, -1], [-1, 1], [-1, -1]], np.float64)
    return GaussianMixture(modes, np.ones((4, 2)), np.ones(4) / 4)


def gmm16_constants():
    """Frozen 16-mode constants: the reference draws them from PRNGKey(0) with
    jax.random.dirichlet (multi_modal.py:39-45); gamma sampling is a further unpinned algorithm,
    so modes/covs follow the threefry oracle and the weights use a fixed NumPy Dirichlet draw.
    Values live in mfm_b200/data/gmm16.txt (written by tests/golden/make_gmm16.py)."""
    p = os.path.join(os.path.dirname(_DATA), "gmm16.txt")
    a = np.loadtxt(p)
    return a[:, 0:2], a[:, 2:4], a[:, 4]


def gmm16():
    return GaussianMixture(*gmm16_constants())


class IndepGaussian(Target):
    """distributions.py:80-97 (reference/base distribution, mean 0 var 1 by default)."""

    def __init__(self, dim, mean=0.0, var=1.0):
        self.dim, self.mean, self.std = dim, mean, np.sqrt(var)

    def loglik(self, x):
        dt = x.dtype
        s2 = dt.type(self.std ** 2)
        d = x - dt.type(self.mean)
        return (-(np.log(dt.type(2 * np.pi) * s2) + d * d / s2) / dt.type(2)).sum(1)

    def grad_loglik(self, x):
        return -(x - x.dtype.type(self.mean)) / x.dtype.type(self.std ** 2)

    def sample(self, keys, dtype=np.float32):
        """vmap(sample_model)(keys) (distributions.py:96-97)."""
        dt = np.dtype(dtype)
        return dt.type(self.mean) + dt.type(self.std) * tf.vmap_normal(keys, self.dim, dtype)


class PhiFourBase(Target):
    """distributions.py:168-226, prior_type='coupled' (the default; ref_dists['phifour'], exe_flow_matching.py:53): the Gaussian
    N(0, P^-1) with the tridiagonal precision P = beta (diag(2 c + 1 / c) - c (sub + super diagonal)), c = alpha dim - the
    quadratic part of the phi-four action.  sample_model = chol_cov @ normal(key, (dim,)) with chol_cov = (L^-1)^T, L the lower
    Cholesky factor of P.  ('coupled_pbc' cannot be constructed in the reference: it assigns into jnp arrays and divides dim to a
    float.)  Oracle only: the device path supports diagonal reference distributions."""

    def __init__(self, dim, alpha=0.1, beta=20.0):
        self.dim = dim
        c = alpha * dim
        prec = np.eye(dim) * (3 * c + 1 / c)
        ones = np.ones((dim, dim))
        prec -= c * np.triu(np.triu(ones, k=-1).T, k=-1)          # as coded: a tridiagonal band of ones (main diagonal included)
        self.prec = beta * prec
        sign, logabs = np.linalg.slogdet(self.prec)
        self.prior_log_det = -sign * logabs                       # = log det of the covariance
        L = np.linalg.cholesky(self.prec)
        import scipy.linalg
        self.chol_cov = scipy.linalg.solve_triangular(L, np.eye(dim), lower=True).T

    def loglik(self, x):
        dt = x.dtype
        P = self.prec.astype(dt)
        q = np.einsum("ni,ij,nj->n", x, P, x)
        return -dt.type(0.5) * q - dt.type(0.5) * dt.type(self.dim * np.log(2 * np.pi) + self.prior_log_det)

    def grad_loglik(self, x):
        return -(x @ self.prec.astype(x.dtype))                   # P is symmetric

    def sample(self, keys, dtype=np.float32):
        """vmap(sample_model)(keys)."""
        z = tf.vmap_normal(keys, self.dim, dtype)
        return z @ self.chol_cov.astype(np.dtype(dtype)).T


class PhiFour(Target):
    """distributions.py:114-165, Dirichlet b.c. with value 0, no tilt."""

    def __init__(self, dim, a=0.1, beta=20.0):
        self.dim, self.a, self.beta = dim, a, beta

    def loglik(self, x):
        dt = x.dtype
        coef = dt.type(self.a * self.dim)
        xp = np.pad(x, ((0, 0), (1, 1)))
        diffs = xp[:, 1:] - xp[:, :-1]
        U = (diffs * diffs).sum(1) / dt.type(2) * coef
        q = dt.type(1.0) - x * x
        V = (q * q).sum(1) / dt.type(4) / coef
        return -dt.type(self.beta) * (U + V)

    def grad_loglik(self, x):
        dt = x.dtype
        coef = dt.type(self.a * self.dim)
        xp = np.pad(x, ((0, 0), (1, 1)))
        lap = dt.type(2) * x - xp[:, :-2] - xp[:, 2:]
        return -dt.type(self.beta) * (coef * lap - x * (dt.type(1) - x * x) / coef)

    def hvp_loglik(self, x, z):
        dt = x.dtype
        coef = dt.type(self.a * self.dim)
        zp = np.pad(z, ((0, 0), (1, 1)))
        lapz = dt.type(2) * z - zp[:, :-2] - zp[:, 2:]
        return -dt.type(self.beta) * (coef * lapz - (dt.type(1) - dt.type(3) * x * x) * z / coef)

    def hdiag_loglik(self, x):
        dt = x.dtype
        coef = dt.type(self.a * self.dim)
        return -dt.type(self.beta) * (dt.type(2) * coef - (dt.type(1) - dt.type(3) * x * x) / coef)

    def init_positions(self, key, n, dtype=np.float32):
        """distributions.py:162-164: uniform(k,(d,))*2-1."""
        dt = np.dtype(dtype)
        return np.stack([tf.uniform(k, (self.dim,), dt) * dt.type(2) - dt.type(1)
                         for k in tf.split(key, n)])


class LogGaussianCoxPines(Target):
    """distributions.py:231-314 (unwhitened), cox_process_utils.py:58-165.

    prior: -0.5 |L^-1 (x-mu)|^2 - 0.5 d log(2pi) - sum log diag L, L = chol(K),
    K_ij = 1.91 exp(-|p_i - p_j|_2 / (40/33)); lik: sum(x c - a exp(x)), a = 1/d."""

    def __init__(self, dim=1600, counts=None):
        import scipy.linalg as sl
        self.dim = dim
        n = int(np.sqrt(dim))
        if counts is None:
            counts = np.loadtxt(_DATA)
        self.counts = np.asarray(counts, np.float64).reshape(dim)
        self.a = 1.0 / dim
        self.signal_variance, self.beta_ls = 1.91, 1.0 / 33
        g = np.arange(n)
        pts = np.array([(i, j) for i in g for j in g], np.float64)      # itertools.product order
        dist = np.sqrt(((pts[:, None, :] - pts[None]) ** 2).sum(-1))
        self.K = self.signal_variance * np.exp(-dist / (n * self.beta_ls))
        self.L = np.linalg.cholesky(self.K)
        self.half_log_det = np.log(np.abs(np.diag(self.L))).sum()
        self.log_norm = -0.5 * dim * np.log(2 * np.pi) - self.half_log_det
        self.mu = np.log(126.0) - 0.5 * self.signal_variance
        self._sl = sl
        self._Kinv = None

    @property
    def Kinv(self):
        if self._Kinv is None:
            Linv = self._sl.solve_triangular(self.L, np.eye(self.dim), lower=True)
            self._Kinv = Linv.T @ Linv
        return self._Kinv

    def loglik(self, x):
        dt = x.dtype
        return (x * self.counts.astype(dt, copy=False) - dt.type(self.a) * np.exp(x)).sum(1)

    def grad_loglik(self, x):
        dt = x.dtype
        return self.counts.astype(dt, copy=False)[None] - dt.type(self.a) * np.exp(x)

    def hvp_loglik(self, x, z):
        return -x.dtype.type(self.a) * np.exp(x) * z

    def hdiag_loglik(self, x):
        return -x.dtype.type(self.a) * np.exp(x)

    def logprior(self, x):
        dt = x.dtype
        # as coded: triangular solve against the Cholesky factor (cox_process_utils.py:162)
        white = self._sl.solve_triangular(self.L.astype(dt, copy=False), (x - dt.type(self.mu)).T, lower=True).T
        return dt.type(-0.5) * (white * white).sum(1) + dt.type(self.log_norm)

    def grad_logprior(self, x):
        dt = x.dtype
        r = (x - dt.type(self.mu)).T
        w = self._sl.solve_triangular(self.L.astype(dt, copy=False), r, lower=True)
        return -self._sl.solve_triangular(self.L.astype(dt, copy=False), w, lower=True, trans="T").T

    def hvp_logprior(self, x, z):
        dt = x.dtype
        w = self._sl.solve_triangular(self.L.astype(dt, copy=False), z.T, lower=True)
        return -self._sl.solve_triangular(self.L.astype(dt, copy=False), w, lower=True, trans="T").T

    def hdiag_logprior(self, x):
        return np.broadcast_to(-np.diag(self.Kinv).astype(x.dtype), x.shape).copy()

    def init_positions(self, key, n, dtype=np.float32):
        """distributions.py:312-314: mu + L @ normal(k,(d,))."""
        dt = np.dtype(dtype)
        eps = tf.vmap_normal(tf.split(key, n), self.dim, dt)
        return (dt.type(self.mu) + eps @ self.L.astype(dt, copy=False).T).astype(dt)


class LogGaussianCoxPinesWhitened(LogGaussianCoxPines):
    """use_whitened=True (distributions.py:276-297): the state is the white noise e; latents f = L e + mu
    (cox_process_utils.get_latents_from_white, :118-140); loglik = Poisson log-likelihood of f; logprior = -|e|^2 / 2 -
    d log(2 pi) / 2.  initialize_model is the SAME function as for the unwhitened target (mu + L eps, :312-314), as coded."""

    def __init__(self, dim=1600, counts=None):
        super().__init__(dim, counts)
        self.log_norm = -0.5 * dim * np.log(2 * np.pi)

    def _f(self, e):
        dt = e.dtype
        return e @ self.L.astype(dt, copy=False).T + dt.type(self.mu)

    def loglik(self, e):
        return LogGaussianCoxPines.loglik(self, self._f(e))

    def grad_loglik(self, e):
        return LogGaussianCoxPines.grad_loglik(self, self._f(e)) @ self.L.astype(e.dtype, copy=False)

    def hvp_loglik(self, e, z):
        dt = e.dtype
        L = self.L.astype(dt, copy=False)
        return -((dt.type(self.a) * np.exp(self._f(e))) * (z @ L.T)) @ L

    def hdiag_loglik(self, e):
        dt = e.dtype
        L = self.L.astype(dt, copy=False)
        return -(dt.type(self.a) * np.exp(self._f(e))) @ (L * L)

    def logprior(self, e):
        dt = e.dtype
        return dt.type(-0.5) * (e * e).sum(1) + dt.type(self.log_norm)

    def grad_logprior(self, e):
        return -e

    def hvp_logprior(self, e, z):
        return -z

    def hdiag_logprior(self, e):
        return -np.ones_like(e)
