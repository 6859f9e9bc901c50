"""Targets of the MFM hot path — device-resident mirror of the reference's distributions.py.

Same class names, constructor arguments and methods (logprob / loglik / logprior /
initialize_model / sample_model) as distributions.py:42-97,114-165,231-314, but every method works
on a batch of positions [N, d] (CUDA tensors) and is backed by the C-ABI target descriptor
(include/mfm_b200.h::mfm_target_t).  `dist.tempered(beta)` is the device replacement for the
closure `lambda x: beta * dist.loglik(x) + dist.logprior(x)` (exe_flow_matching.py:301,316).
"""
from __future__ import annotations

import itertools
import os

import numpy as np
import torch

from . import _lib, random as mrandom

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class DeviceLogDensity:
    """Callable device log-density `x -> beta*loglik(x) + logprior(x)` with a CUDA value_and_grad."""

    def __init__(self, dist: "Distribution", beta: float = 1.0):
        self.dist = dist
        self.beta = float(beta)

    def desc(self) -> _lib.TargetDesc:
        return self.dist._desc(self.beta)

    def value_and_grad(self, x: torch.Tensor, want_loglik: bool = False):
        lib = _lib.load()
        x = x.contiguous()
        n, d = x.shape
        assert d == self.dist.dim
        logp = torch.empty(n, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        ll = torch.empty(n, dtype=torch.float32, device=x.device) if want_loglik else None
        desc = self.desc()
        ws = _lib.workspace(lib.mfm_target_workspace_bytes(desc, n), x.device, "target")
        _lib.check(lib.mfm_logdensity_and_grad(desc, n, _lib.ptr(x), _lib.ptr(logp), _lib.ptr(grad), _lib.ptr(ll),
                                               _lib.ptr(ws), ws.numel(), _lib.stream()))
        return (logp, grad, ll) if want_loglik else (logp, grad)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        single = x.dim() == 1
        v, _ = self.value_and_grad(x[None] if single else x)
        return v[0] if single else v


class Distribution:
    dim: int
    _kind: int

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.log_Z = 0.0
        self.n_plots = 0
        self.can_sample = False

    # -- descriptor --------------------------------------------------------------------------
    def _fill(self, d: _lib.TargetDesc):
        raise NotImplementedError

    def _desc(self, beta: float) -> _lib.TargetDesc:
        d = _lib.TargetDesc()
        d.kind, d.dim, d.beta = self._kind, self.dim, float(beta)
        self._fill(d)
        return d

    def tempered(self, beta: float = 1.0) -> DeviceLogDensity:
        return DeviceLogDensity(self, beta)

    # -- reference method names ----------------------------------------------------------------
    def logprob(self, x):
        return self.tempered(1.0)(x)

    def loglik(self, x):
        single = x.dim() == 1
        _, _, ll = self.tempered(1.0).value_and_grad(x[None] if single else x, want_loglik=True)
        return ll[0] if single else ll

    def logprior(self, x):
        return self.logprob(x) - self.loglik(x)

    def log_prob(self, x):
        return self.logprob(x)

    def _tensor(self, a):
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)


class GaussianMixture(Distribution):
    """distributions.py:42-77 (dim is fixed to 2 there, :54)."""
    _kind = _lib.TARGET_GMM

    def __init__(self, modes, covs, weights, device=None):
        super().__init__(device)
        modes = np.asarray(modes, np.float32)
        covs = np.asarray(covs, np.float32)
        if covs.ndim == 3:                      # full diagonal matrices as in the default args
            covs = np.stack([np.diag(c) for c in covs])
        self.dim = 2
        assert modes.shape[1] == 2, "GaussianMixture is two-dimensional in the reference"
        self.modes = self._tensor(modes)
        self.covs = self._tensor(covs)
        self.chol_covs = self._tensor(np.sqrt(covs))          # distributions.py:51
        self.weights = self._tensor(np.asarray(weights, np.float32))

    def _fill(self, d):
        d.n_modes = self.modes.shape[0]
        d.modes, d.stds, d.weights = self.modes.data_ptr(), self.chol_covs.data_ptr(), self.weights.data_ptr()

    def initialize_model(self, rng_key, n_chain):
        keys = mrandom.split(rng_key, n_chain)
        self.init_params = mrandom.normal(keys, (self.dim,))


class IndepGaussian(Distribution):
    """distributions.py:80-97."""
    _kind = _lib.TARGET_GAUSS

    def __init__(self, dim, mean=0.0, var=1.0, device=None):
        super().__init__(device)
        self.dim, self.mean, self.std = dim, float(mean), float(np.sqrt(var))

    def _fill(self, d):
        d.gauss_mean, d.gauss_std = self.mean, self.std

    def initialize_model(self, rng_key, n_chain):
        self.init_params = mrandom.normal(mrandom.split(rng_key, n_chain), (self.dim,))

    def sample_model(self, rng_key):
        """rng_key uint32[2] -> [d]; uint32[N,2] -> [N,d] (the vmapped form, exe_flow_matching.py:155)."""
        return self.mean + self.std * mrandom.normal(rng_key, (self.dim,))


class PhiFour(Distribution):
    """distributions.py:114-165 (Dirichlet(0) boundary, no tilt)."""
    _kind = _lib.TARGET_PHI4

    def __init__(self, dim, a=0.1, beta=20.0, bc=("dirichlet", 0), tilt=None, device=None):
        super().__init__(device)
        if bc[0] != "dirichlet" or bc[1] != 0 or tilt is not None:
            raise NotImplementedError("only the configuration the reference runs: Dirichlet(0), no tilt")
        self.dim, self.a, self.beta, self.bc, self.tilt = dim, a, beta, bc, tilt

    def _fill(self, d):
        d.phi_a, d.phi_beta = self.a, self.beta

    def initialize_model(self, rng_key, n_chain):
        keys = mrandom.split(rng_key, n_chain)
        self.init_params = mrandom.uniform(keys, (self.dim,)) * 2 - 1     # distributions.py:162-164


def get_bin_counts(points: np.ndarray, n_bins: int) -> np.ndarray:
    """Histogram of [0,1]^2 points on an n_bins x n_bins grid; points on the upper edge fall in
    the last bin (same rule as cox_process_utils.py:28-55)."""
    idx = np.minimum(np.floor(np.asarray(points, np.float64) * n_bins).astype(np.int64), n_bins - 1)
    counts = np.zeros((n_bins, n_bins))
    np.add.at(counts, (idx[:, 0], idx[:, 1]), 1.0)
    return counts


class LogGaussianCoxPines(Distribution):
    """distributions.py:231-314.  Unwhitened parameterisation (what multi_modal.py constructs): the dense Gaussian prior is
    applied as x K^-1 (one GEMM) instead of two triangular solves against chol(K); K^-1 is formed once in float64 on the host
    (cond(K) = 27.6).  use_whitened=True (:276-297): the state is the white noise e, latents f = L e + mu - value and gradient
    are two GEMMs against the Cholesky factor, the field's Hessian terms two more."""
    _kind = _lib.TARGET_PINES

    def __init__(self, dim, file_path=None, use_whitened=False, device=None):
        super().__init__(device)
        self.use_whitened = bool(use_whitened)
        if self.use_whitened:
            self._kind = _lib.TARGET_PINES_WHITE
        self.dim = dim
        n = int(np.sqrt(dim))
        self._num_grid_per_dim = n
        if file_path is not None:
            counts = get_bin_counts(np.genfromtxt(file_path, delimiter=","), n)
        else:
            if n != 40:
                raise ValueError("bundled bin counts are for the 40x40 grid; pass file_path=finpines.csv")
            counts = np.loadtxt(os.path.join(_DATA, "pines_counts_40x40.txt"))
        self._poisson_a = 1.0 / dim
        self._signal_variance, self._beta = 1.91, 1.0 / 33
        pts = np.array(list(itertools.product(range(n), range(n))), np.float64)
        dist = np.sqrt(((pts[:, None, :] - pts[None]) ** 2).sum(-1))
        K = self._signal_variance * np.exp(-dist / (n * self._beta))
        L = np.linalg.cholesky(K)
        self._half_log_det = float(np.log(np.abs(np.diag(L))).sum())
        self._log_norm = float(-0.5 * dim * np.log(2 * np.pi) - self._half_log_det)
        self._mu_zero = float(np.log(126.0) - 0.5 * self._signal_variance)
        Linv = np.linalg.solve(L, np.eye(dim))
        Kinv = Linv.T @ Linv
        Kinv = 0.5 * (Kinv + Kinv.T)
        self._flat_bin_counts = self._tensor(counts.reshape(dim))
        self._kinv = self._tensor(Kinv)
        self._kinv_mu = self._tensor(self._mu_zero * Kinv.sum(0))
        self._kinv_diag = self._tensor(np.diag(Kinv))
        self._cholesky_gram = self._tensor(L)
        if self.use_whitened:
            self._log_norm = float(-0.5 * dim * np.log(2 * np.pi))              # _white_gaussian_log_normalizer (:267)
            self._chol_t = self._tensor(L.T)
            self._chol_sq_t = self._tensor((L * L).T)
            self._mu_vec = self._tensor(np.full(dim, self._mu_zero))
        # K^-1 is the constant B operand of every pines GEMM: its scaled-fp16 split (include/mfm_b200.h, mfm_gemm_presplit) is
        # made once here instead of in shared memory by every CTA of every call
        self._kinv_split = None
        if self.device.type == "cuda" and dim % 16 == 0:
            lib = _lib.load()
            if lib.mfm_gemm_h16_enabled():
                m = torch.empty(dim * dim + 16, dtype=torch.float32, device=self.device)
                with torch.cuda.device(self.device):
                    _lib.check(lib.mfm_gemm_presplit(self._kinv.data_ptr(), m.data_ptr(), dim * dim, _lib.stream()))
                self._kinv_split = m

    def _fill(self, d):
        d.counts, d.kinv = self._flat_bin_counts.data_ptr(), self._kinv.data_ptr()
        d.kinv_mu, d.kinv_diag = self._kinv_mu.data_ptr(), self._kinv_diag.data_ptr()
        d.kinv_split = self._kinv_split.data_ptr() if self._kinv_split is not None else None
        if self.use_whitened:
            d.chol, d.chol_t = self._cholesky_gram.data_ptr(), self._chol_t.data_ptr()
            d.chol_sq_t, d.mu_vec = self._chol_sq_t.data_ptr(), self._mu_vec.data_ptr()
        d.mu, d.log_norm, d.poisson_a = self._mu_zero, self._log_norm, self._poisson_a

    def initialize_model(self, rng_key, n_chain):
        eps = mrandom.normal(mrandom.split(rng_key, n_chain), (self.dim,))
        # initialisation only: mu + L @ eps (distributions.py:312-314)
        self.init_params = (self._mu_zero + eps @ self._cholesky_gram.T).contiguous()
