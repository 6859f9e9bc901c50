"""Mirror of the reference's mcmc_utils.py metrics on the device (same names, same argument meaning).

stein_disc (mcmc_utils.py:28-85) and max_mean_disc (:88-111) as they are called after training
(exe_flow_matching.py:463-488).  The reference differentiates an arbitrary `logprob_fn` with jax.grad; here it must be
a device log-density (a `Distribution` or `dist.tempered(beta)`), for the same reason as in MALA - no fallback."""
from __future__ import annotations

import torch

from . import _lib
from .distributions import DeviceLogDensity, Distribution


def _score(logprob_fn, X):
    if isinstance(logprob_fn, Distribution):
        logprob_fn = logprob_fn.tempered(1.0)
    elif getattr(logprob_fn, "__self__", None) is not None and isinstance(logprob_fn.__self__, Distribution):
        logprob_fn = logprob_fn.__self__.tempered(1.0)          # `dist.logprob`, as the reference passes it
    if not isinstance(logprob_fn, DeviceLogDensity):
        raise TypeError("stein_disc needs a device log-density (a mfm_b200 Distribution, its .logprob, or dist.tempered(beta)); "
                        "arbitrary Python callables cannot be differentiated on the GPU and there is no fallback.")
    return logprob_fn.value_and_grad(X)[1]


def stein_disc(X: torch.Tensor, logprob_fn, beta: float = -1 / 2):
    """Stein discrepancy with the inverse multi-quadric kernel (1 + |x-x'|^2)^beta; returns (U-statistic, V-statistic)."""
    lib = _lib.load()
    X = X.contiguous().float()
    T, d = X.shape
    G = _score(logprob_fn, X).contiguous()
    out = torch.empty(2, dtype=torch.float32, device=X.device)
    ws = _lib.workspace(lib.mfm_pairwise_workspace_bytes(T), X.device, "pairwise")
    _lib.check(lib.mfm_stein_disc(_lib.ptr(X), _lib.ptr(G), T, d, float(beta), _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream()))
    return out[0], out[1]


def max_mean_disc(X: torch.Tensor, Y: torch.Tensor):
    """Squared maximum mean discrepancy with the Gaussian kernel exp(-|x-y|^2/2) (sigma2 = 1)."""
    lib = _lib.load()
    X, Y = X.contiguous().float(), Y.contiguous().float()
    m, d = X.shape
    assert Y.shape == (m, d), "the reference normalises both sums with m = X.shape[0]"
    out = torch.empty(1, dtype=torch.float32, device=X.device)
    ws = _lib.workspace(lib.mfm_pairwise_workspace_bytes(m), X.device, "pairwise")
    _lib.check(lib.mfm_max_mean_disc(_lib.ptr(X), _lib.ptr(Y), m, d, _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream()))
    return out[0]
