"""Oracle: evaluation metrics.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/mcmc_utils.py:28-85 (stein_disc, IMQ kernel, U- and V-statistics) and :88-111 (max_mean_disc,
Gaussian kernel).  Pure NumPy over the full T x T pair matrix (float64): use at T <= a few thousand."""
from __future__ import annotations

import numpy as np


def stein_disc(X, grad_logprob, beta=-1 / 2):
    X = np.asarray(X, np.float64)
    T, d = X.shape
    G = np.asarray(grad_logprob(X), np.float64)
    b = -beta                                                        # :49
    diff = X[:, None, :] - X[None, :, :]                             # disc(x, x_): diff = x - x_
    dot = (diff * diff).sum(-1)
    gd = ((G[:, None, :] - G[None, :, :]) * diff).sum(-1)
    gg = G @ G.T
    disc = (-4 * b * (b + 1) * dot / (1 + dot) ** (b + 2) + 2 * b * (d + gd) / (1 + dot) ** (1 + b) + gg / (1 + dot) ** b)   # :69-73
    mc = disc.sum()
    return (mc - np.trace(disc)) / (T * (T - 1)), mc / T ** 2        # :85


def max_mean_disc(X, Y):
    X, Y = np.asarray(X, np.float64), np.asarray(Y, np.float64)
    m = X.shape[0]

    def ksum(A, B):
        r = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
        return np.exp(-0.5 * r).sum()

    m2 = m * m
    return (ksum(X, X) - m) / (m2 - m) - 2 * ksum(X, Y) / m2 + (ksum(Y, Y) - m) / (m2 - m)   # :104-109
