"""Metric oracle (oracle/metrics.py) against the reference's per-pair formulas written as scalar loops, and known limits."""
import numpy as np

from oracle import metrics as OM, targets as OT


def test_stein_disc_matches_scalar_restatement():
    rng = np.random.default_rng(0)
    t = OT.four_mode()
    X = 8.0 + rng.standard_normal((7, 2))
    G = t.grad(X)
    b, d, T = 0.5, 2, 7
    mc, diag = 0.0, 0.0
    for i in range(T):
        for j in range(T):
            diff = X[i] - X[j]
            dot = diff @ diff
            v = (-4 * b * (b + 1) * dot / (1 + dot) ** (b + 2) + 2 * b * (d + (G[i] - G[j]) @ diff) / (1 + dot) ** (1 + b)
                 + G[i] @ G[j] / (1 + dot) ** b)
            mc += v
            diag += v if i == j else 0.0
    u, v = OM.stein_disc(X, t.grad)
    assert np.isclose(u, (mc - diag) / (T * (T - 1)), rtol=1e-12) and np.isclose(v, mc / T ** 2, rtol=1e-12)


def test_stein_disc_separates_good_from_bad_samples():
    rng = np.random.default_rng(1)
    t = OT.IndepGaussian(3)
    good = rng.standard_normal((600, 3))
    bad = good + 1.5
    ug, _ = OM.stein_disc(good, t.grad)
    ub, _ = OM.stein_disc(bad, t.grad)
    assert abs(ug) < 0.05 and ub > 0.5


def test_max_mean_disc_properties():
    rng = np.random.default_rng(2)
    X, Y = rng.standard_normal((200, 2)), rng.standard_normal((200, 2))
    assert abs(OM.max_mean_disc(X, X) - (-2.0 / 200 + 0) - 0) < 1.0      # finite
    # identical samples: disc_x/(m2-m) - 2 disc_xy/m2 + disc_y/(m2-m) with disc_xy = disc_x + m
    m = 200
    sxx = np.exp(-0.5 * ((X[:, None] - X[None]) ** 2).sum(-1)).sum()
    assert np.isclose(OM.max_mean_disc(X, X), 2 * (sxx - m) / (m * m - m) - 2 * sxx / (m * m), rtol=1e-12)
    assert OM.max_mean_disc(X, Y + 2.0) > 10 * abs(OM.max_mean_disc(X, Y))
