"""Oracle ODE solver (oracle/ode.py, restated jax.experimental.ode.odeint): tableau identities, linear fields with a
closed-form solution, per-chain step control, the per-segment mxstep budget."""
import numpy as np
import scipy.linalg

from oracle import ode as OO


def test_dopri5_tableau_identities():
    assert np.allclose(OO.BETA.sum(1), OO.ALPHA, atol=1e-15)
    assert abs(OO.C_SOL.sum() - 1) < 1e-15 and abs(OO.C_ERR.sum()) < 1e-15 and abs(OO.C_MID.sum() - 0.5) < 1e-15
    assert np.allclose(OO.C_SOL[:6], OO.BETA[5], atol=0)             # FSAL: the 7th stage is f(y1)
    # order conditions up to 3 for the 5th-order weights
    c = np.concatenate([[0.0], OO.ALPHA])
    assert abs((OO.C_SOL * c).sum() - 1 / 2) < 1e-15 and abs((OO.C_SOL * c ** 2).sum() - 1 / 3) < 1e-15
    assert abs((OO.C_SOL * c ** 3).sum() - 1 / 4) < 1e-15 and abs((OO.C_SOL * c ** 4).sum() - 1 / 5) < 1e-14


def test_linear_field_matches_expm():
    rng = np.random.default_rng(0)
    D, N = 6, 5
    A = rng.standard_normal((D, D)) * 0.8
    y0 = rng.standard_normal((N, D))
    fun = lambda y, t: y @ A.T
    stats = {}
    y1 = OO.odeint_final(fun, y0, np.array([0.0, 1.0]), 1e-5, 1e-5, 1000, stats)
    ref = y0 @ scipy.linalg.expm(A).T
    assert np.abs(y1 - ref).max() <= 2e-5 * np.abs(ref).max()
    assert stats["n_eval"] == 2 + 6 * stats["n_try"].max() and (stats["n_acc"] <= stats["n_try"]).all()
    # tighter tolerance -> closer
    y1t = OO.odeint_final(fun, y0, np.array([0.0, 1.0]), 1e-9, 1e-9, 1000)
    assert np.abs(y1t - ref).max() <= 1e-8 * np.abs(ref).max()


def test_time_dependent_field_and_intermediate_times():
    # y' = cos(3 t) y  ->  y(1) = y0 exp(sin(3)/3); output times do not change the step sequence
    y0 = np.array([[1.0], [-2.0], [0.5]])
    fun = lambda y, t: np.cos(3 * t)[:, None] * y
    s2, s5 = {}, {}
    a = OO.odeint_final(fun, y0, np.array([0.0, 1.0]), 1e-6, 1e-6, 1000, s2)
    b = OO.odeint_final(fun, y0, np.linspace(0, 1, 5), 1e-6, 1e-6, 1000, s5)
    ref = y0 * np.exp(np.sin(3.0) / 3)
    assert np.abs(a - ref).max() < 1e-5 and np.abs(b - ref).max() < 1e-5
    assert (s5["n_try"] >= s2["n_try"]).all()


def test_every_chain_runs_its_own_controller():
    # decay rates 1 and 200: the stiff chain needs many more steps; the slow chain's result is what it gets alone
    lam = np.array([1.0, 200.0])
    fun = lambda y, t: -lam[: y.shape[0], None] * y
    y0 = np.ones((2, 1))
    st = {}
    both = OO.odeint_final(fun, y0, np.array([0.0, 1.0]), 1e-5, 1e-5, 1000, st)
    alone = OO.odeint_final(lambda y, t: -y, y0[:1], np.array([0.0, 1.0]), 1e-5, 1e-5, 1000)
    assert st["n_try"][1] > 3 * st["n_try"][0]
    assert np.allclose(both[:1], alone, rtol=1e-13, atol=0)
    assert abs(both[0, 0] - np.exp(-1)) < 1e-5


def test_mxstep_budget_is_per_output_segment():
    fun = lambda y, t: -50.0 * y
    y0 = np.ones((1, 1))
    st = {}
    OO.odeint_final(fun, y0, np.array([0.0, 1.0]), 1e-6, 1e-6, 3, st)
    assert st["n_try"][0] == 3
    st5 = {}
    OO.odeint_final(fun, y0, np.linspace(0, 1, 5), 1e-6, 1e-6, 3, st5)
    assert st5["n_try"][0] == 12                       # i resets per target time (4 segments)


def test_float32_solve_tracks_float64():
    rng = np.random.default_rng(1)
    A = rng.standard_normal((4, 4)) * 0.5
    y0 = rng.standard_normal((3, 4))
    f64 = OO.odeint_final(lambda y, t: y @ A.T, y0, np.array([0.0, 1.0]), 1e-5, 1e-5, 1000)
    A32 = A.astype(np.float32)
    f32 = OO.odeint_final(lambda y, t: y @ A32.T, y0.astype(np.float32), np.array([0.0, 1.0]), 1e-5, 1e-5, 1000)
    assert f32.dtype == np.float32 and np.abs(f32 - f64).max() <= 1e-4 * np.abs(f64).max()


def test_agrees_with_scipy_rk45_on_a_nonlinear_field():
    """Independent implementation of the same Dormand-Prince pair (scipy's RK45 uses its own controller and initial step):
    both must land within the requested tolerance of each other on a smooth nonlinear field."""
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(3)
    W = rng.standard_normal((5, 5)) * 0.7
    y0 = rng.standard_normal((4, 5))

    def f1(t, y):
        return np.tanh(W @ y) * (1.0 + 0.5 * np.sin(4 * t)) - 0.3 * y

    ours = OO.odeint_final(lambda y, t: np.tanh(y @ W.T) * (1.0 + 0.5 * np.sin(4 * t))[:, None] - 0.3 * y, y0,
                           np.array([0.0, 1.0]), 1e-7, 1e-7, 1000)
    for n in range(4):
        ref = solve_ivp(f1, (0.0, 1.0), y0[n], method="RK45", rtol=1e-10, atol=1e-10).y[:, -1]
        assert np.abs(ours[n] - ref).max() < 5e-6


def test_erf_inv_restatement_matches_scipy():
    """XLA's float32 ErfInv polynomial (Giles), as jax.random.normal uses it, against scipy's erfinv: the single-precision
    approximation itself is good to ~3.5e-6 relative in the tails and 2e-7 absolute in the centre (bit-exactness against JAX is
    pinned separately by the Sharp-Bits values in test_oracle_rng.py)."""
    import scipy.special
    from oracle import threefry as tf
    x = np.linspace(-0.99999, 0.99999, 20001).astype(np.float32)
    got = tf.erf_inv_f32(x).astype(np.float64)
    ref = scipy.special.erfinv(x.astype(np.float64))
    assert (np.abs(got - ref) <= 5e-6 * np.maximum(np.abs(ref), 0.1)).all()
    assert np.abs(got - ref)[np.abs(x) < 0.9].max() <= 3e-7
