"""VectorFieldNet / divergence / adaptive Dopri5 / flow-MH vs the float64 oracle.

Tolerances: field values and divergences 1e-4 relative (north_star); ODE outputs 5e-4 relative
(the solver itself runs at rtol=atol=1e-5 and CUDA computes in float32, so step sequences may
differ by a step); accept decisions identical outside a band around the threshold."""
import zlib
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import samplers as OS, targets as OT, threefry as tf, vector_field as VF
from tests.helpers import key_dev, make_targets, rel_err, to_dev

pytestmark = pytest.mark.gpu

CFG = {  # name: (hidden, hutch, n_times, grad_clip, n chains)
    "4-mode": (128, False, 5, None, 24),
    "gmm16": (128, True, 2, None, 24),
    "phi-four": (128, False, 2, None, 12),
    "pines": (1024, True, 2, 1.0, 6),
}


# "trained-like" heads.  phi-four needs a tame head: a random-sign nn_t head turns nn_t*grad logprob
# into gradient *descent* on a quartic potential, which blows up in finite time (oracle too).
HEAD_SCALE = {"4-mode": 0.2, "gmm16": 0.2, "phi-four": 0.002, "pines": 0.1}


@pytest.fixture(scope="module")
def setups(cuda, lib):
    from mfm_b200 import exe_flow_matching as E
    out = {}
    for name, ot, dd in make_targets(cuda):
        H, hutch, n_times, clip, n = CFG[name]
        rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)      # (hash() of a str is salted per process)
        params = VF.init_params(rng, ot.dim, H, 128, head_scale=HEAD_SCALE[name])
        if name == "phi-four":      # keep nn_t > 0 (gradient ascent on log pi: contracting, no blow-up)
            params["params"]["Dense_4"]["bias"] = (np.abs(params["params"]["Dense_4"]["bias"]) * 0.2).astype(np.float32)
        omega = rng.standard_normal(128).astype(np.float32)
        model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", clip)
        P = E.VectorFieldParams(ot.dim, H, 128, cuda).load_dict(params)
        out[name] = SimpleNamespace(ot=ot, dd=dd, params=params, omega=omega, model=model, P=P, hutch=hutch,
                                    n_times=n_times, clip=clip, n=n)
    return out


def _positions(s, n, seed=1):
    x = s.ot.init_positions(tf.PRNGKey(seed), n, np.float32)
    return x.astype(np.float64)


@pytest.mark.parametrize("name", list(CFG))
def test_field_and_divergence(cuda, setups, name):
    s = setups[name]
    n = s.n
    x = _positions(s, n)
    t = np.linspace(0.0, 1.3, n)
    z = np.random.default_rng(5).standard_normal(x.shape) if s.hutch else None
    v_ref, div_ref = VF.field_and_div(s.params, s.omega, x, t, s.ot, z, s.clip)
    v, div = s.model.apply(s.P, to_dev(x, cuda), to_dev(t, cuda), to_dev(z, cuda) if s.hutch else None,
                           hutch=s.hutch, want_div=True)
    assert rel_err(v.cpu().numpy(), v_ref) < 1e-4, name
    assert np.abs(div.cpu().numpy() - div_ref).max() < 1e-4 * max(np.abs(div_ref).max(), 1.0), name


@pytest.mark.parametrize("name,n", [("4-mode", 333), ("gmm16", 333), ("gmm16", 512), ("phi-four", 300)])
def test_field_and_divergence_tensor_core_rows(cuda, lib, setups, name, n):
    """>= 256 rows: the dense layers run on the tcgen05 kernel and hand their outputs over pre-split (d = 2: max |x| is not
    tracked, Dense_2 writes no copy and Dense_3 must split h2 itself; d = 16: Dense_2 runs on mma.sync and writes the copy from
    its scalar epilogue; 333 / 300 rows: a partial last tile)."""
    s = setups[name]
    lib.mfm_debug_set_h16_min_hidden(0)          # (H = 128 here: by default narrow networks stay off the scaled-fp16 path)
    try:
        _field_rows_check(cuda, s, name, n)
    finally:
        lib.mfm_debug_set_h16_min_hidden(-1)


def _field_rows_check(cuda, s, name, n):
    x = _positions(s, n, seed=3)
    t = np.linspace(0.0, 1.0, n)
    z = np.random.default_rng(6).standard_normal(x.shape) if s.hutch else None
    v_ref, div_ref = VF.field_and_div(s.params, s.omega, x, t, s.ot, z, s.clip)
    v, div = s.model.apply(s.P, to_dev(x, cuda), to_dev(t, cuda), to_dev(z, cuda) if s.hutch else None, hutch=s.hutch, want_div=True)
    assert rel_err(v.cpu().numpy(), v_ref) < 1e-4, name
    assert np.abs(div.cpu().numpy() - div_ref).max() < 1e-4 * max(np.abs(div_ref).max(), 1.0), name


@pytest.mark.parametrize("name", list(CFG))
def test_roundtrip_param_dict(cuda, setups, name):
    s = setups[name]
    back = s.P.to_dict()
    for i in range(8):
        assert np.array_equal(back["params"][f"Dense_{i}"]["kernel"], s.params["params"][f"Dense_{i}"]["kernel"])


def _flow(s):
    ts = np.linspace(0.0, 1.0, s.n_times)
    return OS.Flow(s.params, s.omega, s.ot, s.hutch, 1e-5, 1e-5, 1000, s.clip, ts, rng_dtype=np.float32)


def _tol(got, ref64, ref32, floor, scale=None):
    """An adaptive solve of a ReLU (piecewise-smooth) field is only reproducible to its global
    error: accept/reject sequences differ between float32 and float64 arithmetic.  CUDA (float32)
    must be as close to the float64 oracle as the float32 oracle is (x10: the two discrepancies are
    independent draws from a heavy-tailed distribution), with a floor."""
    scale = max(np.abs(ref64).max(), 1.0) if scale is None else scale
    err = np.abs(np.asarray(got, np.float64) - ref64).max() / scale
    ref = np.abs(ref32.astype(np.float64) - ref64).max() / scale
    return err, max(floor, 10.0 * ref)


_MEASURED = {}


def _record(name, nis, key, value):
    """measured parity numbers, written to gpurun_out/r02_flow_mh_parity.json (quoted in DESIGN.md)"""
    import json, os
    _MEASURED.setdefault(f"{name}/{'indep' if nis < 0 else 'rw'}", {})[key] = value
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "r02_flow_mh_parity.json"), "w") as fh:
            json.dump(_MEASURED, fh, indent=1, sort_keys=True)
    except OSError:
        pass


def _gn(s, mcmc=10, nis=0, step=0.1):
    from mfm_b200 import exe_flow_matching as E
    args = SimpleNamespace(hutchs=s.hutch, num_importance_samples=nis, mcmc_per_flow_steps=mcmc, step_size=step)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=s.n_times)
    return E.create_train_data_gn(s.dd, s.model, opts, args)


@pytest.mark.parametrize("name", list(CFG))
def test_push_pull_vs_oracle(cuda, setups, name):
    s = setups[name]
    n = s.n
    gen, init_fn, transform_and_logdet = _gn(s)
    flow = _flow(s)
    keys = tf.split(tf.PRNGKey(77), n)
    u = tf.vmap_normal(tf.split(tf.PRNGKey(4), n), s.ot.dim).astype(np.float64)
    st_o = {}
    x_ref, ldj_ref = flow.transform_and_logdet(keys, u, st_o)
    x32, ldj32 = flow.transform_and_logdet(keys, u.astype(np.float32))
    stats = torch.zeros(8, dtype=torch.int32, device=cuda)
    x, ldj = transform_and_logdet(key_dev(keys, cuda), to_dev(u, cuda), s.P, stats)
    # log-det: a d-term cancelling sum evaluated in float32 carries ~1e-6*d absolute round-off
    ldj_floor = 5e-3 + 1.2e-5 * s.ot.dim      # (pines: 2.4e-2; the float32 oracle itself is off by up to 9e-2 on 96 chains, see the flip-rate test)
    err, tol = _tol(x.cpu().numpy(), x_ref, x32, 1e-3)
    assert err < tol, (name, err, tol)
    err, tol = _tol(ldj.cpu().numpy(), ldj_ref, ldj32, ldj_floor)
    assert err < tol, (name, err, tol)
    acc, tried, mx, nev = stats.cpu().tolist()[:4]
    chain_evals = int(stats[4:6].cpu().view(torch.int64).item())
    assert 2 * n <= chain_evals <= n * nev
    assert abs(tried - int(st_o["n_try"].sum())) <= max(3, 0.15 * tried), (tried, st_o["n_try"].sum())
    assert nev == 2 + 6 * mx
    # inverse direction + round trip
    u_ref, v0_ref = flow.inverse_and_logdet(keys, x_ref)
    u32, v032 = flow.inverse_and_logdet(keys, x_ref.astype(np.float32))
    ub, v0 = gen.inverse_and_logdet(key_dev(keys, cuda), to_dev(x_ref, cuda), s.P)
    err, tol = _tol(ub.cpu().numpy(), u_ref, u32, 1e-3)
    assert err < tol, (name, err, tol)
    err, tol = _tol(v0.cpu().numpy(), v0_ref, v032, ldj_floor)
    assert err < tol, (name, err, tol)
    # exact divergence on a tame field: pull(push(u)) == u and the log-dets cancel (4-mode's score
    # switches sharply between modes, so its round trip amplifies solver error and is not asserted)
    if name == "phi-four":
        assert rel_err(ub.cpu().numpy(), u) < 1e-2
        assert np.abs(v0.cpu().numpy() + ldj_ref).max() < 2e-2 * max(np.abs(ldj_ref).max(), 1.0)


@pytest.mark.parametrize("name", ["4-mode", "gmm16"])
@pytest.mark.parametrize("n", [24, 333])
def test_fused_small_solve_matches_the_general_path(cuda, lib, setups, name, n):
    """The one-launch solve of the small shapes (csrc/ode_small.cuh: 16 chains per CTA, whole Dormand-Prince loop on chip) against
    the multi-kernel lock-step driver it replaces: same algorithm, different summation order inside the dense layers, so the
    two agree to the solver's own reproducibility; the step statistics (per-chain counts) must be of the same size."""
    s = setups[name]
    gen, init_fn, push = _gn(s)
    keys = key_dev(tf.split(tf.PRNGKey(7), n), cuda)
    u = to_dev(tf.vmap_normal(tf.split(tf.PRNGKey(8), n), s.ot.dim), cuda)
    out = {}
    try:
        for mode in (1, 0):
            lib.mfm_debug_set_ode_small(mode)
            stats = torch.zeros(8, dtype=torch.int32, device=cuda)
            x, ldj = push(keys, u, s.P, stats)
            ub, v0 = gen.inverse_and_logdet(keys, x, s.P)
            out[mode] = (x.cpu().numpy(), ldj.cpu().numpy(), ub.cpu().numpy(), v0.cpu().numpy(), stats.cpu().tolist())
    finally:
        lib.mfm_debug_set_ode_small(1)
    f, g = out[1], out[0]
    scale = max(np.abs(g[0]).max(), 1.0)
    assert np.abs(f[0] - g[0]).max() <= 2e-3 * scale
    # the pull starts from each path's own x: 4-mode's score switches sharply between modes and the way back amplifies the
    # difference of the starting points (max over 333 chains 8e-3, median 5e-4)
    assert np.abs(f[2] - g[2]).max() <= 5e-3 * max(np.abs(g[2]).max(), 1.0) and np.median(np.abs(f[2] - g[2])) <= 1e-3
    assert np.abs(f[1] - g[1]).max() <= 2e-2 * max(np.abs(g[1]).max(), 1.0)
    acc_f, try_f, mx_f, nev_f = f[4][:4]; acc_g, try_g, mx_g, nev_g = g[4][:4]
    assert nev_f == 2 + 6 * mx_f and nev_g == 2 + 6 * mx_g
    assert abs(try_f - try_g) <= max(3, 0.1 * try_g) and abs(acc_f - acc_g) <= max(3, 0.1 * acc_g)
    ce_f = int(np.array(f[4][4:6], dtype=np.int32).view(np.int64)[0])
    assert ce_f == 2 * n + 6 * try_f                       # chain-evaluations: every chain's own attempts, no lock-step padding


def test_identity_flow_zero_heads(cuda, setups):
    """Reference initialisation (zero nn_t / nn_xt kernels, :81,86): v == bias, 7 growing steps."""
    from mfm_b200 import exe_flow_matching as E
    s = setups["phi-four"]
    P = E.VectorFieldParams(64, 128, 128, cuda).load_dict(s.params)
    P.kernel(4).zero_(); P.kernel(7).zero_(); P.bias(4).zero_(); P.bias(7).zero_()
    gen, _, push = _gn(s)
    u = to_dev(np.random.default_rng(0).standard_normal((8, 64)), cuda)
    stats = torch.zeros(8, dtype=torch.int32, device=cuda)
    x, ldj = push(key_dev(tf.split(tf.PRNGKey(1), 8), cuda), u, P, stats)
    assert torch.allclose(x, u, atol=1e-6) and torch.all(ldj == 0)
    assert stats.cpu().tolist()[2] == 7       # dt: 1e-6 * 10^k until t >= 1


@pytest.mark.parametrize("name,nis", [("4-mode", 0), ("gmm16", 0), ("phi-four", 0), ("pines", 0), ("gmm16", -1)])
def test_flow_mh_step(cuda, setups, name, nis):
    from mfm_b200.bblackjax.mcmc import mala as M
    s = setups[name]
    n = s.n
    gen, init_fn, _ = _gn(s, nis=nis)
    flow = _flow(s)
    x0 = _positions(s, n, seed=3)
    beta = 0.7
    st_d = init_fn(to_dev(x0, cuda), beta)
    st_o = OS.mala_init(x0, s.ot, beta)
    key = tf.PRNGKey(2024)
    keys = tf.split(key, n)
    dbg, dbg32 = {}, {}
    st_o32 = OS.MALAState(*[a.astype(np.float32) for a in st_o])
    ref = OT.IndepGaussian(s.ot.dim)
    if nis < 0:
        new_o, info_o = OS.indep_flow_mh_step(keys, st_o, s.ot, flow, ref, beta, dbg)
        _, info_32 = OS.indep_flow_mh_step(keys, st_o32, s.ot, flow, ref, beta, dbg32)
    else:
        new_o, info_o = OS.rw_flow_mh_step(keys, st_o, s.ot, flow, beta, dbg)
        _, info_32 = OS.rw_flow_mh_step(keys, st_o32, s.ot, flow, beta, dbg32)
    new_d, info_d = gen.flow_step(key_dev(key, cuda), st_d, s.dd.tempered(beta), s.P)
    la = dbg["log_acc"]
    la32 = dbg32["log_acc"].astype(np.float64)
    err, tol = _tol(info_d.proposed_position.cpu().numpy(), info_o.proposed_position, info_32.proposed_position, 2e-3)
    assert err < tol, (name, err, tol)
    # log acceptance ratio l' - V' - l - V0 (:271-274).  Its float32 evaluation is exact to a few ulps of the largest
    # term (|l| ~ 2e3 for pines / phi-four: 4 ulp = 1e-3) and the two log-dets carry the adaptive solve's own float32
    # sensitivity, which the float32 ORACLE measures (|la32 - la64|; its ensemble maximum, because which chain draws the
    # large deviation is a matter of where an accept / reject sequence happens to split): the device must be that good (x4).
    # No term proportional to |l|: a band of +-10 in log alpha would accept any decision.
    l_mag = np.maximum(np.abs(st_o.logdensity), np.abs(dbg["lp"]) if "lp" in dbg else 0.0)
    ulp = np.spacing(np.maximum(l_mag, 1.0).astype(np.float32)).astype(np.float64)
    fin32 = np.isfinite(la) & np.isfinite(la32) & (np.abs(la) < 80)
    la_tol = np.maximum(1e-3, 4.0 * ulp) + 4.0 * (np.abs(la32 - la)[fin32].max() if fin32.any() else 0.0)
    with np.errstate(over="ignore", divide="ignore"):
        la_d = np.log(info_d.acceptance_rate.cpu().numpy().astype(np.float64))
        logu = np.log(dbg["u"])
    fin = np.isfinite(la) & np.isfinite(la_d) & (np.abs(la) < 80)
    la_err = np.abs(la_d[fin] - la[fin])
    _record(name, nis, "log_alpha_abs_err_max", float(la_err.max()) if la_err.size else 0.0)
    _record(name, nis, "log_alpha_abs_err_f32_oracle_max", float(np.abs(la32 - la)[fin].max()) if la_err.size else 0.0)
    if name in ("phi-four", "pines"):
        # 6-12 chains of a heavy-tailed error (one chain whose accept / reject sequence splits early dominates the maximum; the
        # float32 oracle shows the same tail, 0.05 - 0.5 from run to run of the same algorithm): the median carries the
        # bound here, the distribution is checked on 96 / 256 chains by test_flow_mh_decision_flip_rate
        e32 = np.abs(la32 - la)[fin32]
        assert np.median(la_err) <= 3.0 * np.median(e32) + 5e-3, (name, np.median(la_err), np.median(e32))
        assert la_err.max() <= 10.0 * e32.max() + 5e-2, (name, la_err.max(), e32.max())
        la_tol = np.maximum(la_tol, np.abs(la_d - la) + 1e-6)
    else:
        assert (la_err < la_tol[fin]).all(), (name, la_err.max(), la_tol[fin][np.argmax(la_err - la_tol[fin])])
    acc_d = info_d.is_accepted.cpu().numpy(); acc_o = info_o.is_accepted
    band = np.abs(la - logu) < la_tol
    assert ((acc_d == acc_o) | band).all(), name
    _record(name, nis, "decision_flips", int((acc_d != acc_o).sum())); _record(name, nis, "chains", int(n))
    same = acc_d == acc_o
    err, tol = _tol(new_d.position.cpu().numpy()[same], new_o.position[same], new_o.position[same].astype(np.float32), tol)
    assert err < tol
    assert (info_d.proposed_weight == 0).all()


@pytest.mark.parametrize("name,K", [("gmm16", 4), ("4-mode", 3), ("gmm16", 7)])   # (as coded the weights are exp(log-density): they underflow to 0 / 0 for phi-four and pines, |l| ~ 2e3, in float64 too)
def test_cis_flow_step(cuda, setups, name, K):
    """conditional_importance_sampling (exe_flow_matching.py:280-296, --num_importance_samples K) against the oracle: K + 1
    normalised weights per chain, the index drawn by jax.random.choice from one uniform, the state update (the gradient of the
    previous state is kept, as coded) and the info fields."""
    s = setups[name]
    n = s.n
    gen, init_fn, _ = _gn(s, nis=K)
    flow = _flow(s)
    x0 = _positions(s, n, seed=5)
    beta = 0.8
    st_d = init_fn(to_dev(x0, cuda), beta)
    st_o = OS.mala_init(x0, s.ot, beta)
    key = tf.PRNGKey(77)
    new_o, info_o = OS.cis_flow_step(tf.split(key, n), st_o, s.ot, flow, OT.IndepGaussian(s.ot.dim), K, beta)
    new_d, info_d = gen.flow_step(key_dev(key, cuda), st_d, s.dd.tempered(beta), s.P)
    acc_d, acc_o = info_d.is_accepted.cpu().numpy(), info_o.is_accepted
    w_d, w_o = info_d.acceptance_rate.cpu().numpy().astype(np.float64), info_o.acceptance_rate
    same = (acc_d == acc_o) & (np.abs(w_d - w_o) <= 2e-2 * np.maximum(w_o, 1e-6))      # the same candidate was drawn
    # a draw lands on the other side of a cumulative-weight boundary only when the ODE-level differences in the weights
    # (~1e-2 relative, see test_flow_mh_decision_flip_rate) straddle r: at most one chain of the ensemble here
    assert same.sum() >= n - 1, (name, same.sum(), n)
    assert np.allclose(w_d[same], w_o[same], rtol=2e-2)
    assert torch.equal(info_d.proposed_weight, info_d.acceptance_rate)
    err, tol = _tol(new_d.position.cpu().numpy()[same], new_o.position[same], new_o.position[same].astype(np.float32), 2e-3)
    assert err < tol, (name, err, tol)
    err, tol = _tol(info_d.proposed_position.cpu().numpy()[same], info_o.proposed_position[same], info_o.proposed_position[same].astype(np.float32), 2e-3)
    assert err < tol, (name, err, tol)
    ld, lo = new_d.logdensity.cpu().numpy()[same], new_o.logdensity[same]
    assert np.abs(ld - lo).max() <= 2e-3 * max(np.abs(lo).max(), 1.0)
    assert torch.equal(new_d.logdensity_grad, st_d.logdensity_grad)                     # as coded: not recomputed


def test_train_data_generator_dispatch(cuda, setups):
    """count % (m+1) == 0 -> flow step, else MALA (exe_flow_matching.py:311-313)."""
    s = setups["gmm16"]
    gen, init_fn, _ = _gn(s, mcmc=2, step=0.2)
    st = init_fn(to_dev(_positions(s, 16), cuda), 1.0)
    key = key_dev(tf.PRNGKey(5), cuda)
    _, info = gen(key, st, 1, s.P, 1.0)
    assert (info.proposed_weight != 0).any()           # MALA info carries exp(...) weights
    _, info = gen(key, st, 3, s.P, 1.0)
    assert (info.proposed_weight == 0).all()           # flow-MH info has weight 0


@pytest.mark.parametrize("name,n", [("phi-four", 256), ("pines", 96)])
def test_flow_mh_decision_flip_rate(cuda, setups, name, n):
    """Accept decisions of the flow-MH step against the float64 oracle on a larger ensemble, for the two targets whose
    log-densities are ~2e3 (where a relative band would be vacuous): the absolute log alpha error and the number of flipped
    decisions are MEASURED (gpurun_out/r02_flow_mh_parity.json, quoted in DESIGN.md) and bounded.

    What bounds log alpha is the SOLVER, not the arithmetic: odeint controls the local error of the augmented state to
    rtol * |y| per step with rtol = 1e-5, and the log-det component is O(1e2 - 1e3) here, so two correct executions of the same
    algorithm whose accept / reject sequences differ (float32 vs float64 rounding is enough) legitimately differ by
    ~1e-2 per step in the log-det.  The float32 ORACLE (exact IEEE float32 matmuls, no tensor cores) run on the same chains
    measures that: the device must be as close to the float64 oracle as the float32 oracle is (x4 on the maximum, x3 on the
    median), and a decision may only flip when the oracle's |log alpha - log u| is inside the device's error."""
    s = setups[name]
    gen, init_fn, _ = _gn(s, nis=0)
    flow = _flow(s)
    x0 = _positions(s, n, seed=11)
    beta = 1.0
    st_d = init_fn(to_dev(x0, cuda), beta)
    st_o = OS.mala_init(x0, s.ot, beta)
    st_o32 = OS.MALAState(*[a.astype(np.float32) for a in st_o])
    key = tf.PRNGKey(4242)
    dbg, dbg32 = {}, {}
    new_o, info_o = OS.rw_flow_mh_step(tf.split(key, n), st_o, s.ot, flow, beta, dbg)
    OS.rw_flow_mh_step(tf.split(key, n), st_o32, s.ot, flow, beta, dbg32)
    new_d, info_d = gen.flow_step(key_dev(key, cuda), st_d, s.dd.tempered(beta), s.P)
    la, la32 = dbg["log_acc"], dbg32["log_acc"].astype(np.float64)
    with np.errstate(over="ignore", divide="ignore"):
        la_d = np.log(info_d.acceptance_rate.cpu().numpy().astype(np.float64))
        logu = np.log(dbg["u"])
    fin = np.isfinite(la) & np.isfinite(la_d) & np.isfinite(la32) & (np.abs(la) < 80)
    err, err32 = np.abs(la_d - la)[fin], np.abs(la32 - la)[fin]
    acc_d = info_d.is_accepted.cpu().numpy(); acc_o = info_o.is_accepted
    flips = int((acc_d != acc_o).sum())
    flips32 = int(((dbg32["u"] <= np.exp(np.minimum(dbg32["log_acc"], 80.0))) != acc_o).sum())
    tag = f"{name}-flip"
    for k, v in [("chains", n), ("log_alpha_abs_err_max", float(err.max())), ("log_alpha_abs_err_median", float(np.median(err))),
                 ("f32_oracle_log_alpha_abs_err_max", float(err32.max())), ("f32_oracle_log_alpha_abs_err_median", float(np.median(err32))),
                 ("decision_flips", flips), ("flip_rate", flips / n), ("f32_oracle_decision_flips", flips32),
                 ("oracle_accept_rate", float(acc_o.mean())), ("log_det_abs_max", float(max(np.abs(dbg["V0"]).max(), np.abs(dbg["Vp"]).max())))]:
        _record(tag, 0, k, v)
    assert err.max() <= 4.0 * err32.max() + 1e-3, (name, err.max(), err32.max())
    assert np.median(err) <= 3.0 * np.median(err32) + 1e-3, (name, np.median(err), np.median(err32))
    # every flipped decision sits inside the device's own error band around the threshold, and the rate stays below 1 %
    band = np.abs(la - logu) <= np.abs(la_d - la) + 1e-6
    assert ((acc_d == acc_o) | band | ~fin).all(), name
    assert flips <= max(1, n // 100), (name, flips)
