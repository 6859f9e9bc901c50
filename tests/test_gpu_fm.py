"""Flow-matching loss / gradient / AdamW step vs the oracle (loss 1e-4 relative, north_star)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import optim as OO, targets as OT, threefry as tf, vector_field as VF
from tests.helpers import key_dev, make_targets, rel_err, to_dev
from tests.test_gpu_flow import CFG

pytestmark = pytest.mark.gpu


def _args():
    return SimpleNamespace(ref_dist="stdgauss", cond_flow=True, ot_cond_flow=False, sigma=1e-4, adam_beta1=0.9,
                           adam_beta2=0.999, adam_epsilon=1e-8, weight_decay=1e-4, gradient_clip=1.0)


@pytest.fixture(scope="module")
def setups(cuda, lib):
    from mfm_b200 import exe_flow_matching as E
    out = {}
    for name, ot, dd in make_targets(cuda):
        H, hutch, n_times, clip, n = CFG[name]
        rng = np.random.default_rng(11)
        params = VF.init_params(rng, ot.dim, H, 128, head_scale=0.3)
        omega = rng.standard_normal(128).astype(np.float32)
        model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", clip)
        P = E.VectorFieldParams(ot.dim, H, 128, cuda).load_dict(params)
        lr = E.create_learning_rate_fn(1000, 0, 1e-3)
        state = E.create_train_state(model, P, lr, _args())
        out[name] = SimpleNamespace(ot=ot, dd=dd, params=params, omega=omega, model=model, P=P, clip=clip, state=state)
    return out


def _flat_grads(s, G):
    parts = []
    for i, (fi, fo) in enumerate(s.P.shapes):
        parts.append((s.P.w_off[i], G["params"][f"Dense_{i}"]["kernel"].ravel()))
        parts.append((s.P.b_off[i], G["params"][f"Dense_{i}"]["bias"].ravel()))
    out = np.zeros(s.P.n_params)
    for off, a in parts:
        out[off:off + a.size] = a
    return out


@pytest.mark.parametrize("name,n", [("4-mode", 128), ("gmm16", 77), ("phi-four", 300), ("pines", 40)])
def test_fm_loss_and_grad(cuda, setups, name, n):
    s = setups[name]
    x = s.ot.init_positions(tf.PRNGKey(8), n, np.float32).astype(np.float64)
    key = tf.PRNGKey(31337)
    times, xt, target = VF.fm_batch(key, x, OT.IndepGaussian(s.ot.dim).sample, 1e-4, rng_dtype=np.float32)
    loss_ref, G = VF.fm_loss_and_grad(s.params, s.omega, xt, times, target, s.ot.grad, s.clip)
    loss, grads = s.state.loss_and_grad(key_dev(key, cuda), to_dev(x, cuda))
    assert abs(loss.item() - loss_ref) < 1e-4 * abs(loss_ref), (loss.item(), loss_ref)
    g_ref = _flat_grads(s, G)
    g = grads.cpu().numpy()
    assert rel_err(g, g_ref) < 2e-4
    for i in range(8):   # every layer individually (small layers must not hide behind big ones)
        fi, fo = s.P.shapes[i]
        sl = slice(s.P.w_off[i], s.P.w_off[i] + fi * fo)
        assert rel_err(g[sl], g_ref[sl]) < 5e-4, (name, i)
        sl = slice(s.P.b_off[i], s.P.b_off[i] + fo)
        assert rel_err(g[sl], g_ref[sl]) < 5e-4, (name, "bias", i)


@pytest.mark.parametrize("act", ["tanh", "elu", "gelu", "swish"])
@pytest.mark.parametrize("name,n", [("gmm16", 77), ("pines", 300)])
def test_activations(cuda, setups, name, n, act):
    """--non_linearity (exe_flow_matching.py:40-46, multi_modal.py:181): jax.nn.tanh / elu / gelu (tanh approximation) / swish.
    Field value, Hutchinson divergence (forward-mode tangent through the activation derivatives) and the FM loss / gradient
    (reverse mode through the same derivatives) against the float64 oracle; pines at 300 chains goes through the tcgen05
    kernels incl. the pre-split hand-over, gmm16 through the warp-level ones."""
    from mfm_b200 import exe_flow_matching as E
    s = setups[name]
    H = s.P.hidden
    model = E.VectorFieldNet(to_dev(s.omega, cuda), s.dd, [H, H], [H, H], [H, H], act, s.clip)
    state = E.create_train_state(model, s.P, E.create_learning_rate_fn(1000, 0, 1e-3), _args())
    x = s.ot.init_positions(tf.PRNGKey(8), n, np.float32).astype(np.float64)
    t = np.linspace(0.0, 1.2, n)
    z = np.random.default_rng(5).standard_normal(x.shape)
    v_ref, div_ref = VF.field_and_div(s.params, s.omega, x, t, s.ot, z, s.clip, act=act)
    v, div = model.apply(s.P, to_dev(x, cuda), to_dev(t, cuda), to_dev(z, cuda), hutch=True, want_div=True)
    assert rel_err(v.cpu().numpy(), v_ref) < 1e-4, (name, act)
    assert np.abs(div.cpu().numpy() - div_ref).max() < 1e-4 * max(np.abs(div_ref).max(), 1.0), (name, act)
    key = tf.PRNGKey(31337)
    times, xt, target = VF.fm_batch(key, x, OT.IndepGaussian(s.ot.dim).sample, 1e-4, rng_dtype=np.float32)
    loss_ref, G = VF.fm_loss_and_grad(s.params, s.omega, xt, times, target, s.ot.grad, s.clip, act=act)
    loss, grads = state.loss_and_grad(key_dev(key, cuda), to_dev(x, cuda))
    assert abs(loss.item() - loss_ref) < 1e-4 * abs(loss_ref), (loss.item(), loss_ref)
    g_ref, g = _flat_grads(s, G), grads.cpu().numpy()
    assert rel_err(g, g_ref) < 2e-4
    for i in range(8):
        fi, fo = s.P.shapes[i]
        sl = slice(s.P.w_off[i], s.P.w_off[i] + fi * fo)
        assert rel_err(g[sl], g_ref[sl]) < 5e-4, (name, act, i)


def test_activation_exact_divergence(cuda, setups):
    """trace(jacfwd) path (d tangents) with a smooth activation: phi-four's 64 basis tangents through tanh."""
    from mfm_b200 import exe_flow_matching as E
    s = setups["phi-four"]
    H = s.P.hidden
    model = E.VectorFieldNet(to_dev(s.omega, cuda), s.dd, [H, H], [H, H], [H, H], "tanh", s.clip)
    n = 24
    x = s.ot.init_positions(tf.PRNGKey(8), n, np.float32).astype(np.float64)
    t = np.linspace(0.0, 1.0, n)
    v_ref, div_ref = VF.field_and_div(s.params, s.omega, x, t, s.ot, None, s.clip, act="tanh")
    v, div = model.apply(s.P, to_dev(x, cuda), to_dev(t, cuda), None, hutch=False, want_div=True)
    assert rel_err(v.cpu().numpy(), v_ref) < 1e-4
    assert np.abs(div.cpu().numpy() - div_ref).max() < 1e-4 * max(np.abs(div_ref).max(), 1.0)


def test_fm_sharded_batch_matches_full(cuda, setups):
    """Rows [lo,hi) of an n_total ensemble see the same (t, x0, eps) as in the unsharded call, so the
    shard losses/gradients add up to the full ones (the quantity the NCCL all-reduce sums)."""
    s = setups["phi-four"]
    n = 96
    x = to_dev(s.ot.init_positions(tf.PRNGKey(8), n, np.float32), cuda)
    key = key_dev(tf.PRNGKey(5), cuda)
    loss, grads = s.state.loss_and_grad(key, x)
    loss, grads = loss.clone(), grads.clone()
    tot_l, tot_g = 0.0, torch.zeros_like(grads)
    for lo, hi in ((0, 32), (32, 96)):
        l, g = s.state.loss_and_grad(key, x[lo:hi].contiguous(), chain_offset=lo, n_total=n)
        tot_l += l.item(); tot_g += g
    assert abs(tot_l - loss.item()) < 1e-5 * abs(loss.item())
    assert rel_err(tot_g.cpu().numpy(), grads.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("name", ["phi-four", "pines"])
def test_fm_two_part_backward_is_identical(cuda, lib, setups, name):
    """mfm_fm_loss_grad_part 1 then 2 (the split the multi-GPU host uses to overlap the all-reduce) gives
    bit-identical loss and gradients to the single call; part 1 alone already holds the Dense_4..7 slice."""
    from mfm_b200 import _lib
    s = setups[name]
    n = 160
    x = to_dev(s.ot.init_positions(tf.PRNGKey(3), n, np.float32), cuda)
    key = key_dev(tf.PRNGKey(11), cuda)
    loss, grads = s.state.loss_and_grad(key, x)
    loss, grads = loss.clone(), grads.clone()
    st = s.state
    fd, td = st.model.field_desc(st.P), st.model.dist._desc(1.0)
    ws = _lib.workspace(lib.mfm_fm_workspace_bytes(fd, td, n), cuda, "fm")
    g2 = torch.full_like(grads, float("nan")); l2 = torch.zeros_like(loss)
    split = st.P.w_off[4]
    for part in (1, 2):
        _lib.check(lib.mfm_fm_loss_grad_part(fd, td, _lib.ptr(key), n, 0, n, float(st.args.sigma), _lib.ptr(x), _lib.ptr(l2),
                                             _lib.ptr(g2), _lib.ptr(ws), ws.numel(), part, _lib.stream()))
        if part == 1:
            assert torch.equal(g2[split:], grads[split:]) and torch.equal(l2, loss)
            # the head is untouched so far, except the biases of Dense_3 and Dense_1: their gradients are column sums of
            # signals part 1 produces (and reduces inside the producing epilogue when that runs on the tensor cores)
            head = g2[:split].clone()
            for l in (1, 3):
                head[st.P.b_off[l]:st.P.b_off[l] + st.P.hidden] = 0
            assert (head == 0).all()
    assert torch.equal(g2, grads)


@pytest.mark.parametrize("warmup", [0, 3])
def test_adamw_clip_apply_if_finite(cuda, setups, warmup):
    """warmup = 3: --warmup_steps (exe_flow_matching.py:189-198): the device optimizer evaluates the joined schedule itself; the first
    step runs with learning rate 0, and a rejected update does not advance the schedule."""
    from mfm_b200 import exe_flow_matching as E
    s = setups["4-mode"]
    P = E.VectorFieldParams(2, 128, 128, cuda).load_dict(s.params)
    lr = E.create_learning_rate_fn(100, warmup, 1e-3)
    state = E.create_train_state(s.model, P, lr, _args())
    params = {"params": {k: {n: a.copy() for n, a in v.items()} for k, v in s.params["params"].items()}}
    opt = OO.AdamWClipIfFinite(params, OO.learning_rate_fn(100, warmup, 1e-3))
    rng = np.random.default_rng(0)
    for it in range(5):
        G = {"params": {k: {n: (rng.standard_normal(a.shape) * 10 ** rng.uniform(-6, 2)).astype(np.float32)
                            for n, a in v.items()} for k, v in params["params"].items()}}
        if it == 2:
            G["params"]["Dense_3"]["kernel"][0, 0] = np.nan          # rejected update
        if it == 3:
            G["params"]["Dense_0"]["bias"][5] = np.inf
        flat = torch.from_numpy(_flat_grads(SimpleNamespace(P=P), G).astype(np.float32)).to(cuda)
        state.apply_gradients(flat)
        params = opt.update(G, params)
        got = P.to_dict()
        for k in params["params"]:
            for nme in ("kernel", "bias"):
                np.testing.assert_allclose(got["params"][k][nme], params["params"][k][nme], rtol=2e-6, atol=2e-8)
        st = state.opt_state.cpu().tolist()
        assert st[0] == opt.count and st[1] == opt.notfinite_count and st[2] == opt.total_notfinite
        assert st[3] == int(opt.last_finite)
    assert opt.total_notfinite == 2 and opt.count == 3


def test_lr_schedule_matches_oracle(lib):
    from mfm_b200 import exe_flow_matching as E
    for warm in (0, 1, 40, 400):
        a, b = E.create_learning_rate_fn(400, warm, 1e-3), OO.learning_rate_fn(400, warm, 1e-3)
        for step in (0, 1, 39, 40, 57, 399, 400, 1000):
            assert a(step) == pytest.approx(b(step), rel=1e-12, abs=1e-18)
    assert E.create_learning_rate_fn(400, 40, 1e-3)(20) == pytest.approx(0.5e-3)


def test_tempering_beta_matches_oracle(cuda, lib):
    """beta_fn (exe_flow_matching.py:391-402): ESS-targeted bisection."""
    from mfm_b200 import _lib
    rng = np.random.default_rng(1)
    for n, scale, prev in [(128, 50.0, 0.0), (1024, 5.0, 0.3), (65536, 200.0, 0.0), (128, 1e-3, 0.5)]:
        ll = (rng.standard_normal(n) * scale - 100.0).astype(np.float32)
        exp = OO.tempering_beta(prev, ll, 0.95)
        out = torch.empty(1, dtype=torch.float32, device=cuda)
        ll_d, prev_d = to_dev(ll, cuda), torch.tensor([prev], dtype=torch.float32, device=cuda)   # keep alive
        _lib.check(lib.mfm_tempering_beta(ll_d.data_ptr(), n, prev_d.data_ptr(), 0.95, out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
        got = out.item()
        assert prev <= got <= 1.0
        assert abs(got - exp) <= 2e-3 * max(exp, 1e-3), (n, scale, got, exp)


def test_run_entry_point_small(cuda, lib):
    """run() drives tempering + MALA + flow-MH + FM updates end to end (4-mode, few iterations)."""
    from mfm_b200 import multi_modal as MM
    args = MM.parser().parse_args(["--example", "4-mode", "--learning_iter", "6", "--mcmc_per_flow_steps", "2",
                                   "--seed", "1", "--log_every", "1"])
    dist = MM.build(args, device=cuda)
    res = MM.run(dist, args, None, log_every=1)
    assert len(res["history"]) == 6 and 0.0 < res["final_beta"] <= 1.0
    assert all(np.isfinite(h["loss"]) for h in res["history"])
    loop = res["loop"]
    assert loop.state.opt_state.cpu().tolist()[0] == 6          # six applied AdamW updates
    assert torch.isfinite(loop.states.position).all()


def test_pipelined_update_is_identical(cuda, lib):
    """HotLoop(pipeline=True) - the multi-rank schedule: AdamW applied after the NEXT MALA step, before a flow-MH
    step - gives bit-identical chains, parameters and losses (a MALA step never reads the MLP)."""
    from types import SimpleNamespace
    from mfm_b200 import distributions as Dm, exe_flow_matching as E, random as mr
    d, H, F, n, m = 64, 128, 128, 96, 2
    args = SimpleNamespace(hutchs=False, num_importance_samples=0, mcmc_per_flow_steps=m, step_size=1e-4, ref_dist="stdgauss",
                           cond_flow=True, ot_cond_flow=False, sigma=1e-4, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8,
                           weight_decay=1e-4, gradient_clip=1.0, learning_iter=100, warmup_steps=0, learning_rate=1e-3)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=2)
    rng = np.random.default_rng(3)
    shapes = [(2 * F, H), (H, H), (d, H), (H, H), (H, d), (2 * H, H), (H, H), (H, d)]
    params = {"params": {f"Dense_{i}": {"kernel": (rng.standard_normal(s) / np.sqrt(s[0]) * (0.1 if i in (4, 7) else 1.0)).astype(np.float32),
                                        "bias": (rng.standard_normal(s[1]) * 0.01).astype(np.float32)} for i, s in enumerate(shapes)}}
    omega = torch.from_numpy(rng.standard_normal(F).astype(np.float32)).to(cuda)
    x0 = torch.from_numpy(rng.uniform(-1, 1, (n, d)).astype(np.float32)).to(cuda)
    out = []
    for pipe in (False, True):
        dist = Dm.PhiFour(d, device=cuda)
        model = E.VectorFieldNet(omega, dist, [H, H], [H, H], [H, H], "relu", None)
        P = E.VectorFieldParams(d, H, F, cuda).load_dict(params)
        loop = E.HotLoop(dist, model, P, args, opts, mr.PRNGKey(11, cuda), x0.clone(), pipeline=pipe)
        losses = [float(loop.iteration().item()) for _ in range(2 * (m + 1) + 1)]      # two flow-MH iterations inside
        if pipe:
            assert loop.state._pending is not None
            assert loop.state.opt_state.cpu().tolist()[0] == len(losses) - 1               # the last update is still in flight
        loop.flush()
        assert loop.state._pending is None and loop.state.opt_state.cpu().tolist()[0] == len(losses)
        out.append((losses, loop.states.position.clone(), P.flat.clone(), loop.state.mu.clone()))
    assert out[0][0] == out[1][0]
    for a, b in zip(out[0][1:], out[1][1:]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name", ["phi-four", "4-mode", "phi-four-1024"])
def test_graph_replayed_iterations_are_identical(cuda, lib, name):
    """HotLoop(graph=True) replays the MALA + FM-update iteration from a CUDA graph (latency-bound reference shapes):
    same kernels on the same buffers => bit-identical chains, parameters and losses, flow-MH iterations (eager) included."""
    from types import SimpleNamespace
    from mfm_b200 import distributions as Dm, exe_flow_matching as E, random as mr
    from oracle import targets as OT
    # (phi-four-1024: the reference's own chain count - tensor-core rows, narrow network: captured while no weight mirror exists)
    d, n, m, step = {"phi-four": (64, 96, 3, 1e-4), "phi-four-1024": (64, 1024, 3, 1e-4), "4-mode": (2, 128, 3, 0.2)}[name]
    H, F = 128, 128
    args = SimpleNamespace(hutchs=False, num_importance_samples=0, mcmc_per_flow_steps=m, step_size=step, ref_dist="stdgauss",
                           cond_flow=True, ot_cond_flow=False, sigma=1e-4, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8,
                           weight_decay=1e-4, gradient_clip=1.0, learning_iter=100, warmup_steps=0, learning_rate=1e-3)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=2)
    rng = np.random.default_rng(5)
    shapes = [(2 * F, H), (H, H), (d, H), (H, H), (H, d), (2 * H, H), (H, H), (H, d)]
    params = {"params": {f"Dense_{i}": {"kernel": (rng.standard_normal(s) / np.sqrt(s[0]) * (0.002 if i in (4, 7) else 1.0)).astype(np.float32),
                                        "bias": (rng.standard_normal(s[1]) * 0.01).astype(np.float32)} for i, s in enumerate(shapes)}}
    omega = torch.from_numpy(rng.standard_normal(F).astype(np.float32)).to(cuda)
    x0 = torch.from_numpy(rng.uniform(-1, 1, (n, d)).astype(np.float32)).to(cuda)
    out = []
    for graph in (False, True):
        if name.startswith("phi-four"):
            dist = Dm.PhiFour(d, device=cuda)
        else:
            t4 = OT.four_mode()
            dist = Dm.GaussianMixture(t4.modes, t4.covs, t4.weights, device=cuda)
        model = E.VectorFieldNet(omega, dist, [H, H], [H, H], [H, H], "relu", None)
        P = E.VectorFieldParams(d, H, F, cuda).load_dict(params)
        loop = E.HotLoop(dist, model, P, args, opts, mr.PRNGKey(21, cuda), x0.clone(), graph=graph)
        losses = [float(loop.iteration().item()) for _ in range(3 * (m + 1) + 2)]
        assert (loop._graph is not None) == graph
        assert loop.state.step == len(losses) and loop.state.opt_state.cpu().tolist()[0] == len(losses)
        out.append((losses, loop.states.position.clone(), loop.states.logdensity.clone(), P.flat.clone(), loop.key_sample.clone(),
                    loop.last_info.acceptance_rate.clone()))
    assert out[0][0] == out[1][0]
    for a, b in zip(out[0][1:], out[1][1:]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("name,n", [("phi-four", 200), ("pines", 24)])
def test_fm_loss_non_conditional_flow(cuda, setups, name, n):
    """--cond_flow off: flow_fn (exe_flow_matching.py:139-147) instead of cond_flow_fn; sharded rows see the global draws."""
    import copy
    s = setups[name]
    x = s.ot.init_positions(tf.PRNGKey(8), n, np.float32).astype(np.float64)
    key = tf.PRNGKey(4242)
    times, xt, target = VF.fm_batch_uncond(key, x, 1e-4, rng_dtype=np.float32)
    loss_ref, G = VF.fm_loss_and_grad(s.params, s.omega, xt, times, target, s.ot.grad, s.clip)
    state = copy.copy(s.state)
    state.args = copy.copy(s.state.args); state.args.cond_flow = False
    state.loss, state.grads = s.state.loss.clone(), s.state.grads.clone()
    loss, grads = state.loss_and_grad(key_dev(key, cuda), to_dev(x, cuda))
    assert abs(loss.item() - loss_ref) < 1e-4 * abs(loss_ref), (loss.item(), loss_ref)
    assert rel_err(grads.cpu().numpy(), _flat_grads(s, G)) < 2e-4
    full_l, full_g = loss.item(), grads.clone()
    tot_l, tot_g = 0.0, torch.zeros_like(full_g)
    for lo, hi in ((0, 8), (8, n)):
        l, g = state.loss_and_grad(key_dev(key, cuda), to_dev(x[lo:hi], cuda), chain_offset=lo, n_total=n)
        tot_l += l.item(); tot_g += g
    assert abs(tot_l - full_l) < 1e-5 * abs(full_l) and rel_err(tot_g.cpu().numpy(), full_g.cpu().numpy()) < 2e-5


def test_run_on_real_samples(cuda, lib):
    """mcmc_per_flow_steps < 0 (use_real_samples, exe_flow_matching.py:328,382-386): the FM update trains on draws of the target."""
    from mfm_b200 import multi_modal as MM, random as mr
    args = MM.parser().parse_args(["--example", "4-mode", "--learning_iter", "5", "--mcmc_per_flow_steps", "-1", "--seed", "1",
                                   "--eval_iter", "2", "--log_every", "1"])
    dist = MM.build(args, device=cuda)
    modes = torch.tensor(8.0 * np.array([[1, 1], [1, -1], [-1, 1], [-1, -1]]), dtype=torch.float32, device=cuda)

    def target_gn(keys):
        ks = mr.split(keys, 2)
        comp = (mr.uniform(ks[:, 0].contiguous(), (1,))[:, 0] * 4).long().clamp(max=3)
        return modes[comp] + mr.normal(ks[:, 1].contiguous(), (2,))

    with pytest.raises(ValueError):
        MM.run(dist, args, None)
    res = MM.run(dist, args, target_gn, log_every=1)
    assert len(res["history"]) == 5 and all(np.isfinite(h["loss"]) for h in res["history"])
    loop = res["loop"]
    assert loop.state.opt_state.cpu().tolist()[0] == 5 and loop.states.logdensity is None
    # the last training batch is a draw of the target: every point sits near one of the four modes
    d = (loop.states.position[:, None, :] - modes[None]).norm(dim=-1).min(dim=1).values
    assert d.max().item() < 6.0 and "MMD" in res["table"]
