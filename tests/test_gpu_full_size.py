"""Parity at BASELINE.json's FULL size (65 536 pines chains, d = 1600, hidden 1024) through size-independent properties:
chains are independent, so any subset of the big ensemble must agree with the oracle run on that subset alone (the
per-chain random streams are closed-form functions of the global chain index), and sums over shards must add up.
These are the launch configurations bench.py times (persistent tcgen05 GEMMs over 1 024+ tiles, stream-K remainders,
pre-split weight operand, 419 MB state arrays)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import samplers as OS, targets as OT, threefry as tf, vector_field as VF
from tests.helpers import key_dev, rel_err, to_dev

pytestmark = pytest.mark.gpu

N, D, H, F = 65536, 1600, 1024, 128
ROWS = np.array([0, 1, 255, 256, 12345, 32767, 32768, 65535])


@pytest.fixture(scope="module")
def big(cuda, lib):
    from mfm_b200 import distributions as Dm, exe_flow_matching as E, random as mr
    ot = OT.LogGaussianCoxPines(D)
    dd = Dm.LogGaussianCoxPines(D, device=cuda)
    rng = np.random.default_rng(0)
    params = VF.init_params(rng, D, H, F, head_scale=0.1)
    omega = rng.standard_normal(F).astype(np.float32)
    model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", 1.0)
    P = E.VectorFieldParams(D, H, F, cuda).load_dict(params)
    key = mr.PRNGKey(1, cuda)
    eps = mr.normal(mr.split(key, N), (D,))
    x = (dd._mu_zero + eps @ dd._cholesky_gram.T).contiguous()
    del eps
    return SimpleNamespace(ot=ot, dd=dd, params=params, omega=omega, model=model, P=P, x=x, E=E, mr=mr)


def test_full_size_mala_subset_matches_oracle(cuda, big):
    from mfm_b200.bblackjax.mcmc import mala as M
    fn = big.dd.tempered(1.0)
    st = M.init(big.x, fn)
    sub = tf.PRNGKey(1024)
    new, info = M.mala_step(fn, key_dev(sub, cuda), st, 0.01, per_chain_keys=False)
    idx = torch.from_numpy(ROWS).to(cuda)
    x_sub = big.x[idx].cpu().numpy().astype(np.float64)
    st_o = OS.mala_init(x_sub, big.ot)
    assert rel_err(st.logdensity[idx].cpu().numpy(), st_o.logdensity) < 1e-4
    assert rel_err(st.logdensity_grad[idx].cpu().numpy(), st_o.logdensity_grad) < 1e-4
    keys = tf.split(sub, N)[ROWS]                          # chain i uses row i of split(key, N_total)
    new_o, info_o, dbg = OS.mala_step(keys, st_o, big.ot, 0.01, rng_dtype=np.float32)
    assert rel_err(info.proposed_position[idx].cpu().numpy(), info_o.proposed_position) < 1e-5
    acc_d, acc_o = info.is_accepted[idx].cpu().numpy(), info_o.is_accepted
    band = np.abs(info_o.acceptance_rate - dbg["u"]) < 1e-4 * np.maximum(1.0, np.abs(dbg["delta"]))
    assert ((acc_d == acc_o) | band).all()
    same = acc_d == acc_o
    assert rel_err(new.logdensity[idx].cpu().numpy()[same], new_o.logdensity[same]) < 1e-4
    assert rel_err(new.position[idx].cpu().numpy()[same], new_o.position[same]) < 1e-5
    assert 0.2 < info.is_accepted.float().mean().item() <= 1.0 and torch.isfinite(new.position).all()


def test_full_size_fm_loss_is_the_sum_of_its_shards_and_matches_oracle_rows(cuda, big):
    E = big.E
    args = SimpleNamespace(hutchs=True, num_importance_samples=0, mcmc_per_flow_steps=100, step_size=0.01, ref_dist="stdgauss",
                           cond_flow=True, ot_cond_flow=False, sigma=1e-4, adam_beta1=0.9, adam_beta2=0.999, adam_epsilon=1e-8,
                           weight_decay=1e-4, gradient_clip=1.0, learning_iter=10000, warmup_steps=0, learning_rate=1e-3)
    state = E.create_train_state(big.model, big.P, E.create_learning_rate_fn(10000, 0, 1e-3), args)
    key = key_dev(tf.PRNGKey(5), cuda)
    loss, grads = state.loss_and_grad(key, big.x)
    loss, grads = loss.clone(), grads.clone()
    assert torch.isfinite(grads).all() and np.isfinite(loss.item())
    # 8 shards of 8 192 chains (the 8-GPU decomposition): losses and gradients add up to the full ones
    tot_l, tot_g = 0.0, torch.zeros_like(grads)
    for r in range(8):
        lo = r * 8192
        l, g = state.loss_and_grad(key, big.x[lo:lo + 8192].contiguous(), chain_offset=lo, n_total=N)
        tot_l += l.item(); tot_g += g
    assert abs(tot_l - loss.item()) < 2e-5 * abs(loss.item())
    # weight gradients reduce over the chains: the tensor core's truncating fp32 accumulator biases a split-K slice by
    # ~8e-9 per accumulated term, and full / sharded calls cut the 65 536 terms differently
    assert rel_err(tot_g.cpu().numpy(), grads.cpu().numpy()) < 2e-4
    # a 16-chain shard in the middle of the ensemble against the oracle fed with the global draws of those rows
    lo, m = 40000, 16
    l16, g16 = state.loss_and_grad(key, big.x[lo:lo + m].contiguous(), chain_offset=lo, n_total=N)
    k_time, k_ref, k_gauss, _ = tf.split(tf.PRNGKey(5), 4)
    xs = big.x[lo:lo + m].cpu().numpy().astype(np.float64)
    # rows lo..lo+m of the global uniform(key_time, (N,1)) and normal(key_gauss, (N,D)) draws: words of the halves layout
    bits_t = tf.random_bits(k_time, 32, (N,))[lo:lo + m]
    times = ((bits_t >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32).astype(np.float64) - 1.0
    ref = np.stack([tf.normal(k, (D,), np.float32) for k in tf.split(k_ref, N)[lo:lo + m]]).astype(np.float64)
    total = N * D
    half = total // 2
    w = np.arange(lo * D, (lo + m) * D, dtype=np.uint64)
    first = w < half
    c_lo = np.where(first, w, w - half).astype(np.uint32); c_hi = (c_lo.astype(np.uint64) + half).astype(np.uint32)
    o0, o1 = tf.threefry2x32(k_gauss[0], k_gauss[1], c_lo, c_hi)
    bits = np.where(first, o0, o1).astype(np.uint32)
    lo_f = np.nextafter(np.float32(-1), np.float32(0))
    u = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)
    u = np.maximum(lo_f, u * (np.float32(1) - lo_f) + lo_f)
    eps = (np.float32(np.sqrt(2)) * tf.erf_inv_f32(u)).astype(np.float64).reshape(m, D)
    xt = 1e-4 * eps + times[:, None] * xs + (1 - times[:, None]) * ref
    loss_ref, G = VF.fm_loss_and_grad(big.params, big.omega, xt, times, xs - ref, big.ot.grad, 1.0)
    assert abs(l16.item() - loss_ref) < 1e-4 * abs(loss_ref)
    g_ref = E.VectorFieldParams(D, H, F, cuda).load_dict(G).flat.cpu().numpy()
    assert rel_err(g16.cpu().numpy(), g_ref) < 5e-4


def test_full_size_flow_push_subset_matches_small_run(cuda, big):
    """transform_and_logdet on all 65 536 chains vs the same 8 chains pushed alone (same probe keys): lock-step iteration,
    active-chain compaction and the big-GEMM kernels must not couple chains."""
    E = big.E
    args = SimpleNamespace(hutchs=True, num_importance_samples=0, mcmc_per_flow_steps=100, step_size=0.01)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=2)
    _, _, push = E.create_train_data_gn(big.dd, big.model, opts, args)
    keys = big.mr.split(big.mr.PRNGKey(9, cuda), N)
    u = big.mr.normal(big.mr.split(big.mr.PRNGKey(10, cuda), N), (D,))
    stats = torch.zeros(8, dtype=torch.int32, device=cuda)
    y, ldj = push(keys, u, big.P, stats)
    idx = torch.from_numpy(ROWS).to(cuda)
    keys8 = keys.view(torch.int32)[idx].contiguous().view(torch.uint32)      # torch cannot index uint32 tensors
    y8, ldj8 = push(keys8, u[idx].contiguous(), big.P)
    assert torch.isfinite(y).all() and torch.isfinite(ldj).all()
    acc, tried, mx, nev = stats.cpu().tolist()[:4]
    chain_evals = int(stats[4:6].cpu().view(torch.int64).item())
    assert 2 * N <= chain_evals <= N * nev
    assert nev == 2 + 6 * mx and acc <= tried
    # different kernels serve 8 rows (warp-level MMA) and 65 536 rows (persistent tcgen05): agreement to the ODE tolerance
    assert rel_err(y[idx].cpu().numpy(), y8.cpu().numpy()) < 1e-3
    # log-det: a d-term cancelling Hutchinson sum per field evaluation, ~1e-6 * d absolute round-off each,
    # accumulated over the ~14 accepted steps of the solve by two different GEMM kernels
    dl = np.abs(ldj[idx].cpu().numpy() - ldj8.cpu().numpy()).max()
    assert dl < 3e-2, dl


def test_full_size_flow_mh_decisions_match_oracle_subset(cuda, big):
    """The flow-MH step (random-walk MH in latent space, exe_flow_matching.py:264-278) on all 65 536 chains; eight chains spread
    over the ensemble against the float64 oracle run on those chains alone (their keys are rows of split(key, 65536)).  Same
    criteria as tests/test_gpu_flow.py::test_flow_mh_decision_flip_rate: log alpha no further from the float64 oracle than the
    float32 oracle is (the solver's tolerance, not the arithmetic, bounds it), and a decision may only differ inside that band."""
    E = big.E
    args = SimpleNamespace(hutchs=True, num_importance_samples=0, mcmc_per_flow_steps=100, step_size=0.01)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=2)
    gen, init_fn, _ = E.create_train_data_gn(big.dd, big.model, opts, args)
    st_d = init_fn(big.x.clone(), 1.0)
    key = tf.PRNGKey(777)
    new_d, info_d = gen.flow_step(key_dev(key, cuda), st_d, big.dd.tempered(1.0), big.P)
    assert torch.isfinite(new_d.position).all() and torch.isfinite(new_d.logdensity).all()
    rate = info_d.is_accepted.float().mean().item()
    assert 0.0 <= rate <= 1.0
    idx = torch.from_numpy(ROWS).to(cuda)
    x0 = big.x[idx].cpu().numpy().astype(np.float64)
    st_o = OS.mala_init(x0, big.ot, 1.0)
    st_o32 = OS.MALAState(*[a.astype(np.float32) for a in st_o])
    flow = OS.Flow(big.params, big.omega, big.ot, True, 1e-5, 1e-5, 1000, 1.0, np.linspace(0.0, 1.0, 2), rng_dtype=np.float32)
    keys = tf.split(key, N)[ROWS]
    dbg, dbg32 = {}, {}
    _, info_o = OS.rw_flow_mh_step(keys, st_o, big.ot, flow, 1.0, dbg)
    OS.rw_flow_mh_step(keys, st_o32, big.ot, flow, 1.0, dbg32)
    la, la32 = dbg["log_acc"], dbg32["log_acc"].astype(np.float64)
    with np.errstate(over="ignore", divide="ignore"):
        la_d = np.log(info_d.acceptance_rate[idx].cpu().numpy().astype(np.float64))
        logu = np.log(dbg["u"])
    fin = np.isfinite(la) & np.isfinite(la_d) & np.isfinite(la32) & (np.abs(la) < 80)
    assert fin.sum() >= 4
    err, err32 = np.abs(la_d - la)[fin], np.abs(la32 - la)[fin]
    # 8 chains of a heavy-tailed error: the maximum may sit on one chain whose accept / reject sequence splits early
    assert np.median(err) <= 3.0 * np.median(err32) + 5e-3, (np.median(err), np.median(err32))
    assert err.max() <= 10.0 * err32.max() + 5e-2, (err.max(), err32.max())
    acc_d = info_d.is_accepted[idx].cpu().numpy().astype(bool); acc_o = info_o.is_accepted
    band = np.abs(la - logu) <= np.abs(la_d - la) + 1e-6
    assert ((acc_d == acc_o) | band | ~fin).all()
    assert int((acc_d != acc_o).sum()) <= 1
