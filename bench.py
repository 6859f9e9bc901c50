#!/usr/bin/env python
"""bench.py — chain-steps/sec of the MFM hot path (MALA + flow-MH + FM update), pines 40x40.

Contract: `python bench.py --gpus N --steps K --warmup W` (torchrun for N>1) prints ONE JSON line.

Workload (BASELINE.json configs[4]): pines-shaped ensemble, 65 536 chains in
total, sharded over the N GPUs (strong scaling: total work fixed), d=1600, H=1024, step 0.01,
mcmc_per_flow_steps m, Hutchinson divergence, rtol=atol=1e-5, beta=1, synthetic positions
mu + L eps, "trained-like" MLP fixture (all kernels ~ N(0, 1/fan_in), heads x0.1, numpy seed 0).

A STEP is one cycle-aligned block of (m+1) outer iterations of the reference loop
(exe_flow_matching.py:432-449): m MALA iterations + 1 flow-MH iteration, EACH followed by one
flow-matching AdamW update.  value = n_total * K * (m+1) / seconds  [chain-steps/s].

`--impl reference` times the CPU oracle port of the same path (JAX is not installable here, see
DESIGN.md) on the host cores, on a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chain-steps/sec (MALA+flow-MH+FM update), pines 40x40"
UNIT = "chain-steps/s"
D, H, F = 1600, 1024, 128


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=1)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    p.add_argument("--chains", type=int, default=65536, help="total chains over all GPUs")
    p.add_argument("--mcmc_per_flow_steps", dest="m", type=int, default=100)
    p.add_argument("--head_scale", type=float, default=0.1)
    p.add_argument("--warmup_unit", type=str, default="cycle", choices=["iteration", "cycle"],
                   help="a warm-up step is one full cycle (default, = a timed step) or one outer iteration (>=3 of them touch every MALA/FM kernel, and "
                        "one extra flow-MH iteration is then run untimed)")
    p.add_argument("--no_cpu_baseline", action="store_true")
    p.add_argument("--no_e2e", action="store_true")
    return p.parse_args()


def args_ns(m, learning_iter=10000):
    return SimpleNamespace(hutchs=True, num_importance_samples=0, mcmc_per_flow_steps=m, step_size=0.01,
                           ref_dist="stdgauss", cond_flow=True, ot_cond_flow=False, sigma=1e-4, adam_beta1=0.9,
                           adam_beta2=0.999, adam_epsilon=1e-8, weight_decay=1e-4, gradient_clip=1.0,
                           learning_iter=learning_iter, warmup_steps=0, learning_rate=1e-3)


def fixture_params(head_scale):
    """Trained-like MLP fixture: fixed numpy seed, heads scaled down."""
    rng = np.random.default_rng(0)
    shapes = [(2 * F, H), (H, H), (D, H), (H, H), (H, D), (2 * H, H), (H, H), (H, D)]
    p = {}
    for i, (fi, fo) in enumerate(shapes):
        s = (1.0 / np.sqrt(fi)) * (head_scale if i in (4, 7) else 1.0)
        p[f"Dense_{i}"] = {"kernel": (rng.standard_normal((fi, fo)) * s).astype(np.float32),
                           "bias": (rng.standard_normal(fo) * 0.01).astype(np.float32)}
    omega = rng.standard_normal(F).astype(np.float32)
    return {"params": p}, omega


# ------------------------------------------------------------------------------------------------
# algorithmic work
# ------------------------------------------------------------------------------------------------
def flops_per_chain():
    P = 2 * F * H + 5 * H * H + 3 * D * H
    T = 2 * D * H + 3 * H * H
    return dict(mala=2 * D * D, fm=6 * P + 2 * D * D, field=2 * P + 2 * T + 2 * D * D, logp=2 * D * D)


class ClockSampler:
    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.path)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU oracle timing (reference arm + cpu_baseline)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(m, head_scale, budget_s=20.0, n=None):
    """chain-steps/s of the NumPy oracle (float64, as the reference runs with jax_enable_x64) on a
    bounded sample: n chains, a few MALA+FM iterations and one flow-MH+FM iteration, combined into
    one cycle  n*(m+1) / (m*t_mala_iter + t_flow_iter)."""
    from oracle import optim as OO, samplers as OS, targets as OT, threefry as tf, vector_field as VF
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1 to every rank, and small-batch NumPy GEMMs do not always gain from threads (measured:
    # a 16-chain flow iteration takes 29 s with 8 OpenBLAS threads and 6 s with 1 on a shared 8-core host).  Give the CPU its
    # best shot: the thread count is set explicitly and calibrated below (1 vs all cores); `cores` reports what was used.
    try:
        from threadpoolctl import threadpool_limits
    except Exception:
        threadpool_limits = None
    n = n or 16
    ot = OT.LogGaussianCoxPines(D)
    params, omega = fixture_params(head_scale)
    ref = OT.IndepGaussian(D)
    flow = OS.Flow(params, omega, ot, True, 1e-5, 1e-5, 1000, 1.0, (0.0, 1.0))
    opt = OO.AdamWClipIfFinite(params, OO.learning_rate_fn(10000, 0, 1e-3))
    x0 = ot.init_positions(tf.PRNGKey(1), n, np.float64)
    st = OS.mala_init(x0, ot)
    key = tf.PRNGKey(7)

    def fm_update(key_step, pos, params):
        times, xt, target = VF.fm_batch(key_step, pos, ref.sample, 1e-4)
        loss, G = VF.fm_loss_and_grad(params, omega, xt, times, target, ot.grad, 1.0)
        return opt.update(G, params)

    def pick_threads(fn):
        """faster of 1 / all cores for this phase (BLAS threads; 2 calls each after one warm-up call)"""
        if threadpool_limits is None:
            return cores
        best = None
        for cand in sorted({1, cores}):
            threadpool_limits(limits=cand)
            fn()
            tc = time.perf_counter()
            fn(); fn()
            tc = time.perf_counter() - tc
            if best is None or tc < best[0]:
                best = (tc, cand)
        threadpool_limits(limits=best[1])
        return best[1]

    def one_mala_fm():
        nonlocal key, st, params
        key, k1, k2 = tf.split(key, 3)
        st, _, _ = OS.mala_step(tf.split(k1, n), st, ot, 0.01)
        params = fm_update(k2, st.position, params); flow.params = params

    zc = np.random.default_rng(1).standard_normal((n, D))
    tcal = np.full(n, 0.5)
    th_mala = pick_threads(one_mala_fm)
    # MALA + FM iterations
    t0 = time.perf_counter(); it = 0
    while it < 2 or (time.perf_counter() - t0 < 0.35 * budget_s and it < 50):
        key, k1, k2 = tf.split(key, 3)
        st, _, _ = OS.mala_step(tf.split(k1, n), st, ot, 0.01)
        params = fm_update(k2, st.position, params); flow.params = params
        it += 1
    t_mala = (time.perf_counter() - t0) / it
    th_flow = pick_threads(lambda: VF.field_and_div(params, omega, st.position, tcal, ot, zc, 1.0))
    cores = max(th_mala, th_flow)
    # one flow-MH + FM iteration
    t0 = time.perf_counter()
    key, k1, k2 = tf.split(key, 3)
    stats = {}
    st, _ = OS.rw_flow_mh_step(tf.split(k1, n), st, ot, flow, 1.0, stats)
    params = fm_update(k2, st.position, params)
    t_flow = time.perf_counter() - t0
    rate = n * (m + 1) / (m * t_mala + t_flow)
    sample = (f"{n} chains (d=1600,H=1024), float64 NumPy oracle, BLAS threads = faster of 1 / all {os.cpu_count()} cores per phase ({th_mala} for MALA+FM, {th_flow} for the flow): {it} MALA+FM iterations ({t_mala*1e3:.0f} ms each) + "
              f"1 flow-MH+FM iteration ({t_flow:.1f} s, {int(stats['inv']['n_try'].max())}+{int(stats['fwd']['n_try'].max())} "
              f"RK steps); cycle = {m}*t_mala + t_flow")
    return rate, cores, sample, m * t_mala + t_flow


REFERENCE_SAMPLE_CHAINS = 128    # the reference arm's sample = the reference's own pines ensemble (multi_modal.py:90); the
                                 # cpu_baseline object inside the default run uses 16 chains to stay short


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    rates = []
    n = REFERENCE_SAMPLE_CHAINS
    for _ in range(max(1, min(a.steps, 2))):
        rate, cores, sample, cyc = cpu_oracle_rate(a.m, a.head_scale, budget_s=15.0, n=n)
        rates.append(rate)
    v = float(np.mean(rates))
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * n * (a.m + 1) / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": workload_config(a, a.chains),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle port (NumPy restatement of the reference; JAX/flax/optax cannot be installed in this "
                    f"image); rate measured on a {n}-chain sample (the size of the reference's own pines run) and reported per "
                    "chain-step, i.e. NOT extrapolated to 65536 chains' wall-clock; ms_per_step is the sample's cycle time",
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)


def workload_config(a, n_total):
    return {"workload": "pines-shaped scaling ensemble (BASELINE.json configs[4])", "chains_total": n_total, "dim": D,
            "hidden": H, "fourier_dim": F, "mcmc_per_flow_steps": a.m, "step_size": 0.01, "divergence": "hutchinson",
            "rtol": 1e-5, "atol": 1e-5, "beta": 1.0, "flow_step": "random-walk MH in latent space (reference default)",
            "mlp_params": f"trained-like fixture, heads x{a.head_scale}, numpy seed 0",
            "step_definition": f"{a.m} MALA + 1 flow-MH outer iterations, each followed by one FM AdamW update",
            "l2": "inputs larger than L2 (state arrays 419 MB each at 65536 chains)",
            "parallelism": f"chains sharded over {a.gpus} GPU(s), FM-gradient all-reduce (NCCL)"}


# ------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    # The contract is ONE JSON line on stdout: route everything else that libraries print to fd 1
    # (e.g. NCCL's version banner) to stderr, and keep a private handle for the result line.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as tdist
    from mfm_b200 import _lib, distributions as Dm, exe_flow_matching as E, random as mr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        tdist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    n_total = a.chains
    assert n_total % world == 0
    n = n_total // world
    off = rank * n
    m = a.m
    cyc = m + 1
    args = args_ns(m)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=2)

    dist = Dm.LogGaussianCoxPines(D, device=dev)
    params, omega = fixture_params(a.head_scale)
    model = E.VectorFieldNet(torch.from_numpy(omega).to(dev), dist, [H, H], [H, H], [H, H], "relu", 1.0)
    P = E.VectorFieldParams(D, H, F, dev).load_dict(params)

    # synthetic positions mu + L eps for this rank's rows of split(key_dist, n_total)
    key0 = mr.PRNGKey(1, dev)
    ks = mr.split(key0, 6)
    key_sample, key_dist = ks[1].clone(), ks[3].clone()
    rows = mr.split(key_dist, n_total)[off:off + n].contiguous()
    eps = mr.normal(rows, (D,))
    pos0 = (dist._mu_zero + eps @ dist._cholesky_gram.T).contiguous()
    del eps, rows

    loop = E.HotLoop(dist, model, P, args, opts, key_sample, pos0, beta=1.0, chain_offset=off, n_total=n_total)

    def sync():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def run_cycle():
        # iterations count+1 .. count+cyc with exactly one flow-MH iteration (count % (m+1) == 0)
        for _ in range(cyc):
            loop.iteration()

    # ---- warm-up -------------------------------------------------------------------------------
    if a.warmup_unit == "cycle":
        for _ in range(a.warmup):
            run_cycle()
    else:
        for _ in range(a.warmup):
            loop.iteration()              # MALA + FM update iterations
        # one untimed flow-MH iteration so the ODE kernels are warm too
        saved = loop.count
        loop.count = cyc - 1
        loop.iteration()
        loop.count = saved
        # realign: the timed window must start right after a multiple of (m+1)
        loop.count = 0
    loop.flush()
    sync()

    # ---- timed region --------------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = lib.mfm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for _ in range(a.steps):
        run_cycle()
    loop.flush()                          # multi-rank runs pipeline the last AdamW update: it belongs to the timed work
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = lib.mfm_launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_total * a.steps * cyc / (ms / 1e3)
    ode_stats = loop.gen.last_stats.get("ode")
    ode_stats = ode_stats.cpu().tolist() if ode_stats is not None else None

    # ---- per-phase timing (MALA iteration, FM update, flow iteration) for the roofline ------------
    def timed(fn, reps):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s0.record()
        for _ in range(reps):
            fn()
        s1.record(); torch.cuda.synchronize()
        return s0.elapsed_time(s1) / reps

    fl = flops_per_chain()
    key_t = mr.PRNGKey(99, dev)
    loop.flush()
    ms_fm = timed(lambda: loop.state.loss_and_grad(key_t, loop.states.position, off, n_total), 3)
    from mfm_b200.bblackjax.mcmc.mala import mala_step
    ms_mala = timed(lambda: mala_step(dist.tempered(1.0), key_t, loop.states, 0.01, False, off, n_total, inplace=True), 3)
    # one whole MALA outer iteration (data generator + FM loss/grad + gradient all-reduce + AdamW), max over ranks
    loop.flush()
    loop.count = 0
    ms_iter = timed(loop.iteration, 5)
    loop.flush()
    t_it = torch.tensor([ms_iter], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(t_it, op=tdist.ReduceOp.MAX)
    ms_iter = float(t_it.item())
    # dominant kernel: the persistent CTA-pair dense-layer GEMM, measured on the FM hidden-layer shape
    # [n,H] x [H,H] with K-major operands (how every forward / backward-data layer calls it): a burst of 10
    # launches and a sustained run of >= 1 s (the clocks settle under the 1 kW power cap), CUDA events on
    # the launching stream.
    Ag = torch.randn(n, H, device=dev); Bg = torch.randn(H, H, device=dev); Cg = torch.empty(n, H, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    gemm = lambda: _lib.check(lib.mfm_gemm_tf32x3(n, H, H, Ag.data_ptr(), H, 1, Bg.data_ptr(), H, 0, None, 0, Cg.data_ptr(), H, st))
    # as the layers call it: the weight operand's bf16 cross tile pre-split once (per parameter update) and loaded by TMA
    Bx = torch.empty_like(Bg)
    _lib.check(lib.mfm_gemm_presplit(Bg.data_ptr(), Bx.data_ptr(), H * H, st))
    lib.mfm_gemm_register_mirror(Bg.data_ptr(), H * H, Bx.data_ptr())
    ms_gemm_burst = timed(gemm, 10)
    ms_gemm = timed(gemm, max(10, int(1000.0 / ms_gemm_burst)))
    lib.mfm_gemm_register_mirror(None, 0, None)
    gemm_tflops = 2.0 * n * H * H / (ms_gemm * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1590.0 if not peaks else peaks.get("bf16_tflops", 1590.0)))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1590 (of fallback)"
    step_flops = n * (cyc * fl["fm"] + m * fl["mala"])          # + flow iteration (data dependent), added below
    if ode_stats:
        step_flops += n * (ode_stats[3] * fl["field"] + fl["logp"])
    step_tflops = step_flops * a.steps / (ms * 1e-3) / 1e12 * 1.0
    # per-launch DRAM traffic of this kernel at n = 65536 from the committed ncu --set full capture
    # (profiles/r01_ncu_pair_kernels.md: dram__bytes_read.sum + dram__bytes_write.sum); algorithmic bytes are
    # A + C + W = 2 * n*H*4 + H*H*4.  Only quoted when the run has the profiled shape.
    traffic = 501.6e6 if n == 65536 else None
    roofline = {"bound": "tensor",
                "kernel": "tc2p::gemm_tc2p_kernel (persistent CTA-pair tcgen05/TMEM/TMA dense layer, FM shape [n,1024]x[1024,1024], K-major operands)",
                "achieved": gemm_tflops, "peak": peak_tf, "unit": "TFLOP/s", "frac": gemm_tflops / peak_tf,
                "traffic": traffic, "algorithmic_bytes": 2.0 * n * H * 4 + H * H * 4, "peak_source": peak_src,
                "achieved_burst": 2.0 * n * H * H / (ms_gemm_burst * 1e-3) / 1e12,
                "note": "algorithmic fp32 FLOPs (2*M*N*K per launch), sustained (>= 1 s of back-to-back launches, sw_power_cap active). "
                        "fp32-accurate emulation: per k-step one kind::tf32 MMA (hi*hi) + one kind::f16 bf16 MMA with K=16 (both cross "
                        "terms; the weight operand's cross tile is pre-split in global memory), i.e. 2 tf32-rate MMA slots per fp32 product => the ceiling of this arithmetic is 1/4 of the bf16 peak "
                        "(split-K weight-gradient GEMMs still use 3 tf32 passes, ceiling 1/6)",
                "frac_of_emulation_ceiling": gemm_tflops / (peak_tf / 4.0),
                "whole_step_tflops_per_gpu": step_tflops,
                "phase_ms": {"fm_loss_grad": ms_fm, "mala_iteration": ms_mala, "outer_iteration_mala": ms_iter,
                             "outer_iteration_flow": ms / a.steps - m * ms_iter},
                "phase_tflops": {"fm_loss_grad": n * fl["fm"] / (ms_fm * 1e-3) / 1e12,
                                 "mala_iteration": n * fl["mala"] / (ms_mala * 1e-3) / 1e12},
                "mala_state_gbs": n * (20 * D + 28) / (ms_mala * 1e-3) / 1e9}
    del Ag, Bg, Cg, Bx

    # ---- end-to-end through the public API with HOST buffers ---------------------------------------
    e2e = None
    if not a.no_e2e:
        host_in = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
        host_in.copy_(loop.states.position)
        host_out = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
        host_loss = torch.empty(1, dtype=torch.float32, pin_memory=True)
        dev_in = torch.empty((n, D), dtype=torch.float32, device=dev)
        loop.count = 0
        sync()
        e0.record()
        for _ in range(a.steps):
            dev_in.copy_(host_in, non_blocking=True)          # H2D: this step's chain positions
            loop.reset_positions(dev_in)                      # public init_fn: logdensity + grad
            loss = None
            for _ in range(cyc):
                loss = loop.iteration()
            loop.flush()
            host_out.copy_(loop.states.position, non_blocking=True)   # D2H: new positions + loss
            host_loss.copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            host_in.copy_(host_out)
        e1.record()
        sync()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        e2e = {"value": n_total * a.steps * cyc / (float(t.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": n * D * 4, "d2h_bytes_per_step": n * D * 4 + 4,
               "api": "HotLoop.reset_positions(init_fn) + (m+1) x HotLoop.iteration() per step, pinned host buffers"}

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        rate, cores, sample, _ = cpu_oracle_rate(m, a.head_scale, budget_s=15.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (dense layers: fp32 emulated on tensor cores, tf32 hi*hi + bf16 cross terms / 3xTF32, fp32 accumulate)", "data": "synthetic",
                "config": workload_config(a, n_total), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu,
                "fm_iterations_per_s": a.steps * cyc / (ms / 1e3),
                "ode_stats_last_flow_step": dict(zip(["accepted", "attempted", "max_attempts_per_chain", "field_evals"],
                                                     ode_stats)) if ode_stats else None,
                "warmup_unit": a.warmup_unit}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
