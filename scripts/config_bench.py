"""Outer-iteration timings of the four reference configurations (BASELINE.json configs[0..3]) at THEIR shapes, on one GPU.

These are latency-bound (128 chains x d=2 is 1 KB of state): the numbers are reported as
microseconds per outer iteration and launches per iteration, next to chain-steps/s and FM-iterations/s
over one cycle of m MALA + 1 flow-MH iterations, each followed by an FM update.  'Trained-like' MLP fixture
(heads x0.1) so the ODE takes a realistic number of steps; beta = 1."""
import json
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, ".")
from mfm_b200 import _lib, exe_flow_matching as E, multi_modal as MM, random as mr      # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
CONFIGS = [("4-mode", ["--example", "4-mode", "--mcmc_per_flow_steps", "10"]),
           ("gaussian-mixture", ["--example", "gaussian-mixture", "--mcmc_per_flow_steps", "100", "--hutchs"]),
           ("phi-four", ["--example", "phi-four", "--mcmc_per_flow_steps", "1000"]),
           ("pines", ["--example", "pines", "--mcmc_per_flow_steps", "100", "--hutchs"])]
out = []
for name, argv in CONFIGS:
    args = MM.parser().parse_args(argv + ["--seed", "1"])
    dist = MM.build(args, device=dev)
    d, H, F, n, m = args.dim, args.hidden_x[0], args.fourier_dim, args.num_chain, int(args.mcmc_per_flow_steps)
    rng = np.random.default_rng(0)
    shapes = [(2 * F, H), (H, H), (d, H), (H, H), (H, d), (2 * H, H), (H, H), (H, d)]
    params = {"params": {f"Dense_{i}": {"kernel": (rng.standard_normal(s) / np.sqrt(s[0]) * (0.1 if i in (4, 7) else 1.0)).astype(np.float32),
                                        "bias": (rng.standard_normal(s[1]) * 0.01).astype(np.float32)} for i, s in enumerate(shapes)}}
    omega = torch.from_numpy(rng.standard_normal(F).astype(np.float32)).to(dev)
    model = E.VectorFieldNet(omega, dist, args.hidden_x, args.hidden_t, args.hidden_xt, "relu", args.gradient_clip if d > 128 else None)
    P = E.VectorFieldParams(d, H, F, dev).load_dict(params)
    keys = mr.split(mr.PRNGKey(1, dev), 6)
    dist.initialize_model(keys[3], n)
    opts = SimpleNamespace(rtol=args.rtol, atol=args.atol, mxstep=int(args.mxstep), n_times=5 if name == "4-mode" else 2)
    loop = E.HotLoop(dist, model, P, args, opts, keys[1], dist.init_params.contiguous(), beta=1.0)
    for _ in range(m + 1):                                   # warm-up: one full cycle
        loop.iteration()
    torch.cuda.synchronize()

    def timed(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.mfm_launch_count()
        e0.record()
        for _ in range(k):
            loop.iteration()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1), lib.mfm_launch_count() - l0

    loop.count = 0
    n_mala = min(m, 200)
    ms_mala, l_mala = timed(n_mala)                          # iterations 1..n_mala: MALA + FM update
    loop.count = m
    ms_flow, l_flow = timed(1)                               # iteration m+1: flow-MH + FM update
    stats = loop.gen.last_stats.get("ode")
    stats = stats.cpu().tolist() if stats is not None else None
    cyc_ms = m * ms_mala / n_mala + ms_flow
    rec = {"config": name, "chains": n, "dim": d, "hidden": H, "mcmc_per_flow_steps": m, "divergence": "hutchinson" if args.hutchs else "exact",
           "us_per_mala_iteration": 1e3 * ms_mala / n_mala, "launches_per_mala_iteration": l_mala / n_mala,
           "ms_per_flow_iteration": ms_flow, "launches_per_flow_iteration": l_flow,
           "ode_stats": dict(zip(["accepted", "attempted", "max_attempts_per_chain", "field_evals"], stats)) if stats else None,
           "ms_per_cycle": cyc_ms, "chain_steps_per_s": n * (m + 1) / (cyc_ms * 1e-3), "fm_iterations_per_s": (m + 1) / (cyc_ms * 1e-3)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
json.dump(out, open("gpurun_out/config_bench.json", "w"), indent=1)
