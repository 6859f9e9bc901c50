"""time one fused small-shape ODE solve (4-mode shape) and report us per field evaluation"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zlib
from types import SimpleNamespace
import numpy as np, torch
from oracle import threefry as tf, vector_field as VF
from tests.helpers import key_dev, make_targets, to_dev
from mfm_b200 import exe_flow_matching as E, _lib
lib = _lib.load(); cuda = torch.device("cuda:0")
CFG = {"4-mode": (128, False, 5), "gmm16": (128, True, 2)}
for name, ot, dd in make_targets(cuda):
    if name not in CFG: continue
    H, hutch, n_times = CFG[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)
    params = VF.init_params(rng, ot.dim, H, 128, head_scale=0.2); omega = rng.standard_normal(128).astype(np.float32)
    model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", None)
    P = E.VectorFieldParams(ot.dim, H, 128, cuda).load_dict(params)
    args = SimpleNamespace(hutchs=hutch, num_importance_samples=0, mcmc_per_flow_steps=10, step_size=0.1)
    opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=n_times)
    gen, init_fn, push = E.create_train_data_gn(dd, model, opts, args)
    n = 128
    keys = key_dev(tf.split(tf.PRNGKey(7), n), cuda); u = to_dev(tf.vmap_normal(tf.split(tf.PRNGKey(8), n), ot.dim), cuda)
    stats = torch.zeros(8, dtype=torch.int32, device=cuda)
    for _ in range(2): push(keys, u, P, stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats.zero_(); e0.record()
    for _ in range(5): push(keys, u, P, stats)
    e1.record(); torch.cuda.synchronize()
    st = stats.cpu().tolist()
    ms = e0.elapsed_time(e1) / 5
    print(name, "ms per solve", round(ms, 3), "max attempts", st[2], "evals", st[3], "us per field eval", round(1000 * ms / max(st[3], 1), 2), flush=True)
