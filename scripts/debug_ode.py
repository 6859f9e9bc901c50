import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from types import SimpleNamespace
from oracle import samplers as OS, targets as OT, threefry as tf, vector_field as VF
from tests.helpers import key_dev, make_targets, rel_err, to_dev
from mfm_b200 import exe_flow_matching as E
cuda = torch.device("cuda", 0)
CFG = {"4-mode": (128, False, 5, None, 8), "gmm16": (128, True, 2, None, 8), "phi-four": (128, False, 2, None, 6), "pines": (1024, True, 2, 1.0, 4)}
for name, ot, dd in make_targets(cuda):
    H, hutch, n_times, clip, n = CFG[name]
    for hs in (0.02, 0.2):
        rng = np.random.default_rng(3)
        params = VF.init_params(rng, ot.dim, H, 128, head_scale=hs)
        omega = rng.standard_normal(128).astype(np.float32)
        model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", clip)
        P = E.VectorFieldParams(ot.dim, H, 128, cuda).load_dict(params)
        args = SimpleNamespace(hutchs=hutch, num_importance_samples=0, mcmc_per_flow_steps=10, step_size=0.1)
        opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=n_times)
        gen, init_fn, push = E.create_train_data_gn(dd, model, opts, args)
        keys = tf.split(tf.PRNGKey(77), n)
        u = tf.vmap_normal(tf.split(tf.PRNGKey(4), n), ot.dim).astype(np.float64)
        stats = torch.zeros(4, dtype=torch.int32, device=cuda)
        x, ldj = push(key_dev(keys, cuda), to_dev(u, cuda), P, stats)
        x = x.cpu().numpy(); ldj = ldj.cpu().numpy()
        for dtp in (np.float64, np.float32):
            flow = OS.Flow(params, omega, ot, hutch, 1e-5, 1e-5, 1000, clip, np.linspace(0, 1, n_times), rng_dtype=np.float32)
            st = {}
            xr, lr = flow.transform_and_logdet(keys, u.astype(dtp), st)
            print(name, "hs", hs, np.dtype(dtp).name, "x relerr %.2e" % rel_err(x, xr), "ldj abs %.2e (scale %.2f)" % (np.abs(ldj - lr).max(), np.abs(lr).max()),
                  "cuda stats", stats.cpu().tolist(), "oracle try", st["n_try"].tolist(), "acc", st["n_acc"].tolist(), flush=True)
