"""debug: field evaluation and push at n >= 256 (tcgen05 path) vs the float64 oracle, fused and general ODE paths"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zlib
from types import SimpleNamespace
import numpy as np, torch
from oracle import samplers as OS, threefry as tf, vector_field as VF
from tests.helpers import key_dev, make_targets, rel_err, to_dev
from mfm_b200 import exe_flow_matching as E, _lib
lib = _lib.load()
cuda = torch.device("cuda:0")
CFG = {"4-mode": (128, False, 5, None), "gmm16": (128, True, 2, None)}
for name, ot, dd in make_targets(cuda):
    if name not in CFG: continue
    H, hutch, n_times, clip = CFG[name]
    rng = np.random.default_rng(zlib.crc32(name.encode()) % 1000)      # (hash() of a str is salted per process)
    params = VF.init_params(rng, ot.dim, H, 128, head_scale=0.2)
    omega = rng.standard_normal(128).astype(np.float32)
    model = E.VectorFieldNet(to_dev(omega, cuda), dd, [H, H], [H, H], [H, H], "relu", clip)
    P = E.VectorFieldParams(ot.dim, H, 128, cuda).load_dict(params)
    for n in (24, 256, 333, 512):
        x = ot.init_positions(tf.PRNGKey(1), n, np.float32).astype(np.float64)
        t = np.linspace(0.0, 1.3, n)
        z = np.random.default_rng(5).standard_normal(x.shape) if hutch else None
        v_ref, div_ref = VF.field_and_div(params, omega, x, t, ot, z, clip)
        v, div = model.apply(P, to_dev(x, cuda), to_dev(t, cuda), to_dev(z, cuda) if hutch else None, hutch=hutch, want_div=True)
        print(name, n, "field err", rel_err(v.cpu().numpy(), v_ref), "div err", np.abs(div.cpu().numpy() - div_ref).max() / max(np.abs(div_ref).max(), 1.0), flush=True)
        lib.mfm_set_gemm_h16(0)
        v, div = model.apply(P, to_dev(x, cuda), to_dev(t, cuda), to_dev(z, cuda) if hutch else None, hutch=hutch, want_div=True)
        print(name, n, "  h16 off: field err", rel_err(v.cpu().numpy(), v_ref), "div err", np.abs(div.cpu().numpy() - div_ref).max() / max(np.abs(div_ref).max(), 1.0), flush=True)
        lib.mfm_set_gemm_h16(1)
        if os.environ.get("FIELD_ONLY"): continue
        args = SimpleNamespace(hutchs=hutch, num_importance_samples=0, mcmc_per_flow_steps=10, step_size=0.1)
        opts = SimpleNamespace(rtol=1e-5, atol=1e-5, mxstep=1000, n_times=n_times)
        gen, init_fn, push = E.create_train_data_gn(dd, model, opts, args)
        flow = OS.Flow(params, omega, ot, hutch, 1e-5, 1e-5, 1000, clip, np.linspace(0.0, 1.0, n_times), rng_dtype=np.float32)
        keys = tf.split(tf.PRNGKey(7), n)
        u = tf.vmap_normal(tf.split(tf.PRNGKey(8), n), ot.dim).astype(np.float64)
        x_ref, ldj_ref = flow.transform_and_logdet(keys, u)
        x32, ldj32 = flow.transform_and_logdet(keys, u.astype(np.float32))
        print("   oracle32 vs 64: x", np.abs(x32 - x_ref).max(), "ldj", np.abs(ldj32 - ldj_ref).max())
        for mode in (1, 0):
            lib.mfm_debug_set_ode_small(mode)
            stats = torch.zeros(8, dtype=torch.int32, device=cuda)
            xd, ldj = push(key_dev(keys, cuda), to_dev(u, cuda), P, stats)
            ex = np.abs(xd.cpu().numpy() - x_ref); el = np.abs(ldj.cpu().numpy() - ldj_ref)
            print("   mode", mode, "x err max/median", ex.max(), np.median(ex), "ldj err", el.max(), np.median(el), stats.cpu().tolist()[:4], flush=True)
        lib.mfm_debug_set_ode_small(1)
