"""Oracle: VectorFieldNet forward / JVP / exact divergence / FM loss + gradient.
TEST INFRASTRUCTURE ONLY.

Follows /root/reference/exe_flow_matching.py:56-90 (VectorFieldNet.__call__), :151-179
(cond_flow_fn, flow_matching_loss) and :206-242 (divergence by jvp or trace(jacfwd)).
flax.linen.Dense is restated as y = x @ kernel + bias with kernel [in, out]; modules are
auto-named Dense_0..Dense_7 in construction order: 0,1 = time branch, 2,3 = x branch,
4 = nn_t head, 5,6 = joint branch, 7 = nn_xt head.  Activation: `non_lins` of exe_flow_matching.py:40-46 (relu in every
configuration; tanh / elu / gelu / swish selectable with --non_linearity).  relu'(0) = 0 as in jax.nn.relu's custom JVP;
jax.nn.gelu defaults to the tanh approximation.
"""
from __future__ import annotations

import numpy as np

from . import threefry as tf


def init_params(rng: np.random.Generator, d, H, F=128, head_scale=0.0, dtype=np.float32):
    """NOT the reference initialiser (flax lecun_normal through flax's RNG folding is a 'next'
    row).  Gaussian fan-in init; head_scale=0 reproduces the reference's zero-init heads
    (exe_flow_matching.py:81,86); head_scale>0 gives the 'trained-like' bench fixture."""
    shapes = [(2 * F, H), (H, H), (d, H), (H, H), (H, d), (2 * H, H), (H, H), (H, d)]
    p = {}
    for i, (fi, fo) in enumerate(shapes):
        scale = 1.0 / np.sqrt(fi)
        if i in (4, 7):
            scale *= head_scale
        p[f"Dense_{i}"] = {
            "kernel": (rng.standard_normal((fi, fo)) * scale).astype(dtype),
            "bias": (rng.standard_normal(fo) * 0.01).astype(dtype),
        }
    return {"params": p}


_CAST = {}      # (id(array), dtype) -> (array, converted): the float32 -> float64 cast of 10 M weights per call dominated
                # the oracle's field evaluation; the source array is kept alive in the entry so ids cannot be recycled


def _cast(a, dt):
    if a.dtype == dt:
        return a
    k = (id(a), np.dtype(dt).str)
    hit = _CAST.get(k)
    if hit is None or hit[0] is not a:
        if len(_CAST) > 64:
            _CAST.clear()
        hit = _CAST[k] = (a, a.astype(dt))
    return hit[1]


def _wb(params, i, dt):
    q = params["params"][f"Dense_{i}"]
    return _cast(q["kernel"], dt), _cast(q["bias"], dt)


def activation(name, v):
    """(act(v), act'(v)) for jax.nn.{relu, tanh, elu, gelu (approximate=True), swish}."""
    dt = v.dtype
    if name == "relu":
        return np.maximum(v, 0), (v > 0).astype(dt)
    if name == "tanh":
        t = np.tanh(v); return t, 1 - t * t
    if name == "elu":
        e = np.expm1(v); return np.where(v > 0, v, e), np.where(v > 0, dt.type(1), e + 1)
    if name == "gelu":
        c, a = dt.type(np.sqrt(2 / np.pi)), dt.type(0.044715)
        t = np.tanh(c * (v + a * v ** 3))
        return dt.type(0.5) * v * (1 + t), dt.type(0.5) * (1 + t) + dt.type(0.5) * v * (1 - t * t) * c * (1 + 3 * a * v * v)
    if name == "swish":
        sg = 1 / (1 + np.exp(-v)); return v * sg, sg + v * sg * (1 - sg)
    raise ValueError(name)


def fourier_features(t, omega, dt):
    """exe_flow_matching.py:70-71.  t: [N] (per-chain time), omega: [F]."""
    deg = dt.type(2 * np.pi) * omega.astype(dt)[None, :] * t.astype(dt)[:, None]
    return np.concatenate([np.cos(deg), np.sin(deg)], axis=1)


def forward(params, omega, x, t, grad_logprob, grad_clip=None, want_cache=False, act="relu"):
    """v(x, t) batched: x [N,d], t [N] -> [N,d]."""
    dt = x.dtype
    ff = fourier_features(t, omega, dt)
    W0, b0 = _wb(params, 0, dt); W1, b1 = _wb(params, 1, dt)
    W2, b2 = _wb(params, 2, dt); W3, b3 = _wb(params, 3, dt)
    W4, b4 = _wb(params, 4, dt); W5, b5 = _wb(params, 5, dt)
    W6, b6 = _wb(params, 6, dt); W7, b7 = _wb(params, 7, dt)
    h0, dh0 = activation(act, ff @ W0 + b0)
    st, dst = activation(act, h0 @ W1 + b1)
    h2, dh2 = activation(act, x @ W2 + b2)
    sx, dsx = activation(act, h2 @ W3 + b3)
    gt = st @ W4 + b4
    cat = np.concatenate([sx, st], axis=1)
    h5, dh5 = activation(act, cat @ W5 + b5)
    h6, dh6 = activation(act, h5 @ W6 + b6)
    y = h6 @ W7 + b7
    g = grad_logprob(x)
    gc = np.clip(g, -grad_clip, grad_clip).astype(dt) if grad_clip else g
    v = y + gt * gc
    if want_cache:
        return v, dict(ff=ff, h0=h0, st=st, h2=h2, sx=sx, gt=gt, cat=cat, h5=h5, h6=h6, g=g, gc=gc,
                       dh0=dh0, dst=dst, dh2=dh2, dsx=dsx, dh5=dh5, dh6=dh6)
    return v


def jvp_x(params, cache, z, hvp, grad_clip=None):
    """d/dx v(x,t) . z for a batch of tangents z [N,d] given the forward cache."""
    dt = z.dtype
    H = cache["sx"].shape[1]
    W2, _ = _wb(params, 2, dt); W3, _ = _wb(params, 3, dt)
    W5, _ = _wb(params, 5, dt); W6, _ = _wb(params, 6, dt); W7, _ = _wb(params, 7, dt)
    d2 = (z @ W2) * cache["dh2"]
    d3 = (d2 @ W3) * cache["dsx"]
    d5 = (d3 @ W5[:H]) * cache["dh5"]
    d6 = (d5 @ W6) * cache["dh6"]
    dy = d6 @ W7
    dg = hvp(z)
    if grad_clip:
        dg = dg * ((cache["g"] > -grad_clip) & (cache["g"] < grad_clip))
    return dy + cache["gt"] * dg


def field_and_div(params, omega, x, t, target, hutch_z=None, grad_clip=None, act="relu"):
    """(v, div v) as the augmented field needs them (exe_flow_matching.py:208-218).
    hutch_z [N,d]: Hutchinson probe (div = z.(Jz)); None: exact trace by d tangents."""
    v, c = forward(params, omega, x, t, target.grad, grad_clip, want_cache=True, act=act)
    if hutch_z is not None:
        jz = jvp_x(params, c, hutch_z, lambda zz: target.hvp(x, zz), grad_clip)
        return v, (hutch_z * jz).sum(1)
    N, d = x.shape
    div = np.zeros(N, x.dtype)
    hd = target.hdiag(x)
    if grad_clip:
        hd = hd * ((c["g"] > -grad_clip) & (c["g"] < grad_clip))
    for j in range(d):
        e = np.zeros_like(x); e[:, j] = 1
        jz = jvp_x(params, c, e, lambda zz: np.zeros_like(zz), None)
        div += jz[:, j]
    return v, div + (c["gt"] * hd).sum(1)


def fm_batch(key, samples, ref_sampler, sigma, rng_dtype=None):
    """cond_flow_fn (exe_flow_matching.py:151-169), cond_flow=True, ot_cond_flow=False.
    rng_dtype: dtype of the random draws (float32 = x64 off); arithmetic in samples.dtype."""
    N, d = samples.shape
    dt = samples.dtype
    rdt = np.dtype(rng_dtype or dt)
    key_time, key_ref, key_gauss, _ = tf.split(key, 4)
    times = tf.uniform(key_time, (N, 1), rdt).astype(dt)
    ref = ref_sampler(tf.split(key_ref, N), rdt).astype(dt)
    eps = tf.normal(key_gauss, (N, d), rdt).astype(dt)
    xt = dt.type(sigma) * eps + times * samples + (dt.type(1) - times) * ref
    target = samples - ref
    return times[:, 0], xt, target


def fm_batch_uncond(key, samples, sigma, rng_dtype=None):
    """flow_fn (exe_flow_matching.py:139-147), the --cond_flow-off variant: two-way key split, ONE [N,d] normal draw."""
    N, d = samples.shape
    dt = samples.dtype
    rdt = np.dtype(rng_dtype or dt)
    key_time, key_ref = tf.split(key)
    times = tf.uniform(key_time, (N, 1), rdt).astype(dt)
    ref = tf.normal(key_ref, (N, d), rdt).astype(dt)
    sds = dt.type(1.0) - (dt.type(1.0) - dt.type(sigma)) * times
    xt = times * samples + sds * ref
    target = samples - (dt.type(1) - dt.type(sigma)) * ref
    return times[:, 0], xt, target


def fm_loss_and_grad(params, omega, xt, times, target_v, grad_logprob, grad_clip=None, act="relu"):
    """flow_matching_loss (exe_flow_matching.py:171-179) and its gradient w.r.t. params
    (what jax.value_and_grad(loss_fn, argnums=2) returns, :364-365)."""
    dt = xt.dtype
    v, c = forward(params, omega, xt, times, grad_logprob, grad_clip, want_cache=True, act=act)
    diff = v - target_v
    loss = (diff * diff).sum()
    H = c["sx"].shape[1]
    W = [_wb(params, i, dt)[0] for i in range(8)]
    delta = dt.type(2) * diff
    G = {}

    def put(i, a, dz):
        G[f"Dense_{i}"] = {"kernel": a.T @ dz, "bias": dz.sum(0)}

    put(7, c["h6"], delta)
    d6 = (delta @ W[7].T) * c["dh6"]
    put(6, c["h5"], d6)
    d5 = (d6 @ W[6].T) * c["dh5"]
    put(5, c["cat"], d5)
    dcat = d5 @ W[5].T
    dgt = delta * c["gc"]
    put(4, c["st"], dgt)
    dsx = dcat[:, :H] * c["dsx"]
    dst = (dcat[:, H:] + dgt @ W[4].T) * c["dst"]
    put(3, c["h2"], dsx)
    d2 = (dsx @ W[3].T) * c["dh2"]
    put(2, xt, d2)
    put(1, c["h0"], dst)
    d0 = (dst @ W[1].T) * c["dh0"]
    put(0, c["ff"], d0)
    return loss, {"params": G}
