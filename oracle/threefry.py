"""Oracle: jax.random (legacy uint32[2] keys, threefry2x32, ``jax_threefry_partitionable=False``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates jax 0.4.26 ``jax/_src/prng.py`` (``threefry_2x32``, ``_threefry_split``,
``_threefry_random_bits_original``) and ``jax/_src/random.py`` (``_uniform``, ``_normal_real``,
``_bernoulli``) — third-party code that is NOT under /root/reference (pinned by
/root/reference/environment.yaml:101-102).  Call sites in the reference:
bblackjax/util.py:81, bblackjax/mcmc/proposal.py:179, exe_flow_matching.py:142-166,212,232,
247,257,265,268,275,303,333,433, distributions.py:70-76,93-97,163-164,313-314.

Pinned by: Random123 KATs for threefry2x32-20, and public JAX doc values
(split(PRNGKey(0)), uniform(PRNGKey(0)), normal(PRNGKey(0),(1,)) = -0.20584226 ...).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """20-round Threefry-2x32.  Inputs broadcastable uint32 arrays; returns (o0, o1)."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, U32)
        k1 = np.asarray(k1, U32)
        x0 = np.asarray(x0, U32).copy()
        x1 = np.asarray(x1, U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
        return x0.astype(U32), x1.astype(U32)


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey with x64 off keeps the low 32 bits; with x64 on, hi=seed>>32."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _threefry_2x32_counts(key, counts):
    """prng.threefry_2x32(key, count): halves layout, pad one zero if odd."""
    counts = np.asarray(counts, U32).ravel()
    n = counts.size
    odd = n % 2
    if odd:
        counts = np.concatenate([counts, np.zeros(1, U32)])
    half = counts.size // 2
    o0, o1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def random_bits(key, bit_width, shape):
    """_threefry_random_bits_original for bit_width in (32, 64)."""
    shape = tuple(shape)
    size = int(np.prod(shape)) if shape else 1
    max_count, r = divmod(bit_width * size, 32)
    if r:
        max_count += 1
    bits = _threefry_2x32_counts(key, np.arange(max_count, dtype=U32))
    if bit_width == 64:
        hi, lo = np.split(bits.astype(np.uint64), 2)
        bits = (hi << np.uint64(32)) | lo
    elif bit_width != 32:
        raise NotImplementedError(bit_width)
    return bits.reshape(shape)


def split(key, num=2):
    """jax.random.split -> uint32[num, 2]."""
    key = np.asarray(key, U32)
    return _threefry_2x32_counts(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    """jax.random.uniform: mantissa-fill, subtract 1, affine, clamp below by minval."""
    dtype = np.dtype(dtype)
    key = np.asarray(key, U32)
    if dtype == np.float32:
        bits = random_bits(key, 32, shape)
        fb = (bits >> U32(9)) | U32(0x3F800000)
        floats = fb.view(np.float32) - np.float32(1.0)
    elif dtype == np.float64:
        bits = random_bits(key, 64, shape)
        fb = (bits >> np.uint64(12)) | np.uint64(0x3FF0000000000000)
        floats = fb.view(np.float64) - np.float64(1.0)
    else:
        raise NotImplementedError(dtype)
    minval = dtype.type(minval)
    maxval = dtype.type(maxval)
    out = floats * dtype.type(maxval - minval) + minval
    return np.maximum(minval, out).astype(dtype).reshape(shape)


# XLA's ErfInv for f32 (xla/client/lib/math.cc, Giles 2010 single-precision polynomial).
_W_LT5 = np.array([2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
                   0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941],
                  dtype=np.float32)
_W_GE5 = np.array([-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
                   0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682],
                  dtype=np.float32)


def log1p_f32_xla(t):
    """XLA CPU's float32 log1p (elemental_ir_emitter.cc EmitLog1p; third-party, restated - unpinned): (-0.5 t + 1) t for
    |t| < 1e-4, else log(1 + t) with the sum rounded to float32 first.  The logarithm is evaluated in float64 and rounded
    (= correctly rounded logf; XLA's vectorised polynomial is within an ulp of it) so that the CUDA path can reproduce it
    bit for bit."""
    t = np.asarray(t, np.float32)
    small = ((np.float32(-0.5) * t).astype(np.float32) + np.float32(1.0)).astype(np.float32) * t
    with np.errstate(divide="ignore", invalid="ignore"):
        large = np.log((np.float32(1.0) + t).astype(np.float32).astype(np.float64)).astype(np.float32)
    return np.where(np.abs(t) < np.float32(1e-4), small.astype(np.float32), large).astype(np.float32)


def erf_inv_f32(x):
    """every multiply and add rounds separately (XLA CPU does not contract to FMA)"""
    x = np.asarray(x, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -log1p_f32_xla(-(x * x).astype(np.float32))
        lt = w < np.float32(5.0)
        w = np.where(lt, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(lt, _W_LT5[0], _W_GE5[0]).astype(np.float32)
        for i in range(1, 9):
            p = (np.where(lt, _W_LT5[i], _W_GE5[i]).astype(np.float32) + (p * w).astype(np.float32)).astype(np.float32)
        res = (p * x).astype(np.float32)
        return np.where(np.abs(x) == 1, x * np.float32(np.inf), res).astype(np.float32)


def erf_inv_f64(x):
    from scipy.special import erfinv  # f64 path: XLA uses a rational approx; scipy is within ~1ulp
    return erfinv(np.asarray(x, np.float64))


def normal(key, shape=(), dtype=np.float32):
    """jax.random.normal: sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1))."""
    dtype = np.dtype(dtype)
    lo = np.nextafter(dtype.type(-1.0), dtype.type(0.0))
    u = uniform(key, shape, dtype, lo, dtype.type(1.0))
    if dtype == np.float32:
        return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)
    return np.float64(np.sqrt(2)) * erf_inv_f64(u)


def bernoulli(key, p):
    """jax.random.bernoulli(key, p) for scalar/array p: uniform(key, shape(p), dtype(p)) < p."""
    p = np.asarray(p)
    return uniform(key, p.shape, p.dtype) < p


def vmap_normal(keys, d, dtype=np.float32):
    """vmap(lambda k: normal(k, (d,)))(keys) -> [n, d]."""
    return np.stack([normal(k, (d,), dtype) for k in np.asarray(keys, U32)])
