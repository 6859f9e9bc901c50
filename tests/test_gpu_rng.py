"""CUDA threefry / uniform / normal vs the oracle: bits, uniforms and normals bit-exact (erf_inv restated operation for operation,
the logarithm correctly rounded on both sides)."""
import numpy as np
import pytest
import torch

from mfm_b200 import _lib
from oracle import threefry as tf
from tests.helpers import key_dev, ulp_diff_f32

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,num", [(0, 2), (1, 3), (1024, 6), (59049, 129), (7, 1)])
def test_split_bit_exact(cuda, lib, seed, num):
    from mfm_b200 import random as mr
    got = mr.split(mr.PRNGKey(seed, cuda), num).cpu().numpy()
    assert got.tolist() == tf.split(tf.PRNGKey(seed), num).tolist()


def test_split_batched_bit_exact(cuda, lib):
    from mfm_b200 import random as mr
    keys = tf.split(tf.PRNGKey(5), 9)
    got = mr.split(key_dev(keys, cuda), 4).cpu().numpy()
    exp = np.stack([tf.split(k, 4) for k in keys])
    assert got.tolist() == exp.tolist()


@pytest.mark.parametrize("n", [1, 2, 5, 64, 1601, 100000])
def test_bits_and_uniform_bit_exact(cuda, lib, n):
    from mfm_b200 import random as mr
    k = tf.PRNGKey(11)
    assert mr.bits(key_dev(k, cuda), (n,)).cpu().numpy().tolist() == tf.random_bits(k, 32, (n,)).tolist()
    got = mr.uniform(key_dev(k, cuda), (n,)).cpu().numpy()
    assert got.tobytes() == tf.uniform(k, (n,)).tobytes()
    got = mr.uniform(key_dev(k, cuda), (n,), -12.8, 12.8).cpu().numpy()
    assert got.tobytes() == tf.uniform(k, (n,), np.float32, -12.8, 12.8).tobytes()


def test_scalar_uniform_and_doc_value(cuda, lib):
    from mfm_b200 import random as mr
    assert mr.uniform(mr.PRNGKey(0, cuda), ()).item() == np.float32(0.41845703)
    assert mr.normal(mr.PRNGKey(0, cuda), (1,)).item() == pytest.approx(-0.20584226, abs=1e-7)


@pytest.mark.parametrize("n", [1, 2, 3, 64, 1600, 200001])
def test_normal_is_bit_exact(cuda, lib, n):
    """jax.random.normal float32: the device follows the oracle's restatement of XLA's ErfInv operation for operation (separately
    rounded multiplies and adds, XLA's log1p split, the logarithm correctly rounded on both sides), so the streams agree bit for
    bit - 4 ulp in round 1, when the compiler contracted the polynomial to FMAs and the two log1p implementations differed."""
    from mfm_b200 import random as mr
    k = tf.PRNGKey(1234)
    got = mr.normal(key_dev(k, cuda), (n,)).cpu().numpy()
    exp = tf.normal(k, (n,))
    assert got.tobytes() == exp.tobytes(), int(ulp_diff_f32(got, exp).max())
    assert np.isfinite(got).all()


def test_table_driven_log_gives_the_same_normals(cuda, lib):
    """The hot kernels (pines_propose_kernel, fm_batch_kernel, mala_small_kernel) take the logarithm inside erf_inv from a
    16-entry table + degree-9 series with a Ziv rounding test (csrc/common.cuh::log_rn_f32) instead of the generic double log:
    the resulting normal must be the same float for EVERY 32-bit input (2^32 inputs in 16 interleaved sweeps of 2^28)."""
    import ctypes
    fn = lib.mfm_debug_normal_fast_check
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_ulonglong, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
    out = torch.zeros(2, dtype=torch.int64, device=cuda)
    total = 0
    for start in range(16):
        _lib.check(fn(1 << 28, start, 16, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        total += int(out[0].item())
    assert total == 0, total


@pytest.mark.parametrize("n", [1, 2, 3, 64, 1601, 100000])
def test_x64_draw_layout(cuda, lib, n):
    """jax_enable_x64 (the reference as shipped, multi_modal.py:14): 64 random bits per draw, float64 transform, rounded to
    float32 once.  Uniforms are exact; normals go through two different float64 erfinv implementations (CUDA / scipy standing in
    for XLA's), a few float64 ulps apart - invisible after the rounding except on ties."""
    from mfm_b200 import random as mr
    k = tf.PRNGKey(99)
    lib.mfm_set_rng_x64(1)
    try:
        u = mr.uniform(key_dev(k, cuda), (n,)).cpu().numpy()
        u2 = mr.uniform(key_dev(k, cuda), (n,), -12.8, 12.8).cpu().numpy()
        z = mr.normal(key_dev(k, cuda), (n,)).cpu().numpy()
        keys = tf.split(tf.PRNGKey(3), 5)
        zb = mr.normal(key_dev(keys, cuda), (7,)).cpu().numpy()
    finally:
        lib.mfm_set_rng_x64(0)
    assert u.tobytes() == tf.uniform(k, (n,), np.float64).astype(np.float32).tobytes()
    assert u2.tobytes() == tf.uniform(k, (n,), np.float64, -12.8, 12.8).astype(np.float32).tobytes()
    exp = tf.normal(k, (n,), np.float64).astype(np.float32)
    d = ulp_diff_f32(z, exp)
    assert d.max() <= 1 and (d == 0).mean() >= 0.999
    expb = np.stack([tf.normal(kk, (7,), np.float64) for kk in keys]).astype(np.float32)
    assert ulp_diff_f32(zb, expb).max() <= 1
    # and the float32 layout is back
    assert mr.uniform(mr.PRNGKey(0, cuda), ()).item() == np.float32(0.41845703)


def test_normal_batched_is_vmap(cuda, lib):
    from mfm_b200 import random as mr
    keys = tf.split(tf.PRNGKey(3), 17)
    for d in (2, 7, 64):
        got = mr.normal(key_dev(keys, cuda), (d,)).cpu().numpy()
        exp = tf.vmap_normal(keys, d)
        assert got.tobytes() == exp.tobytes()
        got = mr.uniform(key_dev(keys, cuda), (d,)).cpu().numpy()
        exp = np.stack([tf.uniform(k, (d,)) for k in keys])
        assert got.tobytes() == exp.tobytes()
