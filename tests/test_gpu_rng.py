"""CUDA threefry / uniform / normal vs the oracle: bits and uniforms bit-exact, normals <= 4 ulp
(erf_inv goes through log1pf/sqrtf whose last-bit behaviour differs between libm and CUDA)."""
import numpy as np
import pytest
import torch

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
def test_normal_within_4ulp(cuda, lib, n):
    from mfm_b200 import random as mr
    k = tf.PRNGKey(1234)
    got = mr.normal(key_dev(k, cuda), (n,)).cpu().numpy()
    exp = tf.normal(k, (n,))
    assert ulp_diff_f32(got, exp).max() <= 4
    assert np.isfinite(got).all()


def test_normal_batched_is_vmap(cuda, lib):
    from mfm_b200 import random as mr
    keys = tf.split(tf.PRNGKey(3), 17)
    for d in (2, 7, 64):
        got = mr.normal(key_dev(keys, cuda), (d,)).cpu().numpy()
        exp = tf.vmap_normal(keys, d)
        assert ulp_diff_f32(got, exp).max() <= 4
        got = mr.uniform(key_dev(keys, cuda), (d,)).cpu().numpy()
        exp = np.stack([tf.uniform(k, (d,)) for k in keys])
        assert got.tobytes() == exp.tobytes()
