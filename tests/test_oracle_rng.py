"""Pins the RNG oracle: Random123 KATs + values printed in JAX's public documentation."""
import numpy as np

from oracle import threefry as tf


def test_threefry2x32_random123_kats():
    kats = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
            ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
            ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]
    for key, ctr, exp in kats:
        o = tf.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(o[0]), int(o[1])) == exp


def test_split_layout_matches_jax_docs():
    k = tf.PRNGKey(0)
    assert tf.split(k, 2).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]


def test_uniform_matches_jax_docs():
    assert tf.uniform(tf.PRNGKey(0), ()) == np.float32(0.41845703)


def test_normal_matches_jax_sharp_bits_notebook():
    # "JAX - The Sharp Bits", section on random numbers (jax 0.4.x, threefry, x64 off)
    key = tf.PRNGKey(0)
    assert tf.normal(key, (1,))[0] == np.float32(-0.20584226)
    key, sub = tf.split(key)
    assert tf.normal(sub, (1,))[0] == np.float32(-1.2515389)
    key, sub = tf.split(key)
    assert tf.normal(sub, (1,))[0] == np.float32(-0.58665055)
    key, *subs = tf.split(key, 4)
    got = [tf.normal(s, (1,))[0] for s in subs]
    assert got == [np.float32(-0.37533438), np.float32(0.98645043), np.float32(0.14553197)]


def test_odd_sizes_and_64bit_layout():
    k = tf.PRNGKey(7)
    b5 = tf.random_bits(k, 32, (5,))
    # padded element is count 0, output truncated
    o0, o1 = tf.threefry2x32(k[0], k[1], np.array([0, 1, 2], np.uint32), np.array([3, 4, 0], np.uint32))
    assert b5.tolist() == np.concatenate([o0, o1])[:5].tolist()
    b64 = tf.random_bits(k, 64, (3,))
    w = tf.random_bits(k, 32, (6,))
    assert b64.tolist() == [(int(w[i]) << 32) | int(w[3 + i]) for i in range(3)]
    u64 = tf.uniform(k, (4,), np.float64)
    assert u64.dtype == np.float64 and ((u64 >= 0) & (u64 < 1)).all()


def test_normal_distribution_sanity():
    x = tf.normal(tf.PRNGKey(3), (200000,))
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1) < 0.01
    assert np.isfinite(x).all()
