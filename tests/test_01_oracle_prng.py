"""Pins the RNG contract on published known-answer values (the only external anchors available:
the reference has no tests)."""
import numpy as np
from oracle import prng


def _hex(t):
    return tuple(hex(int(v)) for v in t)


def test_threefry2x32_random123_kat():
    # Random123 KAT vectors; the same three are used by jax's own random_test.py::testThreefry2x32
    assert _hex(prng.threefry2x32(0, 0, 0, 0)) == ("0x6b200159", "0x99ba4efe")
    m = 0xFFFFFFFF
    assert _hex(prng.threefry2x32(m, m, m, m)) == ("0x1cb996fc", "0xbb002be7")
    assert _hex(prng.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3)) == ("0xc4923a9c", "0x483df7a0")


def test_split_matches_jax_docs():
    # values printed in the public JAX documentation ("Pseudo random numbers" / jax.random)
    np.testing.assert_array_equal(prng.split(prng.PRNGKey(0)),
                                  np.array([[4146024105, 967050713], [2718843009, 1272950319]], np.uint32))
    np.testing.assert_array_equal(prng.split(prng.PRNGKey(42)),
                                  np.array([[2465931498, 3679230171], [255383827, 267815257]], np.uint32))


def test_uniform_matches_jax_docs():
    assert float(prng.uniform(prng.PRNGKey(0))) == np.float32(0.41845703)


def test_split_is_batched_consistently():
    keys = prng.split(prng.PRNGKey(5), 7)
    batched = prng.split(keys, 3)
    for i in range(7):
        np.testing.assert_array_equal(batched[i], prng.split(keys[i], 3))
    u = prng.uniform(keys, (5,))
    for i in range(7):
        np.testing.assert_array_equal(u[i], prng.uniform(keys[i], (5,)))


def test_choice_p_is_inverse_cdf():
    keys = prng.split(prng.PRNGKey(1), 20000)
    p = np.array([0.1, 0.2, 0.3, 0.25, 0.15], np.float32)
    a = prng.choice_p(keys, np.broadcast_to(p, (20000, 5)))
    freq = np.bincount(a, minlength=5) / 20000
    assert np.abs(freq - p).max() < 0.015
    assert a.min() >= 0 and a.max() <= 4


def test_randint_and_shuffle_ranges():
    keys = prng.split(prng.PRNGKey(2), 500)
    r = prng.randint(keys, (4,), 3, 9)
    assert r.min() >= 3 and r.max() < 9 and len(np.unique(r)) == 6
    s = prng.shuffle(keys, 50)
    assert (np.sort(s, -1) == np.arange(50)).all()
    t = prng.choice_no_replace_p(keys, (np.arange(50) % 3 != 0).astype(np.float32)[None].repeat(500, 0), 6)
    assert ((t % 3) != 0).all() and all(len(set(row)) == 6 for row in t)


def test_product_host_prng_matches_oracle():
    from to_ued_b200.util import prng as P
    k = P.PRNGKey(7)
    ks = P.split(k, 5)
    np.testing.assert_array_equal(ks, prng.split(prng.PRNGKey(7), 5))
    np.testing.assert_array_equal(P.uniform(ks, (7,), -1, 1), prng.uniform(ks, (7,), -1, 1))
    np.testing.assert_array_equal(P.randint(ks, (3,), 0, 7), prng.randint(ks, (3,), 0, 7))
    np.testing.assert_array_equal(P.shuffle_prefix(ks, 100, 15), prng.choice_no_replace_uniform(ks, 100, 15))
    m = np.random.RandomState(0).rand(5, 100) > 0.3
    np.testing.assert_array_equal(P.masked_topk(ks, m, 6), prng.choice_no_replace_p(ks, m.astype(np.float32), 6))
