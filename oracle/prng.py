"""Counter-based RNG contract (oracle side).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates the published threefry2x32 block function (Salmon et al., "Parallel random numbers: as
easy as 1, 2, 3", Random123) and the way jax==0.4.13 (setup/requirements-cpu.txt:1) derives
``split`` / ``random_bits`` / ``uniform`` / ``bernoulli`` / ``choice`` / ``permutation`` from it
(non-partitionable threefry, the 0.4.13 default).  [3P-recall]: jax is not installable here; the
block function is pinned on the Random123 KATs and ``split`` / ``uniform`` on values printed in
the public JAX docs (tests/test_01_oracle_prng.py).

All functions are vectorised over leading "batch" axes of the key: a key is a ``uint32[..., 2]``
array.  The CUDA kernels (to_ued_b200/csrc/prng.cuh) implement exactly the same derivation, so
every draw on the hot path can be compared bit for bit.

Reference call sites: environments/rollout.py:40,49,61-64; environments/gridworld/gridworld.py:76,
88,116,161; environments/level_sampler.py:159,214,224,341,375,379,401;
environments/gridworld/configs.py:49,112,119.
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_ROT0 = (13, 15, 26, 6)
_ROT1 = (17, 29, 16, 24)
_PARITY = U32(0x1BD11BDA)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32, 20 rounds.  All arguments broadcastable uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            rot = _ROT0 if i % 2 == 0 else _ROT1
            for r in rot:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + U32(i + 1)
    return x0, x1


def PRNGKey(seed: int) -> np.ndarray:
    """jax.random.PRNGKey: (hi32, lo32) of the 64-bit seed."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=U32)


def _bits_flat(key, n: int):
    """threefry_2x32(key, iota(n)) of jax 0.4.13: counts are split into two halves
    (odd n padded with one 0), hashed pairwise, and the two output halves concatenated.
    key: uint32[..., 2]  ->  uint32[..., n]"""
    key = np.asarray(key, dtype=U32)
    cnt = np.arange(n, dtype=U32)
    if n % 2:
        cnt = np.concatenate([cnt, np.zeros(1, dtype=U32)])
    half = cnt.shape[0] // 2
    o0, o1 = threefry2x32(key[..., 0:1], key[..., 1:2], cnt[:half], cnt[half:])
    out = np.concatenate([o0, o1], axis=-1)
    return out[..., :n]


def split(key, num: int = 2):
    """jax.random.split: uint32[..., 2] -> uint32[..., num, 2]"""
    flat = _bits_flat(key, 2 * num)
    return flat.reshape(flat.shape[:-1] + (num, 2))


def random_bits(key, shape=()):
    """32-bit random bits of trailing shape ``shape`` per key."""
    shape = tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    flat = _bits_flat(key, n)
    return flat.reshape(flat.shape[:-1] + shape)


def bits_to_unit_float(bits):
    """jax _uniform: mantissa trick, result in [0, 1)."""
    fb = (np.asarray(bits, dtype=U32) >> U32(9)) | U32(0x3F800000)
    return fb.view(np.float32) - np.float32(1.0)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    f = bits_to_unit_float(random_bits(key, shape))
    minval = np.float32(minval)
    maxval = np.float32(maxval)
    return np.maximum(minval, f * (maxval - minval) + minval).astype(np.float32)


def bernoulli(key, p, shape=None):
    p = np.asarray(p, dtype=np.float32)
    if shape is None:
        shape = p.shape[np.asarray(key).ndim - 1:] if p.ndim >= np.asarray(key).ndim else ()
    return uniform(key, shape) < p


def cumsum_seq(p):
    """Left-to-right float32 running sum over the last axis (the contract; jnp.cumsum's
    association order on a given backend is not knowable here)."""
    p = np.asarray(p, dtype=np.float32)
    out = np.empty_like(p)
    acc = np.zeros(p.shape[:-1], dtype=np.float32)
    for i in range(p.shape[-1]):
        acc = (acc + p[..., i]).astype(np.float32)
        out[..., i] = acc
    return out


def choice_p(key, p):
    """jax.random.choice(key, n, p=p), scalar draw, replace=True:
    r = cumsum(p)[-1] * (1 - uniform(key)); searchsorted(cumsum(p), r, side='left')."""
    pc = cumsum_seq(p)
    u = uniform(key, ())
    r = (pc[..., -1] * (np.float32(1.0) - u)).astype(np.float32)
    return (pc < r[..., None]).sum(axis=-1).astype(np.int32)


def choice_p_many(key, p, n: int):
    """jax.random.choice(key, a, shape=(n,), p=p, replace=True)"""
    pc = cumsum_seq(p)
    u = uniform(key, (n,))
    r = (pc[..., -1:] * (np.float32(1.0) - u)).astype(np.float32)
    return (pc[..., None, :] < r[..., :, None]).sum(axis=-1).astype(np.int32)


def randint(key, shape, minval: int, maxval: int):
    """jax.random.randint for int32 (jax 0.4.13 _randint): two bit draws combined with a
    multiplier so the modulo bias is negligible."""
    shape = tuple(shape)
    k = split(key, 2)
    hi = random_bits(k[..., 0, :], shape).astype(np.uint64)
    lo = random_bits(k[..., 1, :], shape).astype(np.uint64)
    span = np.uint64((int(maxval) - int(minval)) & 0xFFFFFFFF)
    # multiplier = (2**32 % span)**2 % span, all in uint32 arithmetic
    m = (np.uint64(1 << 32) % span) if span else np.uint64(0)
    m = (m * m) & np.uint64(0xFFFFFFFF)
    m = m % span
    off = ((hi % span) * m) & np.uint64(0xFFFFFFFF)
    off = (off + (lo % span)) & np.uint64(0xFFFFFFFF)
    off = off % span
    return (np.int64(minval) + off.astype(np.int64)).astype(np.int32)


def choice_uniform(key, n: int):
    """jax.random.choice(key, a) with no p, scalar draw: a[randint(key, (), 0, n)]"""
    return randint(key, (), 0, n)


def shuffle(key, n: int):
    """jax.random.permutation(key, n) == _shuffle(key, arange(n)): ceil(3 ln n / ln(2^32-1))
    rounds of a stable sort by fresh 32-bit keys."""
    key = np.asarray(key, dtype=U32)
    x = np.broadcast_to(np.arange(n, dtype=np.int32), key.shape[:-1] + (n,)).copy()
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(float(np.iinfo(np.uint32).max))))
    for _ in range(rounds):
        ks = split(key, 2)
        key, sub = ks[..., 0, :], ks[..., 1, :]
        sk = random_bits(sub, (n,))
        order = np.argsort(sk, axis=-1, kind="stable")
        x = np.take_along_axis(x, order, axis=-1)
    return x


def choice_no_replace_uniform(key, n: int, k: int):
    """jax.random.choice(key, arange(n), shape=(k,), replace=False) with p=None."""
    return shuffle(key, n)[..., :k]


def choice_no_replace_p(key, p, k: int):
    """jax.random.choice(key, arange(n), shape=(k,), replace=False, p=p): Gumbel top-k,
    ``argsort(-gumbel(key, n) - log(p))[:k]``.

    Contract note: g_i = log(-log(u_i)) - log(p_i) with u_i = uniform(key, minval=tiny, maxval=1).
    When every admissible p_i is equal (the only way the reference calls it: p is a 0/1 mask,
    configs.py:49 and level_sampler.py:401), the order of g is the order of *decreasing u*
    with p_i = 0 entries last, so the contract ranks on the raw uniform bits (exact, no
    transcendental) and breaks ties by index like a stable argsort."""
    p = np.asarray(p, dtype=np.float32)
    n = p.shape[-1]
    bits = random_bits(key, (n,))
    # larger u  <=>  smaller g ; stable: ties keep index order.  p==0 -> g=+inf -> last.
    mant = (bits >> U32(9)).astype(np.int64)
    sort_key = np.where(p > 0, -mant, np.int64(1 << 40))
    order = np.argsort(sort_key, axis=-1, kind="stable")
    return order[..., :k].astype(np.int32)
