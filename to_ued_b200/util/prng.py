"""Host-side key handling for the B200 path: threefry2x32 keys compatible with the reference's
``jax.random`` usage (jax==0.4.13 derivation rules, see to_ued_b200/csrc/common.cuh for the
device twin).  Only key *plumbing* happens on the host (PRNGKey / split / a few scalar draws for
the level sampler); every per-environment draw happens inside the CUDA kernels.

Keys are ``numpy.uint32[..., 2]`` arrays, exactly like ``jax.random.PRNGKey`` raw keys."""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFF)
_ROUNDS = ((13, 15, 26, 6), (17, 29, 16, 24))


def _tf(k0, k1, c0, c1):
    """threefry2x32 on uint64 lanes masked to 32 bits (vectorised, broadcasting)."""
    k0 = np.asarray(k0, np.uint64); k1 = np.asarray(k1, np.uint64)
    x0 = np.asarray(c0, np.uint64); x1 = np.asarray(c1, np.uint64)
    sched = (k0, k1, (k0 ^ k1 ^ np.uint64(0x1BD11BDA)) & _M)
    x0 = (x0 + sched[0]) & _M
    x1 = (x1 + sched[1]) & _M
    for g in range(5):
        for r in _ROUNDS[g & 1]:
            x0 = (x0 + x1) & _M
            x1 = ((x1 << np.uint64(r)) | (x1 >> np.uint64(32 - r))) & _M
            x1 ^= x0
        x0 = (x0 + sched[(g + 1) % 3]) & _M
        x1 = (x1 + sched[(g + 2) % 3] + np.uint64(g + 1)) & _M
    return x0.astype(np.uint32), x1.astype(np.uint32)


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed) % (1 << 64)
    return np.array([seed >> 32, seed & 0xFFFFFFFF], np.uint32)


_NATIVE = [None]


def _native_iota_bits():
    """toued_host_iota_bits of libtoued.so (a plain C loop on the host), or False if the library is not built."""
    if _NATIVE[0] is None:
        try:
            from .. import _lib
            _NATIVE[0] = _lib.lib().toued_host_iota_bits
        except Exception:
            _NATIVE[0] = False
    return _NATIVE[0]


def iota_bits(key, n: int) -> np.ndarray:
    """threefry_2x32(key, iota(n)): uint32[..., 2] -> uint32[..., n]"""
    key = np.asarray(key, np.uint32)
    fn = _native_iota_bits() if (key.size >= 8 or n >= 64) else False   # batched keys / long draws: native host loop
    if fn:
        k2 = np.ascontiguousarray(key.reshape(-1, 2))
        out = np.empty((k2.shape[0], n), np.uint32)
        if fn(k2.ctypes.data, k2.shape[0], int(n), out.ctypes.data) != 0:
            raise RuntimeError("toued_host_iota_bits failed")
        return out.reshape(key.shape[:-1] + (n,))
    m = n + (n & 1)
    c = np.arange(m, dtype=np.uint32)
    if n & 1:
        c[-1] = 0
    h = m // 2
    a, b = _tf(key[..., :1], key[..., 1:], c[:h], c[h:])
    return np.concatenate([a, b], -1)[..., :n]


def _tf_scalar(k0: int, k1: int, x0: int, x1: int):
    """threefry2x32 on Python ints (the single-key fast path: ~10x less overhead than numpy on 2-element arrays)."""
    M = 0xFFFFFFFF
    sched = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
    x0 = (x0 + k0) & M
    x1 = (x1 + k1) & M
    for g in range(5):
        for r in _ROUNDS[g & 1]:
            x0 = (x0 + x1) & M
            x1 = ((x1 << r) | (x1 >> (32 - r))) & M
            x1 ^= x0
        x0 = (x0 + sched[(g + 1) % 3]) & M
        x1 = (x1 + sched[(g + 2) % 3] + g + 1) & M
    return x0, x1


def split(key, num: int = 2) -> np.ndarray:
    key = np.asarray(key, np.uint32)
    if key.ndim == 1 and num <= 4:                # the per-step host keys: one key -> a few keys
        k0, k1 = int(key[0]), int(key[1])
        pairs = [_tf_scalar(k0, k1, i, num + i) for i in range(num)]       # counts iota(2 num) split into halves
        flat = [p[0] for p in pairs] + [p[1] for p in pairs]
        return np.array(flat, np.uint32).reshape(num, 2)
    f = iota_bits(key, 2 * num)
    return f.reshape(f.shape[:-1] + (num, 2))


def unit_float(bits) -> np.ndarray:
    return ((np.asarray(bits, np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    shape = tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    f = unit_float(iota_bits(key, n))
    f = f.reshape(f.shape[:-1] + shape)
    lo, hi = np.float32(minval), np.float32(maxval)
    return np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)


def bits(key, shape=()) -> np.ndarray:
    shape = tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    f = iota_bits(key, n)
    return f.reshape(f.shape[:-1] + shape)


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """jax.random.randint (int32): hi/lo draws from split(key), (hi % span * mult + lo % span) % span"""
    ks = split(key, 2)
    hi = bits(ks[..., 0, :], shape).astype(np.uint64)
    lo = bits(ks[..., 1, :], shape).astype(np.uint64)
    span = np.uint64(int(maxval) - int(minval))
    mult = (np.uint64(1 << 32) % span)
    mult = (mult * mult & _M) % span
    off = (((hi % span) * mult & _M) + lo % span & _M) % span
    return (off.astype(np.int64) + int(minval)).astype(np.int32)


def shuffle_prefix(key, n: int, k: int) -> np.ndarray:
    """permutation(key, n)[:k] (jax _shuffle: stable sort by fresh 32-bit keys, ceil(3 ln n / ln 2^32) rounds)"""
    key = np.asarray(key, np.uint32)
    x = np.broadcast_to(np.arange(n, dtype=np.int32), key.shape[:-1] + (n,)).copy()
    rounds = int(np.ceil(3 * np.log(max(1, n)) / np.log(float(0xFFFFFFFF))))
    for _ in range(rounds):
        ks = split(key, 2)
        key = ks[..., 0, :]
        order = np.argsort(bits(ks[..., 1, :], (n,)), axis=-1, kind="stable")
        x = np.take_along_axis(x, order, -1)
    return x[..., :k]


def masked_topk(key, mask, k: int) -> np.ndarray:
    """choice(key, arange(n), (k,), replace=False, p=mask) for a 0/1 mask: Gumbel top-k reduces to
    ranking the admissible entries by decreasing uniform draw (ties by index)."""
    mask = np.asarray(mask, bool)
    n = mask.shape[-1]
    mant = (bits(key, (n,)) >> np.uint32(9)).astype(np.int64)
    rank_key = np.where(mask, -mant, np.int64(1) << 40)
    return np.argsort(rank_key, axis=-1, kind="stable")[..., :k].astype(np.int32)


# ---- device-side derivation (csrc/prng.cu): same bits, no host threefry / H2D in front of a launch ----
def to_device(key, device):
    """uint32 key(s) (numpy [.., 2] or tensor) -> int32 cuda tensor [n, 2]."""
    import torch
    if isinstance(key, torch.Tensor):
        return key.to(device).contiguous().view(torch.int32).reshape(-1, 2)
    a = np.ascontiguousarray(np.asarray(key, np.uint32).reshape(-1, 2)).view(np.int32)
    from .. import _lib
    return _lib.h2d(torch.from_numpy(a)).to(device, non_blocking=True)


def split_device(keys, num: int, offset: int = 0, count=None):
    """jax.random.split(keys[i], num)[offset:offset+count] for every key: int32 cuda [n, count, 2]."""
    import torch
    from .. import _lib
    count = num - offset if count is None else count
    n = keys.shape[0]
    out = torch.empty((n, count, 2), dtype=torch.int32, device=keys.device)
    _lib.call("toued_key_split", _lib.ptr(keys), n, int(num), int(offset), int(count), _lib.ptr(out), _lib.stream_ptr())
    return out


def chain_device(keys, length: int, want_carry: bool = False):
    """``rng, _rng = split(rng)`` repeated ``length`` times per key: int32 cuda [length, n, 2] of the _rng
    (and the final rng [n, 2] when want_carry)."""
    import torch
    from .. import _lib
    n = keys.shape[0]
    out = torch.empty((length, n, 2), dtype=torch.int32, device=keys.device)
    carry = torch.empty((n, 2), dtype=torch.int32, device=keys.device) if want_carry else None
    _lib.call("toued_key_chain", _lib.ptr(keys), n, int(length), _lib.ptr(out), _lib.ptr(carry), _lib.stream_ptr())
    return (out, carry) if want_carry else out
