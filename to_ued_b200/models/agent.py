"""Tabular Actor / Critic (reference models/agent.py:7-45 with ``actor_net = critic_net = ()``):
one bias-free Dense on the one-hot-plus-time observation, i.e. a D x C table, + softmax
(no softmax for the 1-wide value critic).  Tables are stored padded to 8 columns
(``[N, D, 8]`` f32; columns >= C are zero and never read)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..util import prng

TABLE_PAD = 8


def lecun_normal(key, shape, fan_in) -> np.ndarray:
    """flax.linen.initializers.lecun_normal: truncated normal on [-2, 2] scaled by
    sqrt(1 / fan_in) / 0.87962566 (jax.random.truncated_normal = erfinv of a uniform on
    (erf(-sqrt 2), erf(sqrt 2)))."""
    from scipy.special import erf, erfinv
    lo, hi = erf(-2.0 / np.sqrt(2.0)), erf(2.0 / np.sqrt(2.0))
    u = prng.uniform(key, shape, np.float32(lo), np.float32(hi)).astype(np.float64)
    v = np.sqrt(2.0) * erfinv(u)
    v = np.clip(v, np.nextafter(np.float32(-2), 0), np.nextafter(np.float32(2), 0))
    return (v * (np.sqrt(1.0 / fan_in) / 0.87962566103423978)).astype(np.float32)


def init_tables(keys, obs_dim: int, n_out: int, device="cuda", out=None, mask=None) -> torch.Tensor:
    """One lecun-normal table per key: keys uint32[N, 2] -> f32[N, D, 8] on the GPU (toued_init_tables;
    same draw as ``lecun_normal`` above up to erfinv rounding).  ``out`` / ``mask`` re-initialise only the
    masked agents of an existing tensor."""
    from .. import _lib
    keys = np.ascontiguousarray(np.asarray(keys, np.uint32).reshape(-1, 2))
    n = keys.shape[0]
    kd = _lib.h2d(torch.from_numpy(keys.view(np.int32))).to(device, non_blocking=True)
    if out is None:
        out = torch.empty((n, obs_dim, TABLE_PAD), dtype=torch.float32, device=device)
    md = None
    if mask is not None:
        md = mask if isinstance(mask, torch.Tensor) else _lib.h2d(torch.from_numpy(np.asarray(mask, np.uint8))).to(device, non_blocking=True)
    _lib.call("toued_init_tables", _lib.ptr(kd), _lib.ptr(md), _lib.ptr(out), n, obs_dim, n_out, _lib.stream_ptr())
    return out


class Actor:
    def __init__(self, layers, n_actions):
        if tuple(layers):
            raise NotImplementedError("only actor_net=() (tabular modes) is implemented")
        self.layers, self.n_actions = tuple(layers), n_actions


class Critic:
    def __init__(self, layers, critic_dims):
        if tuple(layers):
            raise NotImplementedError("only critic_net=() (tabular modes) is implemented")
        self.layers, self.critic_dims = tuple(layers), critic_dims
