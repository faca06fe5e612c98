"""Host side of the device level generator (csrc/levelgen.cu): the mode tables of configs.py packed into the ``GenDesc``
the kernel reads, and ``generate_levels`` = reference ``reset_env_params`` + ``reset_lifetime`` for a batch of keys with
the levels staying on the device.  Bit-exact with the numpy generator in configs.py / oracle/configs.py
(tests/test_16_levelgen_gpu.py)."""
from __future__ import annotations

import struct

import numpy as np
import torch

from ... import _lib
from . import configs as C

_DESC_CACHE = {}
f32 = np.float32


def _param(spec, n_types_total):
    """GenParam: kind, n, lo, hi, vals[8]"""
    vals = [0.0] * 8
    if isinstance(spec, C.LogUniform):
        kind, n, lo, hi = 1, spec.n, float(f32(np.log(f32(spec.lo)))), float(f32(np.log(f32(spec.hi))))
    elif isinstance(spec, C.Uniform):
        kind, n, lo, hi = 2, spec.n, float(f32(spec.lo)), float(f32(spec.hi))
    elif isinstance(spec, C.UniformFirstPos):
        kind, n, lo, hi = 3, spec.n, float(f32(spec.lo)), float(f32(spec.hi))
    else:
        kind, n, lo, hi = 0, len(spec), 0.0, 0.0
        vals[:n] = [float(f32(v)) for v in spec]
    assert n <= min(8, n_types_total)
    return struct.pack("<iiff8f", kind, n, lo, hi, *vals)


def _scalar(spec):
    """GenScalar: kind, a, b, lo, hi"""
    if isinstance(spec, C.LogUniform):
        assert spec.as_int and spec.n is None
        return struct.pack("<iiiff", 1, 0, 0, float(f32(np.log(f32(spec.lo)))), float(f32(np.log(f32(spec.hi)))))
    if isinstance(spec, C.ChoiceRange):
        return struct.pack("<iiiff", 2, spec.lo, spec.hi, 0.0, 0.0)
    return struct.pack("<iiiff", 0, int(spec), 0, 0.0, 0.0)


def _mode(m: C.Mode, kw):
    O, T, G2 = kw["max_n_objs"], kw["max_n_obj_types"], kw["max_grid_size"] ** 2
    mask = np.zeros(256, bool)
    if isinstance(m.wall_idxs, C.UniformWalls):
        assert m.wall_idxs.max_grid_size ** 2 == G2, "uniform walls are drawn over the distribution's grid"
        wall_kind, n_walls = 1, m.wall_idxs.n_walls
    else:
        wall_kind, n_walls = 0, 0
        mask[list(m.wall_idxs)] = True
    words = np.packbits(mask, bitorder="little").view("<u4")
    ids = list(m.obj_ids) + [-1] * (8 - len(m.obj_ids))
    assert len(m.obj_ids) <= O and m.tabular
    return (_param(m.obj_rewards, T) + _param(m.obj_p_terminate, T) + _param(m.obj_p_respawn, T)
            + _scalar(m.max_steps_in_episode) + _scalar(m.n_objs) + _scalar(m.grid_size)
            + struct.pack("<ii", wall_kind, n_walls) + words.tobytes() + struct.pack("<8i", *ids))


def gen_desc(env_mode: str, device="cuda") -> torch.Tensor:
    """The packed GenDesc of ``env_mode`` on the device (cached)."""
    key = (env_mode, str(device))
    if key in _DESC_CACHE:
        return _DESC_CACHE[key]
    spec, kw = C.ENV_MODE_PARAMS[env_mode], C.ENV_MODE_KWARGS[env_mode]
    modes = [C.ENV_MODE_PARAMS[s] for s in spec.modes] if isinstance(spec, C.Distribution) else [spec]
    if len(modes) > 12:
        raise ValueError(f"{env_mode}: more than 12 sub-modes")
    life = C.ENV_MODE_LIFETIME[env_mode]
    blob = struct.pack("<iiii", len(modes), kw["max_n_objs"], kw["max_n_obj_types"], kw["max_grid_size"]) + _scalar(life)
    blob += b"".join(_mode(m, kw) for m in modes)
    total = _lib.lib().toued_generate_levels_desc_bytes()
    blob += b"\0" * (total - len(blob))
    assert len(blob) == total, (len(blob), total)
    t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(device)
    _DESC_CACHE[key] = t
    return t


def generate_levels(keys, env_mode: str, buffer_ids=None, device="cuda", want_lifetimes=False):
    """keys: uint32[n, 2] (numpy, or an int32 device tensor [n, 2]) -> LevelRec uint8[n, 192] on the device
    (+ int32[n] lifetimes on the device when asked)."""
    from ...util import prng
    kd = prng.to_device(keys, device)
    n = kd.shape[0]
    out = torch.empty((n, 192), dtype=torch.uint8, device=device)
    life = torch.empty(n, dtype=torch.int32, device=device) if want_lifetimes else None
    ids = None
    if buffer_ids is not None:
        ids = buffer_ids if isinstance(buffer_ids, torch.Tensor) else \
            _lib.h2d(torch.from_numpy(np.ascontiguousarray(buffer_ids, np.int32))).to(device, non_blocking=True)
    _lib.call("toued_generate_levels", _lib.ptr(gen_desc(env_mode, device)), _lib.ptr(kd), _lib.ptr(ids), _lib.ptr(out),
              _lib.ptr(life), n, _lib.stream_ptr())
    return (out, life) if want_lifetimes else out
