"""Gridworld level distributions and per-mode constants (reference
environments/gridworld/configs.py).  ``reset_env_params`` is vectorised over a batch of keys and
follows the reference's key-splitting order exactly (configs.py:12-53), so that a level drawn from
key k here is the level the reference would draw from key k (up to float rounding of exp/log).

Only tabular modes are registered (non-tabular ``rand_*`` modes are out of scope).

Deviation Q4: the reference's ``"tabular"`` / ``"mazes"`` distributions are ``manual: True`` with a
``modes`` tuple that nothing reads (its ``reset_env_params`` would raise KeyError).  Here:
``rng, k = split(rng)``; sub-mode = ``modes[randint(k, (), 0, len(modes))]``; the sub-mode is then
sampled with ``rng`` and padded to the distribution's ENV_MODE_KWARGS.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple, Union

import numpy as np

from ...util import prng
from .gridworld import EnvParams
from .maze_data import MAZE_WALL_MASKS, maze_wall_idxs

f32 = np.float32

# exp(x) from individually rounded f32 operations (mirrors csrc/common.cuh::exp_portable and oracle/rollout.py)
_LOG2E, _LN2_HI, _LN2_LO = f32(1.4426950408889634), f32(0.693359375), f32(-2.12194440e-4)
_EXP_C = [f32(c) for c in (1.0, 1.0, 0.5, 1.6666667e-1, 4.1666668e-2, 8.3333338e-3, 1.3888889e-3, 1.9841270e-4)]


def exp_portable(x):
    x = np.maximum(np.asarray(x, f32), f32(-80.0))
    n = np.rint((x * _LOG2E).astype(f32)).astype(f32)
    r = (x - (n * _LN2_HI).astype(f32)).astype(f32)
    r = (r - (n * _LN2_LO).astype(f32)).astype(f32)
    p = np.full_like(r, _EXP_C[7])
    for c in (_EXP_C[6], _EXP_C[5], _EXP_C[4], _EXP_C[3], _EXP_C[2], _EXP_C[1], _EXP_C[0]):
        p = ((p * r).astype(f32) + c).astype(f32)
    return (p.view(np.int32) + (n.astype(np.int32) << 23)).view(f32)


@dataclass(frozen=True)
class LogUniform:            # configs.py:117-126
    lo: float
    hi: float
    n: Optional[int] = None  # None -> scalar ()
    as_int: bool = False

    def __call__(self, key):
        shape = () if self.n is None else (self.n,)
        # exp_portable: the float contract shared with the device generator (csrc/levelgen.cu), DESIGN.md section 2
        v = exp_portable(prng.uniform(key, shape, f32(np.log(f32(self.lo))), f32(np.log(f32(self.hi))))).astype(f32)
        return np.round(v).astype(np.int32) if self.as_int else v


@dataclass(frozen=True)
class UniformFirstPos:       # configs.py:98-107
    n: int
    lo: float
    hi: float

    def __call__(self, key):
        ks = prng.split(key, 2)
        return np.concatenate([prng.uniform(ks[..., 0, :], (1,), 0.0, self.hi),
                               prng.uniform(ks[..., 1, :], (self.n - 1,), self.lo, self.hi)], -1)


@dataclass(frozen=True)
class Uniform:
    n: int
    lo: float
    hi: float

    def __call__(self, key):
        return prng.uniform(key, (self.n,), self.lo, self.hi)


@dataclass(frozen=True)
class ChoiceRange:           # partial(random.choice, a=jnp.arange(lo, hi))
    lo: int
    hi: int

    def __call__(self, key):
        return (self.lo + prng.randint(key, (), 0, self.hi - self.lo)).astype(np.int32)


@dataclass(frozen=True)
class UniformWalls:          # configs.py:110-114
    n_walls: int
    max_grid_size: int

    def __call__(self, key):
        return prng.shuffle_prefix(key, self.max_grid_size ** 2, self.n_walls)


@dataclass(frozen=True)
class Mode:
    max_steps_in_episode: object
    obj_ids: Tuple[int, ...]
    obj_rewards: object
    obj_p_terminate: object
    obj_p_respawn: object
    n_objs: object
    grid_size: object
    wall_idxs: object
    tabular: bool = True
    auto_collect: bool = True


@dataclass(frozen=True)
class Distribution:
    modes: Tuple[str, ...]


def _rand_mode(steps, n_types, nobj, grid, n_walls, max_grid):
    return Mode(LogUniform(steps[0], steps[1], None, True), tuple(range(n_types)),
                UniformFirstPos(n_types, -1.0, 1.0), LogUniform(1e-2, 1.0, n_types), LogUniform(1e-3, 1e-1, n_types),
                ChoiceRange(*nobj), ChoiceRange(*grid), UniformWalls(n_walls, max_grid))


def _maze_mode(name):        # configs.py:129-145
    return Mode(LogUniform(25, 50, None, True), (0, 1, 2), Uniform(3, 0.0, 1.0), LogUniform(1e-2, 1.0, 3),
                LogUniform(1e-3, 1e-1, 3), 3, 13, tuple(maze_wall_idxs(name)))


def _wall_line(g, col=None, row=None, gaps=()):
    cells = [r * g + col for r in range(g)] if col is not None else [row * g + c for c in range(g)]
    return [c for c in cells if c not in gaps]


ENV_MODE_PARAMS = {
    "dense": Mode(500, (0, 0, 1, 2), (1.0, -1.0, -1.0), (0.0, 0.5, 0.0), (0.05, 0.1, 0.5), 4, 11, ()),
    "sparse": Mode(50, (0, 1), (1.0, -1.0), (1.0, 1.0), (0.0, 0.0), 2, 13, ()),
    "long": Mode(1000, (0, 0, 1, 1), (1.0, -1.0), (0.0, 0.5), (0.01, 1.0), 4, 11, ()),
    "longer": Mode(2000, (0, 0, 1, 1, 1), (1.0, -1.0), (0.1, 0.8), (0.01, 1.0), 5, 9,
                   tuple(_wall_line(9, col=4, gaps=(9 * 1 + 4, 9 * 7 + 4)))),
    "long_dense": Mode(2000, (0, 0, 0, 0), (1.0,), (0.0,), (0.005,), 4, 11,
                       tuple(sorted(set(_wall_line(11, col=5, gaps=(5, 11 * 7 + 5)))
                                    | set(_wall_line(11, row=4, gaps=(11 * 4 + 2, 11 * 4 + 8)))))),
    "small": _rand_mode((20, 100), 3, (1, 4), (4, 7), 7, 6),
    "medium": _rand_mode((100, 250), 4, (2, 5), (6, 9), 10, 8),
    "large": _rand_mode((250, 750), 5, (2, 6), (8, 11), 15, 10),
    "all": _rand_mode((20, 750), 5, (1, 6), (4, 11), 15, 10),
    "debug": _rand_mode((5, 10), 2, (1, 3), (3, 5), 4, 4),
    **{m: _maze_mode(m) for m in MAZE_WALL_MASKS},
    "tabular": Distribution(("dense", "sparse", "long", "longer", "long_dense")),
    "mazes": Distribution(tuple(MAZE_WALL_MASKS)),
}


def _kw(o, t, g):
    return {"max_n_objs": o, "max_n_obj_types": t, "max_grid_size": g, "tabular": True}


ENV_MODE_KWARGS = {
    "dense": _kw(4, 3, 11), "sparse": _kw(2, 2, 13), "long": _kw(4, 2, 11), "longer": _kw(5, 2, 9),
    "long_dense": _kw(4, 1, 11), "tabular": _kw(5, 3, 13), "small": _kw(3, 3, 6), "medium": _kw(4, 4, 8),
    "large": _kw(5, 5, 10), "all": _kw(5, 5, 10), "debug": _kw(2, 2, 4),
    **{m: _kw(3, 3, 13) for m in MAZE_WALL_MASKS}, "mazes": _kw(3, 3, 13),
}
ENV_MODE_EPISODE_LEN = {
    "dense": 500, "sparse": 50, "long": 1000, "longer": 2000, "long_dense": 2000, "tabular": 2000,
    "small": 100, "medium": 250, "large": 750, "all": 750, "debug": 10,
    **{m: 50 for m in MAZE_WALL_MASKS}, "mazes": 50,
}

_TABULAR_LIFETIME, _SMALL_LIFETIME, _MEDIUM_LIFETIME = 5 * 500, 5 * 50, 5 * 200
_LARGE_LIFETIME, _MAZE_LIFETIME, _DEBUG_LIFETIME = 5 * 500, 5 * 500, 4
ENV_MODE_LIFETIME = {
    **{m: _TABULAR_LIFETIME for m in ("dense", "sparse", "long", "longer", "long_dense", "tabular")},
    "small": _SMALL_LIFETIME, "medium": _MEDIUM_LIFETIME, "large": _LARGE_LIFETIME, "all": _MEDIUM_LIFETIME,
    "all_shortlife": _SMALL_LIFETIME,
    "all_randlife": LogUniform(_SMALL_LIFETIME // 5, _SMALL_LIFETIME, None, True),
    "all_vrandlife": LogUniform(_SMALL_LIFETIME // 25, _SMALL_LIFETIME, None, True),
    "debug": _DEBUG_LIFETIME,
    **{m: _MAZE_LIFETIME for m in MAZE_WALL_MASKS}, "mazes": _MAZE_LIFETIME,
}
for _alias in ("all_shortlife", "all_randlife", "all_vrandlife"):
    ENV_MODE_PARAMS[_alias] = ENV_MODE_PARAMS["all"]
    ENV_MODE_KWARGS[_alias] = ENV_MODE_KWARGS["all"]
    ENV_MODE_EPISODE_LEN[_alias] = ENV_MODE_EPISODE_LEN["all"]
ENV_MODE_LIFETIME_MAX = {m: (v if isinstance(v, int) else int(v.hi)) for m, v in ENV_MODE_LIFETIME.items()}

_TABULAR_HYPERS = {"actor_net": (), "actor_learning_rate": 4e1, "critic_net": (), "critic_learning_rate": 4e0,
                   "optimizer": "SGD", "max_grad_norm": 0.5}
MODE_AGENT_HYPERS = {m: _TABULAR_HYPERS for m in ENV_MODE_KWARGS}


def get_env_spec(mode: str):
    return dict(ENV_MODE_KWARGS[mode]), ENV_MODE_EPISODE_LEN[mode]


def get_max_lifetime(mode: str):
    return ENV_MODE_LIFETIME_MAX[mode]


def get_agent_hypers(mode: str):
    return MODE_AGENT_HYPERS[mode]


def reset_lifetime(rng, env_mode: str):
    """configs.py:56-57, batched over keys [B, 2]."""
    v = ENV_MODE_LIFETIME[env_mode]
    b = np.asarray(rng).shape[0]
    return np.full(b, v, np.int32) if isinstance(v, int) else v(rng).astype(np.int32)


def _next(rng):
    ks = prng.split(rng, 2)
    return ks[:, 0, :], ks[:, 1, :]


def _draw(key, spec, b, dtype):
    """_sample_param (configs.py:83-88): callable -> one more split, constant -> broadcast."""
    if callable(spec):
        return np.asarray(spec(prng.split(key, 2)[:, 1, :]), dtype)
    return np.broadcast_to(np.asarray(spec, dtype), (b,) + np.shape(spec)).copy()


def _sample_mode(rng, mode: Mode, kw) -> EnvParams:
    b = rng.shape[0]
    O, T, G2 = kw["max_n_objs"], kw["max_n_obj_types"], kw["max_grid_size"] ** 2
    out = {}
    ids = list(mode.obj_ids) + [-1] * (O - len(mode.obj_ids))
    out["obj_ids"] = np.broadcast_to(np.asarray(ids, np.int32), (b, O)).copy()
    for name in ("obj_rewards", "obj_p_terminate", "obj_p_respawn"):       # _sample_obj_param
        rng, k = _next(rng)
        spec = getattr(mode, name)
        val = np.asarray(spec(k), f32) if callable(spec) else np.broadcast_to(np.asarray(spec, f32), (b, len(spec)))
        out[name] = np.concatenate([val, np.zeros((b, T - val.shape[1]), f32)], 1)
    out["auto_collect"] = np.full(b, mode.auto_collect)
    out["random_respawn"] = np.full(b, not mode.tabular)
    for name in ("max_steps_in_episode", "n_objs", "grid_size"):
        rng, k = _next(rng)
        out[name] = _draw(k, getattr(mode, name), b, np.int32)
    rng, k = _next(rng)
    walls = np.zeros((b, G2), bool)
    if callable(mode.wall_idxs):
        wi = mode.wall_idxs(prng.split(k, 2)[:, 1, :])
        np.put_along_axis(walls, wi.astype(np.int64), True, axis=1)
    else:
        walls[:, list(mode.wall_idxs)] = True
    out["walls"] = walls
    valid = (np.arange(G2)[None, :] < (out["grid_size"] ** 2)[:, None]) & ~walls
    rng, k = _next(rng)
    pos = prng.masked_topk(k, valid, O + 1)
    out["start_pos"], out["static_obj_poss"] = pos[:, 0].astype(np.int32), pos[:, 1:].astype(np.int32)
    return EnvParams(**out)


def reset_env_params(rng, env_mode: str) -> EnvParams:
    """configs.py:12-53, batched over keys uint32[B, 2]."""
    rng = np.asarray(rng, np.uint32).reshape(-1, 2)
    spec = ENV_MODE_PARAMS[env_mode]
    kw = ENV_MODE_KWARGS[env_mode]
    if isinstance(spec, Distribution):
        rng, k = _next(rng)
        which = prng.randint(k, (), 0, len(spec.modes))
        parts, order = [], []
        for i, sub in enumerate(spec.modes):
            sel = np.nonzero(which == i)[0]
            if len(sel):
                parts.append(_sample_mode(rng[sel], ENV_MODE_PARAMS[sub], kw))
                order.append(sel)
        merged = EnvParams.concat(parts)
        inv = np.argsort(np.concatenate(order), kind="stable")
        return merged[inv]
    return _sample_mode(rng, spec, kw)
