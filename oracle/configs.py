"""Level generator + per-mode tables (oracle).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates environments/gridworld/configs.py:12-145 (``reset_env_params`` and its samplers) and the
per-mode tables :148-707, one level at a time with scalar keys (clarity over speed; the product's
generator in to_ued_b200/environments/gridworld/configs.py is vectorised and is checked against
this one).

Deviation Q4: ``"tabular"`` and ``"mazes"`` are ``manual: True`` distributions
whose ``modes`` tuple is never read by the reference (``reset_env_params`` would KeyError).
Defined here as: ``rng, k = split(rng)``; ``mode = modes[randint(k, (), 0, len(modes))]``; sample
that sub-mode with ``rng`` and pad objects/types/walls to the distribution's ENV_MODE_KWARGS.
"""
from __future__ import annotations

import numpy as np

from . import prng
from .gridworld import EnvParams
from .maze_data import MAZE_WALL_MASKS, maze_wall_idxs

F32 = np.float32


# ---- samplers (configs.py:98-126) -------------------------------------------------------------
def uniform_first_pos(key, n, minval, maxval):
    k = prng.split(key, 2)
    return np.concatenate([prng.uniform(k[0], (1,), 0.0, maxval), prng.uniform(k[1], (n - 1,), minval, maxval)])


def log_uniform(key, shape, minval, maxval):
    """configs.py:117-126 ``exp(uniform(log lo, log hi))``.  Float contract: the exponential is
    ``exp_portable`` (individually rounded f32 operations, the same routine as the policy softmax), so that the host
    numpy generator and the device level generator (csrc/levelgen.cu) produce the same bits; jax's own exp rounding is
    unknowable here."""
    from .rollout import exp_portable
    lo, hi = F32(np.log(F32(minval))), F32(np.log(F32(maxval)))
    return exp_portable(prng.uniform(key, shape, lo, hi)).astype(F32)


def log_uniform_int(key, shape, minval, maxval):
    return np.round(log_uniform(key, shape, minval, maxval)).astype(np.int32)


def uniform_wall_idxs(key, n_walls, max_grid_size):
    return prng.choice_no_replace_uniform(key, max_grid_size ** 2, n_walls)


def choice_arange(key, lo, hi):
    """partial(random.choice, a=jnp.arange(lo, hi))"""
    return np.int32(lo + prng.choice_uniform(key, hi - lo))


def _P(fn, **kw):
    return lambda key: fn(key, **kw)


def _dist(name_steps, n_types, rewards_lo, nobj_range, grid_range, n_walls, max_grid):
    lo, hi = name_steps
    return dict(
        max_steps_in_episode=_P(log_uniform_int, shape=(), minval=lo, maxval=hi),
        obj_ids=list(range(n_types)),
        obj_rewards=_P(uniform_first_pos, n=n_types, minval=rewards_lo, maxval=1.0),
        obj_p_terminate=_P(log_uniform, shape=(n_types,), minval=1e-2, maxval=1.0),
        obj_p_respawn=_P(log_uniform, shape=(n_types,), minval=1e-3, maxval=1e-1),
        n_objs=_P(choice_arange, lo=nobj_range[0], hi=nobj_range[1]),
        grid_size=_P(choice_arange, lo=grid_range[0], hi=grid_range[1]),
        wall_idxs=_P(uniform_wall_idxs, n_walls=n_walls, max_grid_size=max_grid),
        tabular=True, auto_collect=True)


def _maze(name):
    return dict(
        max_steps_in_episode=_P(log_uniform_int, shape=(), minval=25, maxval=50),
        obj_ids=[0, 1, 2],
        obj_rewards=lambda key: prng.uniform(key, (3,), 0.0, 1.0),
        obj_p_terminate=_P(log_uniform, shape=(3,), minval=1e-2, maxval=1.0),
        obj_p_respawn=_P(log_uniform, shape=(3,), minval=1e-3, maxval=1e-1),
        n_objs=3, grid_size=13, wall_idxs=np.array(maze_wall_idxs(name), np.int32),
        tabular=True, auto_collect=True)


def _longer_walls():
    a = np.arange(81)
    return a[(a % 9 == 4) & ~np.isin(a, [9 * 1 + 4, 9 * 7 + 4])]


def _long_dense_walls():
    a = np.arange(121)
    v = (a % 11 == 5) & ~np.isin(a, [5, 11 * 7 + 5])
    h = (a // 11 == 4) & ~np.isin(a, [11 * 4 + 2, 11 * 4 + 8])
    return a[v | h]


_NOWALL = np.array([], np.int32)
ENV_MODE_PARAMS = {   # configs.py:148-420 (tabular modes only)
    "dense": dict(max_steps_in_episode=500, obj_ids=[0, 0, 1, 2], obj_rewards=[1.0, -1.0, -1.0],
                  obj_p_terminate=[0.0, 0.5, 0.0], obj_p_respawn=[0.05, 0.1, 0.5], n_objs=4, grid_size=11,
                  wall_idxs=_NOWALL, tabular=True, auto_collect=True),
    "sparse": dict(max_steps_in_episode=50, obj_ids=[0, 1], obj_rewards=[1.0, -1.0], obj_p_terminate=[1.0, 1.0],
                   obj_p_respawn=[0.0, 0.0], n_objs=2, grid_size=13, wall_idxs=_NOWALL, tabular=True, auto_collect=True),
    "long": dict(max_steps_in_episode=1000, obj_ids=[0, 0, 1, 1], obj_rewards=[1.0, -1.0], obj_p_terminate=[0.0, 0.5],
                 obj_p_respawn=[0.01, 1.0], n_objs=4, grid_size=11, wall_idxs=_NOWALL, tabular=True, auto_collect=True),
    "longer": dict(max_steps_in_episode=2000, obj_ids=[0, 0, 1, 1, 1], obj_rewards=[1.0, -1.0],
                   obj_p_terminate=[0.1, 0.8], obj_p_respawn=[0.01, 1.0], n_objs=5, grid_size=9,
                   wall_idxs=_longer_walls(), tabular=True, auto_collect=True),
    "long_dense": dict(max_steps_in_episode=2000, obj_ids=[0, 0, 0, 0], obj_rewards=[1.0], obj_p_terminate=[0.0],
                       obj_p_respawn=[0.005], n_objs=4, grid_size=11, wall_idxs=_long_dense_walls(),
                       tabular=True, auto_collect=True),
    "small": _dist((20, 100), 3, -1.0, (1, 4), (4, 7), 7, 6),
    "medium": _dist((100, 250), 4, -1.0, (2, 5), (6, 9), 10, 8),
    "large": _dist((250, 750), 5, -1.0, (2, 6), (8, 11), 15, 10),
    "all": _dist((20, 750), 5, -1.0, (1, 6), (4, 11), 15, 10),
    "debug": _dist((5, 10), 2, -1.0, (1, 3), (3, 5), 4, 4),
    **{m: _maze(m) for m in MAZE_WALL_MASKS},
    "tabular": dict(manual=True, modes=("dense", "sparse", "long", "longer", "long_dense")),
    "mazes": dict(manual=True, modes=tuple(MAZE_WALL_MASKS)),
}
for _m in ("all_shortlife", "all_randlife", "all_vrandlife"):
    ENV_MODE_PARAMS[_m] = ENV_MODE_PARAMS["all"]

_K = lambda o, t, g: dict(max_n_objs=o, max_n_obj_types=t, max_grid_size=g, tabular=True)
ENV_MODE_KWARGS = {   # configs.py:423-586
    "dense": _K(4, 3, 11), "sparse": _K(2, 2, 13), "long": _K(4, 2, 11), "longer": _K(5, 2, 9),
    "long_dense": _K(4, 1, 11), "tabular": _K(5, 3, 13), "small": _K(3, 3, 6), "medium": _K(4, 4, 8),
    "large": _K(5, 5, 10), "all": _K(5, 5, 10), "debug": _K(2, 2, 4),
    **{m: _K(3, 3, 13) for m in MAZE_WALL_MASKS}, "mazes": _K(3, 3, 13),
}
ENV_MODE_EPISODE_LEN = {   # configs.py:546-593
    "dense": 500, "sparse": 50, "long": 1000, "longer": 2000, "long_dense": 2000, "tabular": 2000,
    "small": 100, "medium": 250, "large": 750, "all": 750, "debug": 10,
    **{m: 50 for m in MAZE_WALL_MASKS}, "mazes": 50,
}
for _m in ("all_shortlife", "all_randlife", "all_vrandlife"):
    ENV_MODE_KWARGS[_m] = ENV_MODE_KWARGS["all"]
    ENV_MODE_EPISODE_LEN[_m] = ENV_MODE_EPISODE_LEN["all"]

# configs.py:596-650
_TAB, _SMALL, _MED, _LARGE, _MAZE, _DEBUG = 2500, 250, 1000, 2500, 2500, 4
ENV_MODE_LIFETIME = {
    **{m: (lambda _: _TAB) for m in ("dense", "sparse", "long", "longer", "long_dense", "tabular")},
    "small": lambda _: _SMALL, "medium": lambda _: _MED, "large": lambda _: _LARGE, "all": lambda _: _MED,
    "all_shortlife": lambda _: _SMALL,
    "all_randlife": _P(log_uniform_int, shape=(), minval=_SMALL // 5, maxval=_SMALL),
    "all_vrandlife": _P(log_uniform_int, shape=(), minval=_SMALL // 25, maxval=_SMALL),
    "debug": lambda _: _DEBUG,
    **{m: (lambda _: _MAZE) for m in MAZE_WALL_MASKS}, "mazes": lambda _: _MAZE,
}
ENV_MODE_LIFETIME_MAX = {"all_randlife": _SMALL, "all_vrandlife": _SMALL}
ENV_MODE_LIFETIME_MAX.update({m: f(None) for m, f in ENV_MODE_LIFETIME.items() if m not in ENV_MODE_LIFETIME_MAX})

# configs.py:652-659
TABULAR_HYPERS = dict(actor_net=(), actor_learning_rate=4e1, critic_net=(), critic_learning_rate=4e0,
                      optimizer="SGD", max_grad_norm=0.5)


def get_env_spec(mode):
    return dict(ENV_MODE_KWARGS[mode]), ENV_MODE_EPISODE_LEN[mode]


def get_max_lifetime(mode):
    return ENV_MODE_LIFETIME_MAX[mode]


# ---- reset_env_params (configs.py:12-53), one level ---------------------------------------------
def _sample_param(key, param):
    """configs.py:83-88"""
    if callable(param):
        return param(prng.split(key, 2)[1])
    return param


def _sample_obj_param(key, param, T):
    """configs.py:75-80"""
    val = np.asarray(param(key) if callable(param) else param, F32)
    return np.concatenate([val, np.zeros(T - len(val), F32)])


def reset_env_params_one(rng, env_mode, kwargs=None):
    mps = ENV_MODE_PARAMS[env_mode]
    kwargs = kwargs or ENV_MODE_KWARGS[env_mode]
    if mps.get("manual"):                                   # deviation Q4
        ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
        sub = mps["modes"][int(prng.randint(k, (), 0, len(mps["modes"])))]
        return reset_env_params_one(rng, sub, kwargs)
    O, T, G2 = kwargs["max_n_objs"], kwargs["max_n_obj_types"], kwargs["max_grid_size"] ** 2
    out = {"obj_ids": np.array(list(mps["obj_ids"]) + [-1] * (O - len(mps["obj_ids"])), np.int32)}
    for name in ("obj_rewards", "obj_p_terminate", "obj_p_respawn"):
        ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
        out[name] = _sample_obj_param(k, mps[name], T)
    out["auto_collect"] = mps["auto_collect"]
    out["random_respawn"] = not mps["tabular"]
    for name in ("max_steps_in_episode", "n_objs", "grid_size"):
        ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
        out[name] = np.int32(_sample_param(k, mps[name]))
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    wall_idxs = np.asarray(_sample_param(k, mps["wall_idxs"]), np.int64).reshape(-1)
    walls = np.zeros(G2, bool); walls[wall_idxs] = True
    out["walls"] = walls
    all_pos = np.arange(G2)
    valid = (all_pos < int(out["grid_size"]) ** 2) & ~np.isin(all_pos, wall_idxs)
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    pos = prng.choice_no_replace_p(k, valid.astype(F32), O + 1)
    out["start_pos"], out["static_obj_poss"] = np.int32(pos[0]), pos[1:].astype(np.int32)
    return out


def reset_lifetime_one(rng, env_mode):
    """configs.py:56-57"""
    return np.int32(ENV_MODE_LIFETIME[env_mode](rng))


def reset_env_params(rngs, env_mode):
    """environments/environments.py:23-38 vmapped over keys [B, 2] (level_sampler.py:98-101):
    ``p_rng, l_rng = split(rng)``.  Returns (EnvParams[B], lifetimes i32[B])."""
    rngs = np.asarray(rngs, np.uint32).reshape(-1, 2)
    levels, lifetimes = [], []
    for r in rngs:
        ks = prng.split(r, 2)
        levels.append(reset_env_params_one(ks[0], env_mode))
        lifetimes.append(reset_lifetime_one(ks[1], env_mode))
    stack = lambda n, dt: np.stack([np.asarray(l[n]) for l in levels]).astype(dt)
    p = EnvParams(
        max_steps_in_episode=stack("max_steps_in_episode", np.int32), random_respawn=stack("random_respawn", bool),
        auto_collect=stack("auto_collect", bool), grid_size=stack("grid_size", np.int32), walls=stack("walls", bool),
        start_pos=stack("start_pos", np.int32), n_objs=stack("n_objs", np.int32), obj_ids=stack("obj_ids", np.int32),
        static_obj_poss=stack("static_obj_poss", np.int32), obj_rewards=stack("obj_rewards", F32),
        obj_p_terminate=stack("obj_p_terminate", F32), obj_p_respawn=stack("obj_p_respawn", F32))
    return p, np.asarray(lifetimes, np.int32)
