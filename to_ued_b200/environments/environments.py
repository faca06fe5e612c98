"""Name -> environment dispatch (reference environments/environments.py:11-63, gridworld branch;
the gymnax branch is out of scope: no BASELINE config uses it)."""
from __future__ import annotations

import numpy as np

from ..util import prng
from .gridworld import gridworld as grid
from .gridworld import configs as grid_conf


def get_env(env_name: str, env_kwargs: dict):
    if env_name in grid.registered_envs:
        return grid.GridWorld(**env_kwargs)
    raise ValueError(f"Environment {env_name} not registered in any environment sources.")


def reset_env_params(rng, env_name: str, env_mode: str):
    """Reset environment parameters and agent lifetime (environments.py:23-38), batched over keys."""
    if env_name not in grid.registered_envs:
        raise ValueError(f"Environment {env_name} has no parameter reset method.")
    rng = np.asarray(rng, np.uint32).reshape(-1, 2)
    ks = prng.split(rng, 2)
    return grid_conf.reset_env_params(ks[:, 0, :], env_mode), grid_conf.reset_lifetime(ks[:, 1, :], env_mode)


def get_env_spec(env_name: str, env_mode: str):
    if env_name not in grid.registered_envs:
        raise ValueError(f"Environment {env_name} has no get env spec method.")
    kwargs, max_rollout_len = grid_conf.get_env_spec(env_mode)
    return kwargs, max_rollout_len, grid_conf.get_max_lifetime(env_mode)


def get_agent_hypers(env_name: str, env_mode: str = None):
    if env_name in grid.registered_envs:
        return grid_conf.get_agent_hypers(env_mode)
    raise ValueError(f"Environment {env_name} has no get agent hyperparameters method.")
