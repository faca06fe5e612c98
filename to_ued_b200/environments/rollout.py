"""``RolloutWrapper`` (reference environments/rollout.py:13-115) on the fused CUDA rollout kernel.

``batch_reset`` / ``batch_rollout`` operate on a *batch of agents* (the reference vmaps the
single-agent versions over agents; here the agent axis is explicit)."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..util.data import Transition
from .environments import get_env
from .gridworld.gridworld import EnvState, levels_to_device, pack_levels


def _keys_to_device(rng, device):
    if isinstance(rng, torch.Tensor):
        return rng.to(device).contiguous().view(torch.int32)
    a = np.ascontiguousarray(np.asarray(rng, np.uint32)).view(np.int32)
    return _lib.h2d(torch.from_numpy(a)).to(device, non_blocking=True)


class RolloutWrapper:
    def __init__(self, env_name: str = "GridWorld-v0", train_rollout_len: Optional[int] = None,
                 eval_rollout_len: Optional[int] = None, env_kwargs: dict = {}, return_info: bool = False):
        self.env_name = env_name
        self.env_kwargs = env_kwargs
        self.env = get_env(env_name, env_kwargs)
        self.train_rollout_len = train_rollout_len
        self.eval_rollout_len = eval_rollout_len
        self.return_info = return_info

    def _levels(self, env_params):
        if isinstance(env_params, torch.Tensor):
            return env_params
        if hasattr(env_params, "packed") and env_params.packed is not None:
            return env_params.packed
        return levels_to_device(pack_levels(env_params))

    # --- ENVIRONMENT RESET ---  (rollout.py:38-42)
    def batch_reset(self, rng, env_params, num_workers):
        """-> (obs int32[N, W], EnvState).  rng is accepted for signature parity (tabular resets
        consume no randomness)."""
        return self.env.reset(rng, self._levels(env_params), num_workers)

    # --- ENVIRONMENT ROLLOUT ---  (rollout.py:45-102)
    def batch_rollout(self, rng, train_state, env_params, init_obs, init_state, eval=False,
                      forced_actions=None, want_trajectory=True):
        """rng: uint32[N, 2] (one key per agent); train_state: actor TrainState or table tensor
        [N, D, 8]; returns (Transition, end_obs, end_state, first_episode_return[N, W])."""
        env = self.env
        lv = self._levels(env_params)
        table = train_state.params if hasattr(train_state, "params") else train_state
        n, w = init_state.packed.shape
        L = self.eval_rollout_len if eval else self.train_rollout_len
        dev = lv.device
        keys = _keys_to_device(rng, dev)
        st = init_state.packed.clone()
        ret = torch.empty((n, w), dtype=torch.float32, device=dev)
        if want_trajectory:
            obs = torch.empty((n, L + 1, w), dtype=torch.int32, device=dev)
            act = torch.empty((n, L, w), dtype=torch.uint8, device=dev)
            rew = torch.empty((n, L, w), dtype=torch.float32, device=dev)
            don = torch.empty((n, L, w), dtype=torch.uint8, device=dev)
        else:
            obs = act = rew = don = None
        fa = None if forced_actions is None else torch.as_tensor(forced_actions, device=dev).to(torch.uint8).contiguous()
        _lib.call("toued_rollout", _lib.ptr(lv), _lib.ptr(keys), _lib.ptr(table), _lib.ptr(fa), _lib.ptr(st),
                  _lib.ptr(obs), _lib.ptr(act), _lib.ptr(rew), _lib.ptr(don), _lib.ptr(ret), n, w, L,
                  env.obs_dim, env.max_grid_size, env.max_n_objs, 0, _lib.stream_ptr())
        traj = Transition(obs, act, rew, don) if want_trajectory else None
        end_obs = obs[:, -1] if want_trajectory else None
        return traj, end_obs, EnvState(st, env.max_n_objs), ret
