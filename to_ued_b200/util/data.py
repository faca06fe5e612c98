"""Batched containers mirroring the reference's pytrees (util/data.py:8-68).

All tensors carry a leading agent axis N and live on the GPU; host-side level descriptions
(EnvParams, lifetimes, buffer ids) stay in numpy and are packed into LevelRec records for the
kernels (environments/gridworld/gridworld.py::pack_levels)."""
from __future__ import annotations

from dataclasses import dataclass, replace as _replace
from typing import Any, Optional

import numpy as np
import torch


@dataclass
class LpgHyperparams:
    """util/data.py:8-34"""
    num_agent_updates: int
    agent_target_coeff: float
    policy_entropy_coeff: float
    target_entropy_coeff: float
    policy_l2_coeff: float
    target_l2_coeff: float

    @staticmethod
    def from_run_args(args):
        return LpgHyperparams(
            num_agent_updates=args.num_agent_updates, agent_target_coeff=args.lpg_agent_target_coeff,
            policy_entropy_coeff=args.lpg_policy_entropy_coeff, target_entropy_coeff=args.lpg_target_entropy_coeff,
            policy_l2_coeff=args.lpg_policy_l2_coeff, target_l2_coeff=args.lpg_target_l2_coeff)

    def replace(self, **kw):
        return _replace(self, **kw)


@dataclass
class Transition:
    """util/data.py:37-43 in compact form, layout [N, L, W] (obs: [N, L+1, W]).
    ``obs[:, t]`` is the observation at step t and ``obs[:, t + 1]`` is ``next_obs`` at step t
    (the carried observation; after a done it is the reset observation, as in gymnax)."""
    obs: torch.Tensor       # int32 [N, L+1, W]  row | time << 16
    action: torch.Tensor    # uint8 [N, L, W]
    reward: torch.Tensor    # f32   [N, L, W]
    done: torch.Tensor      # uint8 [N, L, W]

    @property
    def next_obs(self):
        return self.obs[:, 1:]


@dataclass
class Level:
    """util/data.py:46-50 (host side) + its packed device record."""
    env_params: Any                 # EnvParams (numpy, [N])
    lifetime: np.ndarray            # i32[N]
    buffer_id: np.ndarray           # i32[N]
    packed: Optional[torch.Tensor] = None   # uint8 [N, 192] on device

    def __len__(self):
        return len(self.lifetime)


@dataclass
class TrainState:
    """flax TrainState for a tabular network: ``params`` is the padded table [N, D, 8]
    (``kernel`` = params[..., :C]); ``step`` is int32 [N]."""
    params: torch.Tensor
    step: torch.Tensor
    n_out: int
    learning_rate: float
    max_grad_norm: float
    optimizer: str = "SGD"

    @property
    def kernel(self):
        return self.params[..., : self.n_out]

    def replace(self, **kw):
        return _replace(self, **kw)


@dataclass
class AgentState:
    """util/data.py:53-59"""
    actor_state: TrainState
    critic_state: TrainState
    level: Level
    env_obs: torch.Tensor           # int32 [N, W] packed observation
    env_state: Any                  # EnvState ([N, W] packed)
    host_step: Optional[np.ndarray] = None   # host mirror of actor_state.step (deterministic evolution)

    def replace(self, **kw):
        return _replace(self, **kw)
