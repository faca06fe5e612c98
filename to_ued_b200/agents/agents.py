"""Agent construction / evaluation (reference agents/agents.py:11-116), batched over agents."""
from __future__ import annotations

from dataclasses import dataclass, replace as _replace

import numpy as np
import torch

from .. import _lib
from ..util import prng
from ..util.data import TrainState
from ..models.agent import Actor, Critic, init_tables
from ..environments.environments import get_agent_hypers


@dataclass
class AgentHyperparams:
    """agents/agents.py:11-28"""
    actor_net: tuple
    actor_learning_rate: float
    critic_net: tuple
    critic_learning_rate: float
    optimizer: str
    max_grad_norm: float
    critic_dims: int = 1
    convert_nchw: bool = False

    @staticmethod
    def from_args(args):
        d = dict(get_agent_hypers(args.env_name, args.env_mode))
        return AgentHyperparams(**d, critic_dims=args.lpg_target_width)

    def replace(self, **kw):
        return _replace(self, **kw)


def _obs_dim(obs_shape):
    if isinstance(obs_shape, int):
        return obs_shape
    if len(obs_shape) != 1:
        raise NotImplementedError("only flat (tabular) observations are implemented")
    return int(obs_shape[0])


def _create_train_state(keys, n_out, obs_dim, optimizer, learning_rate, max_grad_norm, device):
    if optimizer != "SGD":
        raise NotImplementedError("tabular agents use SGD (configs.py:652-659); Adam agents are the out-of-scope rand_* modes")
    params = init_tables(keys, obs_dim, n_out, device)
    step = torch.zeros(params.shape[0], dtype=torch.int32, device=params.device)
    return TrainState(params=params, step=step, n_out=n_out, learning_rate=learning_rate,
                      max_grad_norm=max_grad_norm, optimizer=optimizer)


def create_agent(rng, agent_params: AgentHyperparams, action_n: int, obs_shape, device="cuda"):
    """agents/agents.py:31-56, batched: rng uint32[N, 2] -> (actor TrainState, critic TrainState).
    Q8 reproduced: the critic model is built from ``actor_net`` (both are () here)."""
    rng = np.asarray(rng, np.uint32).reshape(-1, 2)
    ks = prng.split(rng, 2)
    Actor(agent_params.actor_net, action_n)
    Critic(agent_params.actor_net, agent_params.critic_dims)
    d = _obs_dim(obs_shape)
    actor = _create_train_state(ks[:, 0, :], action_n, d, agent_params.optimizer,
                                agent_params.actor_learning_rate, agent_params.max_grad_norm, device)
    critic = _create_train_state(ks[:, 1, :], agent_params.critic_dims, d, agent_params.optimizer,
                                 agent_params.critic_learning_rate, agent_params.max_grad_norm, device)
    return actor, critic


def create_value_critic(rng, agent_params: AgentHyperparams, obs_shape, device="cuda"):
    """agents/agents.py:59-75, batched over keys."""
    rng = np.asarray(rng, np.uint32).reshape(-1, 2)
    agent_params = agent_params.replace(critic_dims=1)
    Critic(agent_params.actor_net, 1)
    return _create_train_state(rng, 1, _obs_dim(obs_shape), agent_params.optimizer,
                               agent_params.critic_learning_rate, agent_params.max_grad_norm, device)


def eval_agent(rng, rollout_manager, env_params, actor_train_state, num_workers):
    """agents/agents.py:98-106, batched: rng uint32[N, 2] -> mean first-episode return f32[N]."""
    table = actor_train_state.params if hasattr(actor_train_state, "params") else actor_train_state
    k_reset, k_roll = prng.chain_device(prng.to_device(rng, table.device), 2)     # agents.py:99-103
    env_obs, env_state = rollout_manager.batch_reset(k_reset, env_params, num_workers)
    _, _, _, tot = rollout_manager.batch_rollout(k_roll, actor_train_state, env_params, env_obs, env_state,
                                                 eval=True, want_trajectory=False)
    return tot.mean(dim=1)
