"""A2C antagonist (reference agents/a2c.py:12-125), batched over agents, on csrc/a2c_update.cu.

Used by the algorithmic-regret level score (environments/level_sampler.py:293-329): a fresh A2C agent
is trained on the level for ``max_lifetime`` updates and its return is compared with the LPG agent's."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .. import _lib
from ..util import prng
from ..util.data import AgentState


@dataclass
class A2CHyperparams:
    """agents/a2c.py:12-16"""
    gamma: float
    gae_lambda: float
    entropy_coeff: float


def train_a2c_agent(rng, agent_state: AgentState, rollout_manager, num_train_steps: int, hypers: A2CHyperparams,
                    outer_product_quirk: bool = True, record=None):
    """agents/a2c.py:79-125, batched: rng uint32[N, 2].  Returns (agent_state, metrics dict of f32[N]).
    ``record`` (a list) receives the per-update Transition objects (tests only)."""
    from ..environments.gridworld.gridworld import EnvState
    from ..util.data import Transition
    env = rollout_manager.env
    actor, critic = agent_state.actor_state, agent_state.critic_state
    N, W = agent_state.env_state.packed.shape
    L, D, K = rollout_manager.train_rollout_len, env.obs_dim, num_train_steps
    dev = actor.params.device
    i32, f32, u8 = torch.int32, torch.float32, torch.uint8
    a = [actor.params.clone(), torch.empty_like(actor.params)]
    c = [critic.params.clone(), torch.empty_like(critic.params)]
    step = actor.step.clone()
    state = agent_state.env_state.packed.clone()
    obs = torch.empty((N, L + 1, W), dtype=i32, device=dev)
    act = torch.empty((N, L, W), dtype=u8, device=dev)
    rew = torch.empty((N, L, W), dtype=f32, device=dev)
    don = torch.empty((N, L, W), dtype=u8, device=dev)
    st = torch.empty((N, L * W), dtype=torch.int16, device=dev)
    scal = torch.empty((N, 4), dtype=f32, device=dev)
    msum = torch.zeros((N, 2), dtype=f32, device=dev)
    levels = agent_state.level.packed
    keys_d = prng.chain_device(prng.to_device(rng, dev), K)   # a2c.py:95-96, derived on the device
    p, s = _lib.ptr, _lib.stream_ptr()
    if record is None:
        # the whole lifetime in one native call (a Python loop of 3 launches per update is host-bound)
        _lib.call("toued_a2c_train", p(levels), p(keys_d), p(a[0]), p(a[1]), p(c[0]), p(c[1]), p(state), p(obs), p(act),
                  p(rew), p(don), p(st), p(step), p(scal), p(msum), K, N, W, L, D, env.max_grid_size, env.max_n_objs,
                  float(actor.learning_rate), float(critic.learning_rate), float(actor.max_grad_norm),
                  float(hypers.gamma), float(hypers.gae_lambda), float(hypers.entropy_coeff), int(outer_product_quirk), s)
    for k in range(K if record is not None else 0):
        i, o = k & 1, (k & 1) ^ 1
        _lib.call("toued_rollout", p(levels), p(keys_d[k]), p(a[i]), None, p(state), p(obs), p(act), p(rew), p(don),
                  None, N, W, L, D, env.max_grid_size, env.max_n_objs, 0, s)
        _lib.call("toued_sort_tokens", p(obs), p(st), N, W, L, s)
        _lib.call("toued_a2c_update", p(obs), p(act), p(rew), p(don), p(st), p(a[i]), p(c[i]), p(a[o]), p(c[o]),
                  p(levels), p(step), p(scal), N, W, L, D, float(actor.learning_rate), float(critic.learning_rate),
                  float(actor.max_grad_norm), float(hypers.gamma), float(hypers.gae_lambda),
                  float(hypers.entropy_coeff), int(outer_product_quirk), s)
        msum += scal[:, :2]
        if record is not None:
            record.append(Transition(obs.clone(), act.clone(), rew.clone(), don.clone()))
    f = K & 1
    out = agent_state.replace(actor_state=actor.replace(params=a[f], step=step),
                              critic_state=critic.replace(params=c[f], step=step.clone()),
                              env_obs=obs[:, -1].clone(), env_state=EnvState(state, env.max_n_objs))
    return out, {"actor_loss": msum[:, 0] / K, "critic_loss": msum[:, 1] / K}
