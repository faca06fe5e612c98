"""LPG-driven agent training (reference agents/lpg_agent.py:13-140), batched over agents and
executed by the CUDA kernels in csrc/{rollout,lpg_forward,agent_update}.cu.

``train_lpg_agent`` keeps the reference's signature and return value (agent_state, rollouts,
metrics); when a ``Tape`` is passed it additionally records what the meta-gradient needs
(per-update tables, trajectories, LPG activations) so that meta/train.py can run the hand-written
backward pass where the reference calls ``jax.grad``."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, List, Optional

import numpy as np
import torch

from .. import _lib
from ..util import prng
from ..util.data import AgentState, Transition


@dataclass
class LPGAgentMetrics:
    """agents/lpg_agent.py:13-28; fields are f32[N] tensors (mean over the K updates)."""
    policy_l2: Any
    policy_entropy: Any
    critic_loss: Any
    critic_l2: Any
    critic_entropy: Any

    def as_dict(self):
        return {"policy_l2": self.policy_l2, "policy_entropy": self.policy_entropy,
                "critic_loss": self.critic_loss, "critic_l2": self.critic_l2,
                "critic_entropy": self.critic_entropy}


class Tape:
    """Saved-for-backward buffers of K agent updates (all on the GPU, allocated once and reused)."""

    def __init__(self, n_agents, n_workers, rollout_len, obs_dim, num_updates, device, keep_gates=True,
                 precision=None):
        N, W, L, D, K = n_agents, n_workers, rollout_len, obs_dim, num_updates
        R, dev = N * W, device
        f32, u8, i32 = torch.float32, torch.uint8, torch.int32
        self.N, self.W, self.L, self.D, self.K, self.R = N, W, L, D, K, R
        self.actor = torch.empty((K + 1, N, D, 8), dtype=f32, device=dev)
        self.critic = torch.empty((K + 1, N, D, 8), dtype=f32, device=dev)
        # K train rollouts + 1 eval rollout
        self.obs = torch.empty((K + 1, N, L + 1, W), dtype=i32, device=dev)
        self.action = torch.empty((K + 1, N, L, W), dtype=u8, device=dev)
        self.reward = torch.empty((K + 1, N, L, W), dtype=f32, device=dev)
        self.done = torch.empty((K + 1, N, L, W), dtype=u8, device=dev)
        self.sorted_tok = torch.empty((K + 1, N, L * W), dtype=torch.int16, device=dev)
        import to_ued_b200
        self.precision = precision or to_ued_b200.GRU_PRECISION
        self.x = torch.empty((K, L, R, 8), dtype=f32, device=dev)
        if self.precision == "tc":
            f16 = torch.float16
            kk = K if keep_gates else 1                      # without a reverse pass one slot is enough
            Rp = (R + 63) // 64 * 64
            R32 = (R + 31) // 32 * 32                         # RB32 layout pads rows to blocks of 32
            self.h16 = torch.empty((kk, L, R32, 256), dtype=f16, device=dev)
            self.fac = torch.empty((kk, 5, L, R32, 256), dtype=f16, device=dev) if keep_gates else None
            # bf16 token-tile image of the masked carry (rows >= R of a partial 64-token block stay zero)
            self.hpimg = torch.zeros((kk, L * Rp * 256 * 2), dtype=torch.uint8, device=dev) if keep_gates else None
            # bf16 token-tile image of the LPG input rows (one 64-column group; columns 8..63 stay zero)
            self.ximg = torch.zeros((K, L * Rp * 128), dtype=torch.uint8, device=dev) if keep_gates else None
            self.wh_img = torch.empty(256 * 768, dtype=f16, device=dev)
            self.wh_img_version = None
            self.h = self.gates = None
        else:
            self.h = torch.empty((K, L, R, 256), dtype=f32, device=dev)
            self.gates = torch.empty((K, 4, L, R, 256), dtype=f32, device=dev) if keep_gates else None
        self.pi_hat = torch.empty((K, L, R), dtype=f32, device=dev)
        self.y_hat = torch.empty((K, L, R, 8), dtype=f32, device=dev)
        self.scalars = torch.empty((K, N, 8), dtype=f32, device=dev)
        self.step_in = torch.empty((K, N), dtype=i32, device=dev)
        self.ep_return = torch.empty((N, W), dtype=f32, device=dev)

    def transition(self, k) -> Transition:
        return Transition(self.obs[k], self.action[k], self.reward[k], self.done[k])


def lpg_agent_train_step(k, tape: Tape, levels, step, lpg_params, lifetime_conditioning, agent_target_coeff,
                         lr_actor, lr_critic, max_grad_norm):
    """agents/lpg_agent.py:31-85 for update ``k`` of the tape: reads tape.actor[k] / critic[k] and
    the k-th rollout, writes tape.actor[k+1] / critic[k+1], the LPG activations and scalars."""
    N, W, L, D = tape.N, tape.W, tape.L, tape.D
    s = _lib.stream_ptr()
    p = _lib.ptr
    _lib.call("toued_sort_tokens", p(tape.obs[k]), p(tape.sorted_tok[k]), N, W, L, s)
    tape.step_in[k].copy_(step)
    _lib.call("toued_lpg_prepare", p(tape.obs[k]), p(tape.action[k]), p(tape.reward[k]), p(tape.done[k]),
              p(tape.actor[k]), p(tape.critic[k]), p(lpg_params), p(step), p(levels), p(tape.x[k]),
              p(tape.ximg[k]) if getattr(tape, "ximg", None) is not None else None,
              N, W, L, D, int(lifetime_conditioning), s)
    if tape.precision == "tc":
        ks = k % tape.h16.shape[0]
        _lib.call("toued_gru_forward_tc", p(tape.x[k]), p(tape.done[k]), p(lpg_params), p(tape.wh_img),
                  p(tape.h16[ks]), p(tape.fac[ks]) if tape.fac is not None else None,
                  p(tape.hpimg[ks]) if tape.hpimg is not None else None, p(tape.pi_hat[k]), p(tape.y_hat[k]),
                  N, W, L, int(lifetime_conditioning), s)
    else:
        _lib.call("toued_gru_forward", p(tape.x[k]), p(tape.done[k]), p(lpg_params), p(tape.h[k]),
                  p(tape.gates[k]) if tape.gates is not None else None, p(tape.pi_hat[k]), p(tape.y_hat[k]),
                  N, W, L, int(lifetime_conditioning), s)
    _lib.call("toued_agent_update", p(tape.obs[k]), p(tape.action[k]), p(tape.sorted_tok[k]), p(tape.pi_hat[k]),
              p(tape.y_hat[k]), p(tape.actor[k]), p(tape.critic[k]), p(tape.actor[k + 1]), p(tape.critic[k + 1]),
              p(levels), p(step), p(tape.scalars[k]), N, W, L, D, float(lr_actor), float(lr_critic),
              float(max_grad_norm), float(agent_target_coeff), s)


def train_lpg_agent(rng, lpg_train_state, agent_state: AgentState, rollout_manager, num_train_steps: int,
                    agent_target_coeff: float, tape: Optional[Tape] = None):
    """agents/lpg_agent.py:88-140, batched: rng uint32[N, 2].
    Returns (agent_state, rollouts (list of K Transition), LPGAgentMetrics of f32[N])."""
    env = rollout_manager.env
    actor, critic = agent_state.actor_state, agent_state.critic_state
    N, W = agent_state.env_state.packed.shape
    L, K = rollout_manager.train_rollout_len, num_train_steps
    dev = actor.params.device
    if tape is None:
        tape = Tape(N, W, L, env.obs_dim, K, dev, keep_gates=False)
    levels = agent_state.level.packed
    step = actor.step.clone()
    state = agent_state.env_state.packed.clone()
    tape.actor[0].copy_(actor.params)
    tape.critic[0].copy_(critic.params)
    rng = np.asarray(rng, np.uint32).reshape(-1, 2)
    keys = np.empty((K, N, 2), np.uint32)
    for k in range(K):                                        # lpg_agent.py:104-105
        ks = prng.split(rng, 2)
        rng, keys[k] = ks[:, 0, :], ks[:, 1, :]
    keys_d = torch.from_numpy(keys.view(np.int32)).to(dev, non_blocking=True)
    s = _lib.stream_ptr()
    p = _lib.ptr
    lpg = lpg_train_state.params if hasattr(lpg_train_state, "params") else lpg_train_state
    cond = lpg_train_state.model.lifetime_conditioning if hasattr(lpg_train_state, "model") else (lpg.numel() > 204000)
    if tape.precision == "tc":                                # recurrent matrix -> fp16 SW128 pass images
        _lib.call("toued_pack_wh_forward", p(lpg), p(tape.wh_img), s)
    for k in range(K):
        _lib.call("toued_rollout", p(levels), p(keys_d[k]), p(tape.actor[k]), None, p(state), p(tape.obs[k]),
                  p(tape.action[k]), p(tape.reward[k]), p(tape.done[k]), None, N, W, L, env.obs_dim,
                  env.max_grid_size, env.max_n_objs, 0, s)
        lpg_agent_train_step(k, tape, levels, step, lpg, cond, agent_target_coeff, actor.learning_rate,
                             critic.learning_rate, actor.max_grad_norm)
    sc = tape.scalars[:K].mean(dim=0)                         # lpg_agent.py:140 mean over updates
    metrics = LPGAgentMetrics(policy_l2=sc[:, 4], policy_entropy=sc[:, 6], critic_loss=sc[:, 3],
                              critic_l2=sc[:, 5], critic_entropy=sc[:, 7])
    from ..environments.gridworld.gridworld import EnvState
    new_agent = agent_state.replace(
        actor_state=actor.replace(params=tape.actor[K].clone(), step=step),
        critic_state=critic.replace(params=tape.critic[K].clone(), step=step.clone()),
        env_obs=tape.obs[K - 1][:, -1].clone(), env_state=EnvState(state, env.max_n_objs))
    rollouts = [tape.transition(k) for k in range(K)]
    return new_agent, rollouts, metrics
