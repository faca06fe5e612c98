"""Rollout oracle (numpy).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates environments/rollout.py:38-102 (``RolloutWrapper.batch_reset`` / ``batch_rollout`` /
``single_rollout``) for the tabular actor of models/agent.py:7-17 with ``actor_net=()``: one
bias-free Dense + softmax on a one-hot-plus-time observation, i.e.

    logits = W[idx, :] + (time * 0.001) * W[D-1, :]          (obs @ W with two non-zeros)

Float contract for the *sampling* path (must be bit-identical to the CUDA rollout kernel so that
actions, transitions and dones can be compared bit for bit):
  * every op is an individually rounded IEEE f32 add / mul / div (no FMA contraction),
  * softmax uses ``exp_portable`` below (range reduction + degree-7 polynomial, <= 2 ulp),
  * cumsum is left to right,
  * action = #{ j : cumsum[j] < cumsum[-1] * (1 - u) }   (jax.random.choice with p).
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np

from . import prng
from .gridworld import GridWorld, EnvParams, EnvState

F32 = np.float32

_LOG2E = F32(1.4426950408889634)
_LN2_HI = F32(0.693359375)            # 355/512, exact in f32 with trailing zeros
_LN2_LO = F32(-2.12194440e-4)
_C = [F32(c) for c in (1.0, 1.0, 0.5, 1.6666667e-1, 4.1666668e-2, 8.3333338e-3, 1.3888889e-3, 1.9841270e-4)]
_XMIN = F32(-80.0)


def exp_portable(x):
    """exp(x) for x <= 0 using only individually-rounded f32 mul/add and integer exponent
    arithmetic; x is clamped to >= -80 (exp(-80) = 1.8e-35, far below f32 softmax resolution).
    Mirrors to_ued_b200/csrc/common.cuh::exp_portable."""
    x = np.maximum(np.asarray(x, F32), _XMIN)
    n = np.rint((x * _LOG2E).astype(F32)).astype(F32)
    r = (x - (n * _LN2_HI).astype(F32)).astype(F32)
    r = (r - (n * _LN2_LO).astype(F32)).astype(F32)
    p = np.full_like(r, _C[7])
    for c in (_C[6], _C[5], _C[4], _C[3], _C[2], _C[1], _C[0]):
        p = ((p * r).astype(F32) + c).astype(F32)
    bits = p.view(np.int32) + (n.astype(np.int32) << 23)
    return bits.view(F32)


def softmax_portable(z):
    z = np.asarray(z, F32)
    m = z.max(axis=-1, keepdims=True)
    e = exp_portable((z - m).astype(F32))
    s = np.zeros(e.shape[:-1], F32)
    for j in range(e.shape[-1]):
        s = (s + e[..., j]).astype(F32)
    return (e / s[..., None]).astype(F32)


def tab_logits(table, idx, time):
    """obs @ W for the one-hot-plus-time observation.  table: [N, D, C]; idx, time: [N, ...]"""
    N = table.shape[0]
    flat = idx.reshape(N, -1)
    rows = np.take_along_axis(table, flat[:, :, None], axis=1).reshape(idx.shape + (table.shape[-1],))
    tf = (time.astype(F32) * F32(0.001)).astype(F32)
    last = table[:, -1, :].reshape((N,) + (1,) * (idx.ndim - 1) + (table.shape[-1],))
    return (rows + (tf[..., None] * last).astype(F32)).astype(F32)


def actor_probs(table, idx, time):
    """models/agent.py:13-17 on the (idx, time) observation."""
    return softmax_portable(tab_logits(table, idx, time))


@dataclass
class Trajectory:
    """util/data.py:37-43 ``Transition`` in compact (idx, time) form, layout [N, L, W]."""
    obs_idx: np.ndarray      # i32[N, L+1, W]  obs_idx[:, t] is obs at t; [:, t+1] is next_obs at t
    obs_time: np.ndarray     # i32[N, L+1, W]
    action: np.ndarray       # i32[N, L, W]
    reward: np.ndarray       # f32[N, L, W]
    done: np.ndarray         # bool[N, L, W]


class RolloutWrapper:
    """environments/rollout.py:13-35"""

    def __init__(self, env: GridWorld, train_rollout_len=20, eval_rollout_len=None):
        self.env = env
        self.train_rollout_len = train_rollout_len
        self.eval_rollout_len = eval_rollout_len

    def batch_reset(self, rng, p: EnvParams, num_workers: int) -> EnvState:
        """rollout.py:38-42.  rng: uint32[N, 2].  The per-worker keys are dead in tabular mode."""
        return self.env.reset(None, p, num_workers)

    def batch_rollout(self, rng, actor_table, p: EnvParams, state: EnvState, eval=False,
                      forced_actions=None):
        """rollout.py:45-102.  rng: uint32[N, 2]; actor_table: f32[N, D, A].
        Returns (Trajectory, end_state, first_episode_return[N, W])."""
        env = self.env
        N, W = state.pos.shape
        L = self.eval_rollout_len if eval else self.train_rollout_len
        wk = prng.split(rng, W)                                  # rollout.py:49  [N, W, 2]
        obs_idx = np.zeros((N, L + 1, W), np.int32)
        obs_time = np.zeros((N, L + 1, W), np.int32)
        action = np.zeros((N, L, W), np.int32)
        reward = np.zeros((N, L, W), F32)
        done = np.zeros((N, L, W), bool)
        cum = np.zeros((N, W), F32)
        valid = np.ones((N, W), F32)
        s = state
        obs_idx[:, 0], obs_time[:, 0] = env.obs_index(s)
        for t in range(L):                                       # rollout.py:59-81
            ks = prng.split(wk, 2); wk, k_act = ks[..., 0, :], ks[..., 1, :]
            probs = actor_probs(actor_table, obs_idx[:, t], obs_time[:, t])
            a = prng.choice_p(k_act, probs)                      # rollout.py:63
            if forced_actions is not None:
                a = forced_actions[:, t].astype(np.int32)
            ks = prng.split(wk, 2); wk, k_env = ks[..., 0, :], ks[..., 1, :]
            s, r, d = env.step(k_env, s, a, p)                   # rollout.py:65
            cum = (cum + (r * valid).astype(F32)).astype(F32)    # rollout.py:68
            valid = (valid * (F32(1.0) - d.astype(F32))).astype(F32)
            action[:, t], reward[:, t], done[:, t] = a, r, d
            obs_idx[:, t + 1], obs_time[:, t + 1] = env.obs_index(s)
        return Trajectory(obs_idx, obs_time, action, reward, done), s, cum
