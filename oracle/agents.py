"""Agent-side oracle (torch, CPU, differentiable).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates, batched over a leading agent axis N:
  models/agent.py:7-45          Actor / Critic with actor_net = critic_net = ()
  models/optim.py:5-11          SGD = clip_by_global_norm -> scale(lr) -> scale(-1)   [optax 0.1.5]
  util/metrics.py:5-38          batch_rollout_entropy, kl_divergence, gae
  agents/lpg_agent.py:31-140    lpg_agent_train_step, train_lpg_agent
  agents/a2c.py:19-125          a2c_agent_train_step, train_a2c_agent
  agents/agents.py:98-116       eval_agent, compute_advantage

Tables are ``[N, D, C]`` tensors; an observation is (row index, time) and
``obs @ W == W[row] + 0.001 * time * W[D-1]`` (two non-zeros of the one-hot-plus-time vector).
Everything that the reference differentiates with ``jax.grad`` is differentiated here with
``torch.autograd`` (``create_graph=True`` so the meta-gradient flows through the updates).
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
import torch

from . import prng
from .lpg import LPGLayout, lpg_forward
from .rollout import RolloutWrapper, Trajectory


def tab_forward(table, obs_idx, obs_time, softmax=True):
    """table [N, D, C]; obs_idx/time integer arrays [N, ...] -> [N, ..., C]"""
    N, D, C = table.shape
    idx = torch.as_tensor(np.asarray(obs_idx), dtype=torch.long)
    tf = (torch.as_tensor(np.asarray(obs_time)).to(torch.float32) * np.float32(0.001)).to(table.dtype)
    flat = idx.reshape(N, -1)
    rows = torch.gather(table, 1, flat[:, :, None].expand(-1, -1, C)).reshape(idx.shape + (C,))
    last = table[:, D - 1, :].reshape((N,) + (1,) * (idx.dim() - 1) + (C,))
    z = rows + tf[..., None] * last
    if not softmax:
        return z
    return torch.softmax(z, dim=-1)


def clip_sgd(param, grad, lr, max_norm):
    """optax.chain(clip_by_global_norm(max_norm), scale(lr), scale(-1)) applied per agent
    (each agent's network has its own TrainState, agents.py:92-95)."""
    g_norm = grad.flatten(1).norm(dim=1)
    trigger = g_norm < max_norm
    safe = torch.where(trigger, torch.ones_like(g_norm), g_norm)
    scale = torch.where(trigger, torch.ones_like(g_norm), max_norm / safe)
    return param - lr * grad * scale[:, None, None]


def entropy(table, obs_idx, obs_time):
    """util/metrics.py:5-9 per agent: -mean(sum((p + 1e-8) log(p + 1e-8)))"""
    p = tab_forward(table, obs_idx, obs_time) + 1e-8
    return -(p * torch.log(p)).sum(-1).flatten(1).mean(1)


def gae(value, reward, done, discount, lam):
    """util/metrics.py:17-38 along axis 1 (time).  value [N, L+1, W]; reward/done [N, L, W]"""
    L = reward.shape[1]
    adv = [None] * L
    g = torch.zeros_like(value[:, 0])
    for t in reversed(range(L)):
        nd = 1.0 - done[:, t]
        delta = reward[:, t] + discount * value[:, t + 1] * nd - value[:, t]
        g = delta + discount * lam * nd * g
        adv[t] = g
    adv = torch.stack(adv, 1)
    return adv, adv + value[:, :-1]


@dataclass
class AgentTables:
    actor: torch.Tensor        # [N, D, 5]
    critic: torch.Tensor       # [N, D, Y]
    step: torch.Tensor         # int64 [N]   (actor_state.step; critic step advances identically)


@dataclass
class Hypers:
    actor_lr: float = 4e1
    critic_lr: float = 4e0
    max_grad_norm: float = 0.5


def _traj_t(traj: Trajectory, dt):
    return (torch.as_tensor(traj.reward).to(dt), torch.as_tensor(traj.done).to(dt),
            torch.as_tensor(traj.action.astype(np.int64)))


def lpg_agent_train_step(layout: LPGLayout, lpg_flat, ag: AgentTables, traj: Trajectory, lifetime,
                         alpha_y: float, hy: Hypers):
    """agents/lpg_agent.py:31-85.  Returns (new AgentTables, critic_loss[N], pi_l2[N], y_l2[N])."""
    dt = ag.actor.dtype
    N, L, W = traj.action.shape
    reward, done, action = _traj_t(traj, dt)
    oi, ot = traj.obs_idx, traj.obs_time

    probs = tab_forward(ag.actor, oi[:, :L], ot[:, :L])
    pi = torch.gather(probs + 1e-8, -1, action[..., None])[..., 0]          # [N, L, W]
    y_t = tab_forward(ag.critic, oi[:, :L], ot[:, :L])
    y_tp1 = tab_forward(ag.critic, oi[:, 1:], ot[:, 1:])
    seq = lambda a: a.transpose(1, 2).reshape((N * W, L) + a.shape[3:])      # [N, L, W, ..] -> [N*W, L, ..]
    lifetime_t = torch.as_tensor(np.asarray(lifetime))
    pi_hat, y_hat = lpg_forward(layout, lpg_flat, seq(reward), seq(done), seq(pi.detach()), seq(y_t.detach()),
                                seq(y_tp1.detach()), ag.step.repeat_interleave(W), lifetime_t.repeat_interleave(W))
    unseq = lambda a: a.reshape((N, W, L) + a.shape[2:]).transpose(1, 2)
    pi_hat, y_hat = unseq(pi_hat), unseq(y_hat)
    y_l2 = (y_hat ** 2).sum(-1).flatten(1).mean(1)
    critic_loss = (y_t * (torch.log(y_t + 1e-8) - torch.log(y_hat + 1e-8))).sum(-1)   # metrics.py:12-14
    actor_loss = torch.log(pi) * pi_hat
    pi_l2 = (pi_hat ** 2).flatten(1).mean(1)
    loss = actor_loss.flatten(1).mean(1) + alpha_y * critic_loss.flatten(1).mean(1)
    ga, gc = torch.autograd.grad(loss.sum(), (ag.actor, ag.critic), create_graph=True)
    new_actor = clip_sgd(ag.actor, ga, hy.actor_lr, hy.max_grad_norm)
    new_critic = clip_sgd(ag.critic, gc, hy.critic_lr, hy.max_grad_norm)
    new_step = ag.step + 1
    keep = (new_step <= lifetime_t)                                          # lpg_agent.py:78-82
    k3 = keep[:, None, None]
    out = AgentTables(torch.where(k3, new_actor, ag.actor), torch.where(k3, new_critic, ag.critic),
                      torch.where(keep, new_step, ag.step))
    return out, critic_loss.flatten(1).mean(1), pi_l2, y_l2, (ga, gc, pi_hat, y_hat)


def train_lpg_agent(rng, layout, lpg_flat, ag: AgentTables, ro: RolloutWrapper, env_params, env_state,
                    lifetime, num_train_steps, alpha_y, hy: Hypers, trajectories=None):
    """agents/lpg_agent.py:88-140.  rng: uint32[N, 2].  ``trajectories`` (list of K Trajectory)
    replays given rollouts instead of sampling (used for float-tolerance parity, where a 1-ulp
    table difference must not change the sampled data).
    Returns (AgentTables, env_state, [Trajectory]*K, metrics dict of [N] tensors, per-step debug)."""
    rollouts, mets, dbg = [], [], []
    for k in range(num_train_steps):
        ks = prng.split(rng, 2); rng, _rng = ks[:, 0, :], ks[:, 1, :]
        if trajectories is None:
            table32 = ag.actor.detach().to(torch.float32).numpy()
            traj, env_state, _ = ro.batch_rollout(_rng, table32, env_params, env_state)
        else:
            traj = trajectories[k]
        ag, critic_loss, pi_l2, y_l2, aux = lpg_agent_train_step(layout, lpg_flat, ag, traj, lifetime, alpha_y, hy)
        L = traj.action.shape[1]
        a_ent = entropy(ag.actor, traj.obs_idx[:, :L], traj.obs_time[:, :L])      # lpg_agent.py:119-120
        c_ent = entropy(ag.critic, traj.obs_idx[:, :L], traj.obs_time[:, :L])
        rollouts.append(traj)
        mets.append(dict(policy_l2=pi_l2, policy_entropy=a_ent, critic_loss=critic_loss, critic_l2=y_l2,
                         critic_entropy=c_ent))
        dbg.append(dict(actor=ag.actor, critic=ag.critic, step=ag.step, ga=aux[0], gc=aux[1], pi_hat=aux[2], y_hat=aux[3]))
    metrics = {k: torch.stack([m[k] for m in mets]).mean(0) for k in mets[0]}
    return ag, env_state, rollouts, metrics, dbg


def eval_agent(rng, ro: RolloutWrapper, env_params, actor_table, num_workers):
    """agents/agents.py:98-106.  rng: uint32[N, 2] -> mean first-episode return [N] (numpy f32)."""
    ks = prng.split(rng, 2); rng = ks[:, 0, :]
    s0 = ro.batch_reset(ks[:, 1, :], env_params, num_workers)
    ks = prng.split(rng, 2)
    table32 = actor_table.detach().to(torch.float32).numpy() if torch.is_tensor(actor_table) else actor_table
    _, _, ret = ro.batch_rollout(ks[:, 1, :], table32, env_params, s0, eval=True)
    return ret.mean(axis=1, dtype=np.float32)


def compute_advantage(value_table, traj: Trajectory, gamma, lam):
    """agents/agents.py:109-116 per agent & worker.  value_table [N, D, 1].
    Returns (mse [N, W], adv [N, L, W]).  NOTE (Q16): in the reference the value critic output
    keeps a trailing axis of size 1, so the reference's ``adv`` is [W, L, 1]; consumers that
    multiply it with a [L] vector get an outer product (see meta.py / a2c)."""
    dt = value_table.dtype
    reward, done, _ = _traj_t(traj, dt)
    value = tab_forward(value_table, traj.obs_idx, traj.obs_time, softmax=False)[..., 0]   # [N, L+1, W]
    adv, target = gae(value, reward, done, gamma, lam)
    adv, target = adv.detach(), target.detach()
    return ((target - value[:, :-1]) ** 2).mean(1), adv


# ------------------------------------------------------------------------------------------- A2C
@dataclass
class A2CHyperparams:
    gamma: float = 0.99
    gae_lambda: float = 0.95
    entropy_coeff: float = 0.01


def a2c_agent_train_step(actor, critic, step, traj: Trajectory, lifetime, hypers: A2CHyperparams, hy: Hypers):
    """agents/a2c.py:19-76.  actor [N, D, 5], critic [N, D, 1] (value critic)."""
    dt = actor.dtype
    N, L, W = traj.action.shape
    reward, done, action = _traj_t(traj, dt)
    critic = critic.detach().requires_grad_(True)
    actor = actor.detach().requires_grad_(True)
    value = tab_forward(critic, traj.obs_idx, traj.obs_time, softmax=False)[..., 0]
    adv, target = gae(value, reward, done, hypers.gamma, hypers.gae_lambda)
    adv, target = adv.detach(), target.detach()
    closs = ((target - value[:, :-1]) ** 2).mean(1).mean(1)                    # mean over t, then workers
    a = adv.flatten(1)
    adv = (adv - a.mean(1)[:, None, None]) / (a.std(1, unbiased=False)[:, None, None] + 1e-8)
    gc, = torch.autograd.grad(closs.sum(), critic)
    new_critic = clip_sgd(critic, gc, hy.critic_lr, hy.max_grad_norm)
    p = tab_forward(actor, traj.obs_idx[:, :L], traj.obs_time[:, :L]) + 1e-8
    logp = torch.log(p)
    sel = torch.gather(logp, -1, action[..., None])[..., 0]
    ent = -(p * logp).sum(-1).mean(1)                                          # per worker
    # Quirk Q16 (reproduced): the value critic's output keeps its trailing axis of size 1, so adv
    # is [L, 1] per worker and ``-jnp.multiply(selected_log_probs, adv)`` (a2c.py:60) broadcasts
    # [L] x [L, 1] to the OUTER product [L, L]; its mean is -mean_t(logp) * mean_t(adv).
    aloss = (-(sel.mean(1) * adv.mean(1)) - hypers.entropy_coeff * ent).mean(1)
    ga, = torch.autograd.grad(aloss.sum(), actor)
    new_actor = clip_sgd(actor, ga, hy.actor_lr, hy.max_grad_norm)
    lifetime_t = torch.as_tensor(np.asarray(lifetime))
    new_step = step + 1
    keep = new_step <= lifetime_t
    k3 = keep[:, None, None]
    return (torch.where(k3, new_actor, actor).detach(), torch.where(k3, new_critic, critic).detach(),
            torch.where(keep, new_step, step), aloss.detach(), closs.detach())


def train_a2c_agent(rng, actor, critic, step, ro, env_params, env_state, lifetime, num_train_steps,
                    hypers: A2CHyperparams, hy: Hypers, trajectories=None):
    """agents/a2c.py:79-125"""
    al, cl = [], []
    for k in range(num_train_steps):
        ks = prng.split(rng, 2); rng, _rng = ks[:, 0, :], ks[:, 1, :]
        if trajectories is None:
            traj, env_state, _ = ro.batch_rollout(_rng, actor.detach().to(torch.float32).numpy(), env_params, env_state)
        else:
            traj = trajectories[k]
        actor, critic, step, a, c = a2c_agent_train_step(actor, critic, step, traj, lifetime, hypers, hy)
        al.append(a); cl.append(c)
    return actor, critic, step, env_state, dict(actor_loss=torch.stack(al).mean(0), critic_loss=torch.stack(cl).mean(0))
