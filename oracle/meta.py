"""Meta-optimiser oracle (torch, CPU).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates meta/train.py:14-130 ``lpg_meta_grad_train_step`` (meta-gradient through the K agent
updates, by torch.autograd where the reference uses jax.grad), meta/train.py:133-227
``lpg_es_train_step`` and optax==0.1.5 ``scale_by_adam`` [3P-recall].

Reproduced quirks (SURVEY.md §2.1 + one found while restating):
  Q2   the value critic is never trained: ``value_critic_state.replace(params=...)`` is discarded
       (meta/train.py:62), so its gradient is structurally zero and advantages come from the
       initial critic; only its step counter advances (K + 1 per meta-step).
  Q11  eval_agent uses 4 workers on the meta-gradient path.
  Q16  ``compute_advantage`` returns adv with a trailing axis of size 1 ([L, 1] per worker,
       agents.py:109-116) and ``-jnp.multiply(sampled_log_probs, adv)`` (meta/train.py:92)
       broadcasts [L] x [L, 1] to the OUTER product [L, L]; its mean is
       ``-mean_t(log pi) * mean_t(adv)`` per worker.  ``outer_product_quirk=False`` gives the
       element-wise product instead (not the reference's behaviour).
"""
from __future__ import annotations

import numpy as np
import torch

from . import prng
from .agents import (AgentTables, Hypers, train_lpg_agent, eval_agent, compute_advantage, tab_forward)
from .lpg import LPGLayout
from .rollout import RolloutWrapper


def lpg_meta_grad_train_step(rng, layout: LPGLayout, lpg_flat, ag: AgentTables, value_tables, ro: RolloutWrapper,
                             env_params, env_state, lifetime, *, num_agent_updates=5, gamma=0.99, gae_lambda=0.95,
                             alpha_y=0.5, beta=(5e-2, 1e-3, 5e-3, 1e-3), hy: Hypers = Hypers(),
                             trajectories=None, eval_trajectory=None, outer_product_quirk=True,
                             eval_workers=4, do_eval=True):
    """meta/train.py:14-130 for a batch of N agents.  rng: the scalar key (uint32[2]) of the step.
    beta = (policy_entropy, target_entropy, policy_l2, target_l2) coefficients.
    Returns dict(grad=mean LPG grad [P], per_agent_grad [N, P] (optional), agents, env_state,
    metrics, trajectories, eval_trajectory, adv)."""
    N = ag.actor.shape[0]
    dt = lpg_flat.dtype
    lpg_flat = lpg_flat.detach().clone().requires_grad_(True)
    rngs = prng.split(rng, N)                                                # train.py:120
    ks = prng.split(rngs, 2); rngs, r_train = ks[:, 0, :], ks[:, 1, :]       # :40
    ag0 = AgentTables(ag.actor.detach().clone().requires_grad_(True),
                      ag.critic.detach().clone().requires_grad_(True), ag.step.clone())
    agK, env_state, rollouts, am, dbg = train_lpg_agent(
        r_train, layout, lpg_flat, ag0, ro, env_params, env_state, lifetime, num_agent_updates, alpha_y, hy,
        trajectories=trajectories)
    ks = prng.split(rngs, 2); rngs, r_eval = ks[:, 0, :], ks[:, 1, :]        # :47
    if eval_trajectory is None:
        eval_traj, env_state, _ = ro.batch_rollout(r_eval, agK.actor.detach().to(torch.float32).numpy(),
                                                   env_params, env_state)
    else:
        eval_traj = eval_trajectory
    # value "update" (Q2: params never change) and advantage          train.py:60-85
    value_loss_w, adv = compute_advantage(value_tables.detach(), eval_traj, gamma, gae_lambda)
    value_loss = value_loss_w.mean(1)
    a = adv.flatten(1)
    adv = (adv - a.mean(1)[:, None, None]) / (a.std(1, unbiased=False)[:, None, None] + 1e-8)
    L = eval_traj.action.shape[1]
    probs = tab_forward(agK.actor, eval_traj.obs_idx[:, :L], eval_traj.obs_time[:, :L])
    action = torch.as_tensor(eval_traj.action.astype(np.int64))
    logp = torch.gather(torch.log(probs + 1e-8), -1, action[..., None])[..., 0]           # [N, L, W]
    if outer_product_quirk:
        lpg_loss = -(logp.mean(1) * adv.mean(1)).mean(1)                                   # Q16
    else:
        lpg_loss = -(logp * adv).flatten(1).mean(1)
    b0, b1, b2, b3 = beta
    reg = (lpg_loss - b0 * am["policy_entropy"] + b2 * am["policy_l2"]
           - b1 * am["critic_entropy"] + b3 * am["critic_l2"])                             # :94-100
    grad, = torch.autograd.grad(reg.sum(), lpg_flat)
    grad = grad / N                                                                        # :128 mean over agents
    metrics = {"lpg_loss": lpg_loss.detach().mean(), "reg_lpg_loss": reg.detach().mean(),
               "value_loss": value_loss.mean(),
               "lpg_agent": {k: v.detach().mean() for k, v in am.items()}}
    per_agent = {"lpg_loss": lpg_loss.detach(), "reg_lpg_loss": reg.detach(), "value_loss": value_loss,
                 **{k: v.detach() for k, v in am.items()}}
    if do_eval:
        ks = prng.split(rngs, 2)                                                           # :109
        ret = eval_agent(ks[:, 1, :], ro, env_params, agK.actor, eval_workers)
        metrics["lpg_agent_return"] = float(ret.mean())
        per_agent["lpg_agent_return"] = ret
    agents_out = AgentTables(agK.actor.detach(), agK.critic.detach(), agK.step)
    return dict(grad=grad, agents=agents_out, env_state=env_state, metrics=metrics, per_agent=per_agent,
                trajectories=rollouts, eval_trajectory=eval_traj, adv=adv, debug=dbg)


class SGDClip:
    """optax.chain(clip_by_global_norm(max_norm), scale(lr), scale(-1)) (models/optim.py:6-11, ``--lpg_opt SGD``);
    optax.clip_by_global_norm: g if |g| < max_norm else g / |g| * max_norm."""

    def __init__(self, lr, max_norm):
        self.lr, self.max_norm = lr, max_norm

    def step(self, params, grad):
        norm = torch.sqrt((grad * grad).sum())
        g = grad if float(norm) < self.max_norm else grad / norm * self.max_norm
        return params - self.lr * g


class Adam:
    """optax.scale_by_adam(b1=.9, b2=.999, eps=1e-8, eps_root=0) -> scale(lr) -> scale(-1)
    (models/optim.py:12-17; no clipping on this branch, Q9)."""

    def __init__(self, size, lr, dtype=torch.float32):
        self.mu = torch.zeros(size, dtype=dtype)
        self.nu = torch.zeros(size, dtype=dtype)
        self.count = 0
        self.lr = lr

    def step(self, params, grad, b1=0.9, b2=0.999, eps=1e-8):
        self.mu = b1 * self.mu + (1 - b1) * grad
        self.nu = b2 * self.nu + (1 - b2) * grad * grad
        self.count += 1
        mu_hat = self.mu / (1 - b1 ** self.count)
        nu_hat = self.nu / (1 - b2 ** self.count)
        return params - self.lr * mu_hat / (torch.sqrt(nu_hat) + eps)
