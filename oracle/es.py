"""OpenES + the ES meta-step oracle.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates evosax==0.1.4 ``OpenES`` with the Adam ``GradientOptimizer`` [3P-recall] as used by
models/optim.py:21-34, and meta/train.py:133-227 ``lpg_es_train_step`` (antithetic task sampling).

evosax semantics restated: ``ask``: z_plus ~ N(0, I) [popsize/2, P]; x = mean + sigma * [z_plus; -z_plus].
``tell`` (maximize=True -> fitness negated, no other shaping): noise = (x - mean) / sigma;
theta_grad = noise^T fitness / (popsize * sigma); Adam(beta1 .99, beta2 .999, eps 1e-8) with bias
correction exponent gen_counter + 1; lrate <- max(lrate * lrate_decay, lrate_limit);
sigma <- max(sigma * sigma_decay, sigma_limit); mean <- (1 - mean_decay) * mean.
``initialize``: mean = uniform(init_min = 0, init_max = 0) = zeros (so ES starts from zero LPG parameters;
train_state.params is never refreshed, Q12)."""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
import torch

from . import prng
from .agents import AgentTables, Hypers, train_lpg_agent, eval_agent


@dataclass
class ESState:
    mean: torch.Tensor
    sigma: float
    m: torch.Tensor
    v: torch.Tensor
    lrate: float
    gen_counter: int = 0


def es_init(P, lrate_init, sigma_init, dtype=torch.float64):
    z = torch.zeros(P, dtype=dtype)
    return ESState(z.clone(), sigma_init, z.clone(), z.clone(), lrate_init)


def normal(key, shape):
    """jax.random.normal: sqrt(2) * erfinv(uniform(nextafter(-1, 0), 1))"""
    from scipy.special import erfinv
    u = prng.uniform(key, shape, np.nextafter(np.float32(-1), np.float32(0)), 1.0)
    return np.sqrt(2.0) * erfinv(u.astype(np.float64))


def es_ask(key, st: ESState, popsize):
    z = torch.tensor(normal(key, (popsize // 2, st.mean.numel())), dtype=st.mean.dtype)
    x = st.mean + st.sigma * torch.cat([z, -z])
    return x


def es_tell(x, fitness, st: ESState, popsize, lrate_decay=0.999, lrate_limit=1e-5, sigma_decay=1.0, sigma_limit=0.1,
            mean_decay=0.0, b1=0.99, b2=0.999, eps=1e-8):
    fit = -fitness.to(x.dtype)
    noise = (x - st.mean) / st.sigma
    g = (noise.T @ fit) / (popsize * st.sigma)
    m = (1 - b1) * g + b1 * st.m
    v = (1 - b2) * g * g + b2 * st.v
    mhat = m / (1 - b1 ** (st.gen_counter + 1))
    vhat = v / (1 - b2 ** (st.gen_counter + 1))
    mean = st.mean - st.lrate * mhat / (torch.sqrt(vhat) + eps)
    mean = mean * (1 - mean_decay)
    return ESState(mean, max(st.sigma * sigma_decay, sigma_limit), m, v, max(st.lrate * lrate_decay, lrate_limit),
                   st.gen_counter + 1)


def lpg_es_train_step(rng, layout, st: ESState, ag: AgentTables, ro, env_params, env_state, lifetime, *,
                      num_agent_updates, alpha_y=0.5, hy: Hypers = Hypers(), env_workers=64, candidates=None,
                      trajectories=None, fitness_override=None):
    """meta/train.py:133-227 for N agents (popsize 2N).  ``candidates`` (pair-adjacent order) and
    ``trajectories`` (K Trajectory objects over the 2N repeated agents) can be injected so that the float
    path is compared on identical data.  Returns dict(fitness, rank_fitness, first_greater, agents, es_state)."""
    N = ag.actor.shape[0]
    popsize = 2 * N
    ks = prng.split(rng, 2); rng, k_ask = ks[0], ks[1]
    if candidates is None:
        x = es_ask(k_ask, st, popsize)
        idx = np.stack([np.arange(N), np.arange(N) + N], 1).reshape(-1)        # train.py:152-158
        candidates = x[idx]
    rep = lambda a: np.repeat(a, 2, axis=0)
    rag = AgentTables(ag.actor.repeat_interleave(2, 0).detach().requires_grad_(True),
                      ag.critic.repeat_interleave(2, 0).detach().requires_grad_(True), ag.step.repeat_interleave(2, 0))
    p2 = env_params.index(np.repeat(np.arange(N), 2))
    from .gridworld import EnvState
    es2 = EnvState(*[rep(getattr(env_state, f)) for f in ("time", "pos", "obj_poss", "obj_existss", "early_term")])
    ks = prng.split(rng, 2); rng, k = ks[0], ks[1]
    keys = prng.split(k, popsize)
    # per-candidate parameters: train each repeated agent with its own LPG vector
    fitness = np.zeros(popsize, np.float32)
    actors, critics, steps = [], [], []
    for i in range(popsize):
        sl = slice(i, i + 1)
        agi = AgentTables(rag.actor[sl], rag.critic[sl], rag.step[sl])
        ki = prng.split(keys[i:i + 1], 2)
        kk, k_train = ki[:, 0, :], ki[:, 1, :]
        esi = EnvState(*[getattr(es2, f)[sl] for f in ("time", "pos", "obj_poss", "obj_existss", "early_term")])
        tr = None if trajectories is None else [type(t)(t.obs_idx[sl], t.obs_time[sl], t.action[sl], t.reward[sl], t.done[sl]) for t in trajectories]
        out, _, _, _, _ = train_lpg_agent(k_train, layout, candidates[i].detach(), agi, ro, p2.index(sl), esi,
                                          lifetime[np.repeat(np.arange(N), 2)][sl], num_agent_updates, alpha_y, hy,
                                          trajectories=tr)
        actors.append(out.actor.detach()); critics.append(out.critic.detach()); steps.append(out.step)
        if fitness_override is None:
            fitness[i] = eval_agent(kk, ro, p2.index(sl), out.actor, env_workers)[0]     # rng (not _rng), train.py:182-189
    if fitness_override is not None:
        fitness = np.asarray(fitness_override, np.float32)
    first_greater = fitness[::2] > fitness[1::2]
    rank = np.zeros(popsize, np.float32)
    rank[::2] = first_greater; rank[1::2] = 1.0 - first_greater
    A, C, S = torch.cat(actors), torch.cat(critics), torch.cat(steps)
    sel = torch.tensor(np.where(first_greater, np.arange(N) * 2, np.arange(N) * 2 + 1))
    new_st = es_tell(candidates, torch.tensor(rank), st, popsize)
    return dict(fitness=fitness, rank_fitness=rank, first_greater=first_greater, candidates=candidates,
                agents=AgentTables(A[sel], C[sel], S[sel]), all_actor=A, es_state=new_st)
