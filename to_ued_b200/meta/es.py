"""ES meta-optimisation of LPG (TA-LPG): reference meta/train.py:133-227 ``lpg_es_train_step``,
models/optim.py:21-34 ``create_es_strategy`` (evosax==0.1.4 OpenES, restated — [3P-recall]) and
util/data.py:62-68 ``ESTrainState``.

Antithetic task sampling: candidate 2i and 2i+1 are mean +/- sigma z_i and both train a copy of agent i on
the same level; fitness is the rank within the pair.  Every candidate has its own LPG parameter vector, so
the LPG forward runs with a per-agent parameter stride (one CTA per agent in the exact-fp32 GRU kernel).

Multi-GPU: rank r trains the pairs of its own agents (``toued_es_ask_shard`` draws exactly
the global population's candidates for them), the pairwise rank fitness is local to a pair, the gradient estimate
``noise^T fitness`` is summed over ranks with one all-reduce, the raw fitness is all-gathered for the metrics, and every
rank applies the identical Adam step to its replica of the ES mean.

Reproduced: Q11 (fitness from ``env_workers`` eval workers), Q12 (``train_state.params`` is never refreshed;
the ES mean — which evosax initialises to zeros — is the result; exposed as ``ESTrainState.mean``)."""
from __future__ import annotations

from dataclasses import dataclass, replace as _replace

import numpy as np
import torch

from .. import _lib
from ..util import prng
from ..util import dist as udist
from ..util.data import AgentState, LpgHyperparams, Level
from ..agents.lpg_agent import train_lpg_agent
from ..agents.agents import eval_agent
from ..environments.gridworld.gridworld import EnvState


@dataclass
class OpenES:
    """models/optim.py:21-34 arguments of evosax.OpenES."""
    popsize: int
    num_dims: int
    opt_name: str = "adam"
    lrate_init: float = 1e-4
    lrate_decay: float = 0.999
    lrate_limit: float = 1e-5
    sigma_init: float = 0.1
    sigma_decay: float = 1.0
    sigma_limit: float = 0.1
    mean_decay: float = 0.0
    maximize: bool = True
    beta_1: float = 0.99
    beta_2: float = 0.999
    eps: float = 1e-8

    def __post_init__(self):
        if self.popsize & 1:
            raise ValueError("Population size must be even")
        if self.opt_name != "adam":
            raise NotImplementedError("only the Adam gradient optimizer of evosax is implemented (reference default)")


@dataclass
class ESTrainState:
    """util/data.py:62-68"""
    train_state: object          # LPGTrainState (params never refreshed, Q12)
    strategy: OpenES
    es_params: object
    es_state: dict               # mean, sigma, m, v, lrate, gen_counter

    @property
    def mean(self):
        return self.es_state["mean"]

    def replace(self, **kw):
        return _replace(self, **kw)


def create_es_strategy(args, n_params):
    return OpenES(popsize=args.num_agents * 2, num_dims=n_params, opt_name=args.lpg_opt.lower(),
                  lrate_init=args.lpg_learning_rate, lrate_decay=args.es_lrate_decay, lrate_limit=args.es_lrate_limit,
                  sigma_init=args.es_sigma_init, sigma_decay=args.es_sigma_decay, sigma_limit=args.es_sigma_limit,
                  mean_decay=args.es_mean_decay)


def create_es_train_state(rng, args, train_state):
    strategy = create_es_strategy(args, train_state.params.numel())
    z = torch.zeros_like(train_state.params)
    es_state = {"mean": z.clone(), "sigma": strategy.sigma_init, "m": z.clone(), "v": z.clone(),
                "lrate": strategy.lrate_init, "gen_counter": 0}
    return ESTrainState(train_state, strategy, None, es_state)


def _repeat2(agents: AgentState, W, max_n_objs) -> AgentState:
    """jnp.repeat(x, 2, axis=0) over the agent axis (meta/train.py:193-195)."""
    r = lambda t: t.repeat_interleave(2, dim=0).contiguous()
    a, c, lv = agents.actor_state, agents.critic_state, agents.level
    level = Level(lv.env_params, np.repeat(lv.lifetime, 2), np.repeat(lv.buffer_id, 2), r(lv.packed))
    return AgentState(a.replace(params=r(a.params), step=r(a.step)), c.replace(params=r(c.params), step=r(c.step)),
                      level, r(agents.env_obs), EnvState(r(agents.env_state.packed), max_n_objs),
                      None if agents.host_step is None else np.repeat(agents.host_step, 2))


def lpg_es_train_step(rng, lpg_train_state: ESTrainState, agent_states: AgentState, value_critic_states,
                      rollout_manager, num_mini_batches: int, lpg_hypers: LpgHyperparams, *, candidates=None):
    """Train a batch of agents with LPG candidates, then update LPG with ES (meta/train.py:133-227).
    Returns (lpg_train_state, agent_states, None, metrics)."""
    env = rollout_manager.env
    strat, st = lpg_train_state.strategy, lpg_train_state.es_state
    N, W = agent_states.env_state.packed.shape                  # local agents
    rank_, world = udist.rank_world()
    pop_global, P = strat.popsize, strat.num_dims
    popsize = 2 * N                                             # local members
    if pop_global != popsize * world:
        raise ValueError(f"ES population ({pop_global}) must be twice the number of agents ({N} x {world} ranks)")
    pair_off = rank_ * N
    dev = agent_states.actor_state.params.device
    p, s = _lib.ptr, _lib.stream_ptr()
    rng = np.asarray(rng, np.uint32)
    rng, k_ask = prng.split(rng, 2)
    Pp = (P + 3) // 4 * 4                                   # 16-byte aligned candidate rows
    cand = torch.zeros((popsize, Pp), dtype=torch.float32, device=dev)
    if candidates is None:
        kd = torch.from_numpy(np.ascontiguousarray(k_ask).view(np.int32)).to(dev)
        if world == 1:
            _lib.call("toued_es_ask", p(kd), p(st["mean"]), float(st["sigma"]), p(cand), popsize, P, Pp, s)
        else:
            _lib.call("toued_es_ask_shard", p(kd), p(st["mean"]), float(st["sigma"]), p(cand), pop_global, P, Pp,
                      pair_off, N, s)
    else:
        cand[:, :P] = candidates
    rep = _repeat2(agent_states, W, env.max_n_objs)
    rng, k = prng.split(rng, 2)
    keys = prng.split(k, pop_global)[2 * pair_off:2 * pair_off + popsize]
    ks = prng.split(keys, 2)
    k_eval, k_train = ks[:, 0, :], ks[:, 1, :]                  # rng, _rng = split(rng) (meta/train.py:172)

    class _Cand:                                                # per-candidate "train state"
        params = cand
        model = lpg_train_state.train_state.model
    if popsize % num_mini_batches != 0:
        raise ValueError("population must be divisible by num_mini_batches")
    nb = popsize // num_mini_batches
    fitness = torch.empty(popsize, dtype=torch.float32, device=dev)
    new_actor = torch.empty_like(rep.actor_state.params)
    new_critic = torch.empty_like(rep.critic_state.params)
    new_step = torch.empty_like(rep.actor_state.step)
    new_state = torch.empty_like(rep.env_state.packed)
    new_obs = torch.empty_like(rep.env_obs)
    msum = torch.zeros(5, dtype=torch.float32, device=dev)
    for mb in range(num_mini_batches):
        sl = slice(mb * nb, (mb + 1) * nb)
        lv = rep.level
        sub = AgentState(rep.actor_state.replace(params=rep.actor_state.params[sl], step=rep.actor_state.step[sl]),
                         rep.critic_state.replace(params=rep.critic_state.params[sl], step=rep.critic_state.step[sl]),
                         Level(lv.env_params, lv.lifetime[sl], lv.buffer_id[sl], lv.packed[sl]), rep.env_obs[sl],
                         EnvState(rep.env_state.packed[sl], env.max_n_objs))

        class _Sub:
            params = cand[sl]
            model = lpg_train_state.train_state.model
        out, _, am = train_lpg_agent(k_train[sl], _Sub, sub, rollout_manager, lpg_hypers.num_agent_updates,
                                     lpg_hypers.agent_target_coeff, lpg_stride=Pp)
        fitness[sl] = eval_agent(k_eval[sl], rollout_manager, sub.level.packed, out.actor_state, W)
        new_actor[sl], new_critic[sl], new_step[sl] = out.actor_state.params, out.critic_state.params, out.actor_state.step
        new_state[sl], new_obs[sl] = out.env_state.packed, out.env_obs
        msum += torch.stack([am.policy_l2.sum(), am.policy_entropy.sum(), am.critic_loss.sum(), am.critic_l2.sum(),
                             am.critic_entropy.sum()])
    # --- rank transformation per antithetic pair, keep the better agent (meta/train.py:203-211) ---
    first_greater = fitness[0::2] > fitness[1::2]
    rank = torch.empty_like(fitness)
    rank[0::2] = first_greater.float()
    rank[1::2] = 1.0 - first_greater.float()
    pick = torch.where(first_greater, torch.arange(N, device=dev) * 2, torch.arange(N, device=dev) * 2 + 1)
    a, c = agent_states.actor_state, agent_states.critic_state
    host_step = None
    if agent_states.host_step is not None:
        K = lpg_hypers.num_agent_updates
        host_step = np.minimum(agent_states.host_step + K, np.maximum(agent_states.host_step, agent_states.level.lifetime)).astype(np.int32)
    agents_out = agent_states.replace(
        actor_state=a.replace(params=new_actor[pick], step=new_step[pick]),
        critic_state=c.replace(params=new_critic[pick], step=new_step[pick].clone()),
        env_obs=new_obs[pick], env_state=EnvState(new_state[pick], env.max_n_objs), host_step=host_step)
    # --- tell ---
    mean = st["mean"].clone(); m = st["m"].clone(); v = st["v"].clone()
    if world == 1:
        _lib.call("toued_es_tell", p(cand), p(rank), p(mean), p(m), p(v), popsize, P, Pp, float(st["sigma"]), float(st["lrate"]),
                  float(strat.beta_1), float(strat.beta_2), float(strat.eps), int(st["gen_counter"]), float(strat.mean_decay), s)
        fitness_all = fitness
    else:
        # this rank's part of noise^T fitness -> all-reduce -> identical Adam step on every rank (meta/train.py:203-216)
        gsum = torch.empty(P, dtype=torch.float32, device=dev)
        _lib.call("toued_es_grad_partial", p(cand), p(rank), p(mean), p(gsum), popsize, P, Pp, float(st["sigma"]), s)
        udist.all_reduce_sum(gsum)
        _lib.call("toued_es_adam", p(gsum), p(mean), p(m), p(v), pop_global, P, float(st["sigma"]), float(st["lrate"]),
                  float(strat.beta_1), float(strat.beta_2), float(strat.eps), int(st["gen_counter"]), float(strat.mean_decay), s)
        fitness_all = udist.all_gather_device(fitness)           # 2 N_global floats, for the metrics
        udist.all_reduce_sum(msum)
    new_es = {"mean": mean, "m": m, "v": v,
              "sigma": max(st["sigma"] * strat.sigma_decay, strat.sigma_limit),
              "lrate": max(st["lrate"] * strat.lrate_decay, strat.lrate_limit),
              "gen_counter": st["gen_counter"] + 1}
    metrics = {
        "fitness": {"mean": fitness_all.mean(), "min": fitness_all.min(), "max": fitness_all.max(),
                    "var": fitness_all.var(unbiased=False)},
        "lpg_agent": {k_: msum[i] / pop_global for i, k_ in enumerate(("policy_l2", "policy_entropy", "critic_loss",
                                                                    "critic_l2", "critic_entropy"))},
        "_fitness": fitness_all, "_candidates": cand[:, :P],
    }
    return lpg_train_state.replace(es_state=new_es), agents_out, None, metrics
