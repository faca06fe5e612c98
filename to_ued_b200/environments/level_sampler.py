"""Level sampler: domain randomisation and prioritised level replay (reference
environments/level_sampler.py:25-426), batched over agents.

Division of labour.  The level *buffer* (<= buffer_size small records: score, active, new flags,
level parameters) and its index logic live on the host in numpy: it is a few thousand scalars per
meta-step and its choices must be bit-exact, which integer numpy code gives for free.  Everything
that touches per-agent state — table initialisation, environment resets, A2C antagonist training and
the evaluation rollouts that produce the regret scores — runs in the CUDA kernels.

The reference re-creates *every* agent on every call of ``sample`` and then masks with
``terminated`` (level_sampler.py:236-265); here only the terminated agents are re-created, which is
the same result.  ``actor_state.step`` is mirrored on the host (it evolves deterministically:
``step <- step + 1`` while ``step < lifetime``), so ``sample`` never synchronises with the device.

Multi-GPU: rank r holds agents [r n, (r + 1) n) of the global batch.  ``sample`` and
``initial_sample`` derive every key for the GLOBAL batch (``split(rng, n_global)``) and keep the local slice, the
level buffer is replicated and updated identically on every rank from all-gathered ``(buffer_id, score,
terminated)`` triples, so the levels, ids and init keys of agent i do not depend on the number of ranks.  The
host-side decisions are separated from the device work (``_plan_sample`` / ``_recreate``) so that the
rank-independence is testable without a GPU (tests/test_05_multi_rank_cpu.py).

Reproduced quirks: Q3 (``new=level_buffer.active.at[reset_ids].set(True)``), Q10 (buffer ids of random
levels are 0)."""
from __future__ import annotations

from dataclasses import dataclass, replace as _replace
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..util import prng
from ..util import dist as udist
from ..util.data import AgentState, Level, TrainState
from ..agents.agents import AgentHyperparams, eval_agent
from ..models.agent import init_tables
from .environments import get_env, reset_env_params, get_env_spec
from .gridworld.gridworld import EnvParams, EnvState, pack_levels, levels_to_device
from .rollout import RolloutWrapper

# TODO(reference): positive_value_loss and l1_value_loss are not implemented upstream either
SCORE_FUNCTIONS = ["random", "frozen", "alg_regret"]
SCORE_TRANSFORMS = ["proportional", "rank"]


@dataclass
class LevelBuffer:
    """level_sampler.py:29-54 (host side)."""
    level: Level
    score: np.ndarray       # f32[B]
    active: np.ndarray      # bool[B]
    new: np.ndarray         # bool[B]

    @staticmethod
    def create_buffer(params: EnvParams, lifetimes):
        n = len(lifetimes)
        return LevelBuffer(Level(params, np.asarray(lifetimes, np.int32), np.arange(n, dtype=np.int32)),
                           np.zeros(n, np.float32), np.zeros(n, bool), np.ones(n, bool))

    def replace(self, **kw):
        return _replace(self, **kw)

    def __len__(self):
        return self.score.shape[0]


def _index_level(level: Level, ids) -> Level:
    ids = np.asarray(ids)
    return Level(level.env_params[ids], level.lifetime[ids], level.buffer_id[ids])


def _where_level(mask, a: Level, b: Level) -> Level:
    return Level(a.env_params.where(mask, b.env_params), np.where(mask, a.lifetime, b.lifetime).astype(np.int32),
                 np.where(mask, a.buffer_id, b.buffer_id).astype(np.int32))


class LevelSampler:
    """Level sampler, containing methods for domain randomisation and prioritised level replay."""

    def __init__(self, args, device="cuda"):
        self.device = device
        self.rank, self.world = udist.rank_world()
        self.env_name, self.env_mode, self.env_workers = args.env_name, args.env_mode, args.env_workers
        self.env_kwargs, self.max_rollout_len, self.max_lifetime = get_env_spec(self.env_name, self.env_mode)
        self.env = get_env(self.env_name, self.env_kwargs)
        self.rollout_manager = RolloutWrapper(self.env_name, args.train_rollout_len, self.max_rollout_len, self.env_kwargs)
        self.agent_hypers = AgentHyperparams.from_args(args)
        if args.score_function not in SCORE_FUNCTIONS:
            raise ValueError(f"Level score function {args.score_function} not in known functions: {SCORE_FUNCTIONS}")
        if args.score_transform not in SCORE_TRANSFORMS:
            raise ValueError(f"Level score transform {args.score_transform} not in known transforms: {SCORE_TRANSFORMS}")
        self.score_function, self.score_transform = args.score_function, args.score_transform
        self.score_temperature, self.buffer_size = args.score_temperature, args.buffer_size
        self.p_replay, self.num_mini_batches = args.p_replay, args.num_mini_batches
        from ..agents.a2c import A2CHyperparams
        self.a2c_hypers = A2CHyperparams(args.gamma, args.gae_lambda, args.entropy_coeff)

    # ------------------------------------------------------------------------------ buffer
    def initialize_buffer(self, rng):
        """level_sampler.py:90-96"""
        if self.score_function == "random":
            return None
        params, lifetimes = self._sample_env_params(prng.split(np.asarray(rng, np.uint32), self.buffer_size))
        return LevelBuffer.create_buffer(params, lifetimes)

    def _sample_env_params(self, rng):
        """level_sampler.py:98-101 (vmapped over keys)."""
        return reset_env_params(rng, self.env_name, self.env_mode)

    def _sample_random_levels(self, rng, batch_size: int, sl: slice = slice(None)) -> Level:
        """level_sampler.py:268-271 (Q10: buffer ids are zeros).  ``sl``: the part of the batch to generate (one key
        per level, so a slice of the keys gives exactly that slice of the levels)."""
        keys = prng.split(np.asarray(rng, np.uint32), batch_size)[sl]
        params, lifetimes = self._sample_env_params(keys)
        return Level(params, lifetimes, np.zeros(len(keys), np.int32))

    # ------------------------------------------------------------------------------ agents
    def _device_level(self, level: Level) -> Level:
        level.packed = levels_to_device(pack_levels(level.env_params, level.lifetime, level.buffer_id), self.device)
        return level

    def _create_agent(self, rng, level: Level, value_critic=False) -> AgentState:
        """level_sampler.py:273-291, batched over keys [N, 2] / levels [N]."""
        rng = np.asarray(rng, np.uint32).reshape(-1, 2)
        ks = prng.split(rng, 2)
        worker_rng, agent_rng = ks[:, 0, :], ks[:, 1, :]
        if level.packed is None:
            self._device_level(level)
        env_obs, env_state = self.rollout_manager.batch_reset(worker_rng, level.packed, self.env_workers)
        hy = self.agent_hypers.replace(critic_dims=1) if value_critic else self.agent_hypers
        from ..agents.agents import create_agent
        actor, critic = create_agent(agent_rng, hy, self.num_actions, self.obs_shape, self.device)
        return AgentState(actor, critic, level, env_obs, env_state, np.zeros(len(level), np.int32))

    def initial_sample(self, rng, level_buffer: Optional[LevelBuffer], batch_size: int, create_value_critics: bool):
        """level_sampler.py:103-132.  ``batch_size`` is the GLOBAL number of agents; with several ranks the agents
        [rank * n, (rank + 1) * n) are created here (same levels and keys as in a single-rank run)."""
        rng = np.asarray(rng, np.uint32)
        sl = udist.local_slice(batch_size)
        if self.score_function == "random":
            rng, _rng = prng.split(rng, 2)
            levels = self._sample_random_levels(_rng, batch_size, sl)
        else:
            levels = _index_level(level_buffer.level, np.arange(batch_size)[sl])
            level_buffer = level_buffer.replace(active=np.arange(self.buffer_size) < batch_size)
        rng, _rng = prng.split(rng, 2)
        agent_states = self._create_agent(prng.split(_rng, batch_size)[sl], levels)
        value_critics = None
        if create_value_critics:
            from ..agents.agents import create_value_critic
            rng, _rng = prng.split(rng, 2)
            value_critics = create_value_critic(prng.split(_rng, batch_size)[sl], self.agent_hypers, self.obs_shape, self.device)
        return level_buffer, agent_states, value_critics

    # ------------------------------------------------------------------------------ sample
    def sample(self, rng, level_buffer: Optional[LevelBuffer], old_agents: AgentState, old_value_critics):
        """Update level buffer and sample new levels for terminated agents (level_sampler.py:134-266)."""
        score_fn = lambda keys, only: self._compute_algorithmic_regret(keys, old_agents, only=only)
        level_buffer, plan = self._plan_sample(rng, level_buffer, old_agents.host_step, old_agents.level,
                                               old_value_critics is not None, score_fn)
        if plan is None:
            return level_buffer, old_agents, old_value_critics
        terminated, new_levels, agent_keys, value_keys = plan
        return (level_buffer,) + self._recreate(terminated, new_levels, agent_keys, value_keys, old_agents, old_value_critics)

    def _plan_sample(self, rng, level_buffer, host_step, old_level: Level, have_value_critics: bool, score_fn):
        """Host-side decisions of ``sample`` for this rank's agents: (new level buffer, plan) with plan = None when no
        agent of the GLOBAL batch terminated, else (terminated[n_local], new levels[n_local], agent init keys
        [n_local, 2], value-critic init keys or None).  Every draw is made for the global batch and sliced, and the
        buffer is updated from all-gathered triples: the result does not depend on the number of ranks.
        ``score_fn(keys[n_local, 2], only=mask) -> f32[n_local]`` is the regret evaluation (device work)."""
        rng = np.asarray(rng, np.uint32)
        terminated = np.asarray(host_step) >= old_level.lifetime
        n_local = terminated.shape[0]
        batch_size = n_local * self.world                      # global
        sl = slice(self.rank * n_local, (self.rank + 1) * n_local)
        terminated_g = udist.all_gather_host(terminated)

        if self.score_function == "random":
            rng, _rng = prng.split(rng, 2)
            if terminated.any():
                new_levels = _where_level(terminated, self._sample_random_levels(_rng, batch_size, sl), old_level)
            else:
                new_levels = old_level
        elif self.score_function == "frozen":
            rng, _rng = prng.split(rng, 2)
            p_uniform = np.ones(self.buffer_size, np.float32) / np.float32(self.buffer_size)
            level_ids = _choice_p_many(_rng, p_uniform, batch_size)[sl]
            new_levels = _where_level(terminated, _index_level(level_buffer.level, level_ids), old_level)
        else:
            rng, _rng = prng.split(rng, 2)
            level_buffer = self._reset_lowest_scoring(_rng, level_buffer, batch_size)
            if self.score_function != "alg_regret":
                raise NotImplementedError(f"Level score function {self.score_function} is not implemented.")
            rng, _rng = prng.split(rng, 2)
            score = np.asarray(score_fn(prng.split(_rng, batch_size)[sl], terminated), np.float32)
            # every rank applies the same update: all-gather of (buffer_id, score, terminated) per agent
            old_ids = udist.all_gather_host(old_level.buffer_id.astype(np.int32))
            score = udist.all_gather_host(score)
            # sequential scatter == .at[old_ids].set(...) with duplicate ids resolved last-wins
            sc, ac, nw = level_buffer.score.copy(), level_buffer.active.copy(), level_buffer.new.copy()
            t_score = np.where(terminated_g, score, level_buffer.score[old_ids])
            t_active = np.where(terminated_g, False, level_buffer.active[old_ids])
            t_new = np.where(terminated_g, False, level_buffer.new[old_ids])
            sc[old_ids], ac[old_ids], nw[old_ids] = t_score, t_active, t_new
            level_buffer = level_buffer.replace(score=sc, active=ac, new=nw)
            rng, replay_rng, random_rng = prng.split(rng, 3)
            replay_ids = self._replay_ids(replay_rng, level_buffer, batch_size)
            random_ids = self._random_ids(random_rng, level_buffer, batch_size)
            rng, _rng = prng.split(rng, 2)
            n_to_replay = int((prng.uniform(_rng, (batch_size,)) < np.float32(self.p_replay)).sum())
            use_replay = np.arange(batch_size) < n_to_replay
            n_replayable = self.buffer_size - int((level_buffer.new | level_buffer.active).sum())
            use_replay = use_replay & (n_replayable >= batch_size)
            rng, _rng = prng.split(rng, 2)
            use_replay = use_replay[prng.shuffle_prefix(_rng, batch_size, batch_size)]      # random.permutation
            new_ids = np.where(terminated_g, np.where(use_replay, replay_ids, random_ids), old_ids).astype(np.int32)
            new_levels = _where_level(terminated, _index_level(level_buffer.level, new_ids[sl]), old_level)
            ac = level_buffer.active.copy()
            ac[new_ids] = True
            level_buffer = level_buffer.replace(active=ac)

        # --- Initialise new agents and environment workers for terminated agents ---
        rng, _rng = prng.split(rng, 2)
        agent_keys = prng.split(_rng, batch_size)[sl]
        value_keys = None
        if have_value_critics:
            rng, _rng = prng.split(rng, 2)
            value_keys = prng.split(_rng, batch_size)[sl]
        if not terminated_g.any():
            return level_buffer, None
        return level_buffer, (terminated, new_levels, agent_keys, value_keys)

    def _recreate(self, terminated, new_levels: Level, agent_keys, value_keys, old_agents: AgentState, old_vc):
        """Masked version of level_sampler.py:236-265: only terminated agents get new tables / envs."""
        dev = self.device
        D, W = self.obs_shape[0], self.env_workers
        mask_d = _lib.h2d(torch.from_numpy(terminated.astype(np.uint8))).to(dev, non_blocking=True)
        new_levels = self._device_level(new_levels)
        ks = prng.split(agent_keys, 2)                       # worker_rng, agent_rng (level_sampler.py:275)
        ks2 = prng.split(ks[:, 1, :], 2)                     # actor_rng, critic_rng (agents.py:37)
        actor, critic = old_agents.actor_state, old_agents.critic_state
        a_params = init_tables(ks2[:, 0, :], D, actor.n_out, dev, out=actor.params.clone(), mask=mask_d)
        c_params = init_tables(ks2[:, 1, :], D, critic.n_out, dev, out=critic.params.clone(), mask=mask_d)
        state = old_agents.env_state.packed.clone()
        obs = old_agents.env_obs.clone()
        step = actor.step.clone()
        _lib.call("toued_masked_reset", _lib.ptr(new_levels.packed), _lib.ptr(mask_d), _lib.ptr(state), _lib.ptr(obs),
                  _lib.ptr(step), len(terminated), W, self.env.max_grid_size, _lib.stream_ptr())
        agents = AgentState(actor.replace(params=a_params, step=step), critic.replace(params=c_params, step=step.clone()),
                            new_levels, obs, EnvState(state, self.env.max_n_objs),
                            np.where(terminated, 0, old_agents.host_step).astype(np.int32))
        vc = None
        if old_vc is not None:
            v_params = init_tables(value_keys, D, 1, dev, out=old_vc.params.clone(), mask=mask_d)
            vc = old_vc.replace(params=v_params, step=torch.where(mask_d.bool(), torch.zeros_like(old_vc.step), old_vc.step))
        return agents, vc

    # ------------------------------------------------------------------------------ PLR helpers
    def _compute_algorithmic_regret(self, rng, lpg_agent_state: AgentState, only=None):
        """level_sampler.py:293-329: A2C antagonist return minus LPG agent return, f32[N] (host)."""
        from ..agents.a2c import train_a2c_agent
        rng = np.asarray(rng, np.uint32).reshape(-1, 2)
        n = rng.shape[0]
        score = np.zeros(n, np.float32)
        sel = np.arange(n) if only is None else np.nonzero(only)[0]
        if len(sel) == 0:
            return score                      # the reference computes (and discards) all scores; only
        rng = rng[sel]                        # terminated agents' scores are ever used (:188)
        level = _index_level(lpg_agent_state.level, sel)
        ks = prng.split(rng, 2); rng, _rng = ks[:, 0, :], ks[:, 1, :]
        a2c_agent = self._create_agent(_rng, level, value_critic=True)
        ks = prng.split(rng, 2); rng, _rng = ks[:, 0, :], ks[:, 1, :]
        a2c_agent, _ = train_a2c_agent(_rng, a2c_agent, self.rollout_manager, self.max_lifetime, self.a2c_hypers)
        ks = prng.split(rng, 2)
        lpg_rng, a2c_rng = ks[:, 0, :], ks[:, 1, :]
        sel_d = _lib.h2d(torch.from_numpy(sel)).to(self.device)
        lpg_ret = eval_agent(lpg_rng, self.rollout_manager, a2c_agent.level.packed,
                             lpg_agent_state.actor_state.params[sel_d], self.env_workers)
        a2c_ret = eval_agent(a2c_rng, self.rollout_manager, a2c_agent.level.packed, a2c_agent.actor_state,
                             self.env_workers)
        score[sel] = (a2c_ret - lpg_ret).cpu().numpy()
        return score

    def _reset_lowest_scoring(self, rng, level_buffer: LevelBuffer, minimum_new: int) -> LevelBuffer:
        """level_sampler.py:331-353 (Q3 reproduced: ``new`` is rebuilt from ``active``)."""
        level_scores = np.where(level_buffer.new, -np.inf, level_buffer.score).astype(np.float32)
        level_scores = np.where(level_buffer.active, np.inf, level_scores)
        reset_ids = np.argsort(level_scores, kind="stable")[:minimum_new]
        new_params, new_lifetimes = self._sample_env_params(prng.split(np.asarray(rng, np.uint32), minimum_new))
        lv = level_buffer.level
        params = EnvParams(**{f: _set_rows(getattr(lv.env_params, f), reset_ids, getattr(new_params, f))
                              for f in lv.env_params.__dataclass_fields__})
        level = Level(params, _set_rows(lv.lifetime, reset_ids, new_lifetimes), _set_rows(lv.buffer_id, reset_ids, reset_ids))
        return level_buffer.replace(level=level, score=_set_rows(level_buffer.score, reset_ids, 0.0),
                                    active=_set_rows(level_buffer.active, reset_ids, False),
                                    new=_set_rows(level_buffer.active, reset_ids, True))

    def _replay_ids(self, rng, level_buffer: LevelBuffer, batch_size: int) -> np.ndarray:
        """level_sampler.py:355-387 (ids of the replayed levels)."""
        invalid = level_buffer.new | level_buffer.active
        scores = np.exp(level_buffer.score / np.float32(self.score_temperature)).astype(np.float32)
        scores = np.where(invalid, np.float32(0.0), scores)
        scores = (scores / scores.sum(dtype=np.float32)).astype(np.float32)
        p_replay = np.where(self.buffer_size - invalid.sum() < batch_size, np.ones_like(scores), scores)
        if self.score_transform == "rank":
            return np.argsort(p_replay, kind="stable")[::-1][:batch_size].astype(np.int32)
        if self.score_transform == "proportional":
            _rng = prng.split(np.asarray(rng, np.uint32), 2)[1]
            return _gumbel_topk(_rng, p_replay, batch_size)
        raise NotImplementedError(f"Level score transform {self.score_transform} is not implemented.")

    def _replay_from_buffer(self, rng, level_buffer: LevelBuffer, batch_size: int) -> Level:
        return _index_level(level_buffer.level, self._replay_ids(rng, level_buffer, batch_size))

    def _random_ids(self, rng, level_buffer: LevelBuffer, batch_size: int) -> np.ndarray:
        """level_sampler.py:389-408: new (unevaluated), inactive levels, without replacement."""
        mask = level_buffer.new & ~level_buffer.active
        return prng.masked_topk(np.asarray(rng, np.uint32), mask, batch_size)

    def _sample_random_from_buffer(self, rng, level_buffer: LevelBuffer, batch_size: int) -> Level:
        return _index_level(level_buffer.level, self._random_ids(rng, level_buffer, batch_size))

    @property
    def num_actions(self):
        return self.env.num_actions

    @property
    def obs_shape(self):
        return self.env.observation_space(self.env.default_params).shape


def _set_rows(arr, ids, val):
    out = np.array(arr, copy=True)
    out[ids] = val
    return out


def _choice_p_many(key, p, n):
    """jax.random.choice(key, arange(len(p)), (n,), replace=True, p=p): inverse-CDF on a left-to-right cumsum."""
    pc = np.cumsum(p.astype(np.float32), dtype=np.float32)
    u = prng.uniform(key, (n,))
    r = (pc[-1] * (np.float32(1.0) - u)).astype(np.float32)
    return (pc[None, :] < r[:, None]).sum(-1).astype(np.int32)


def _gumbel_topk(key, p, k):
    """choice(..., replace=False, p=p) for general p: argsort(-gumbel - log p)[:k]."""
    u = prng.uniform(key, (len(p),), np.finfo(np.float32).tiny, 1.0)
    g = -(-np.log(-np.log(u))) - np.log(p.astype(np.float32))
    return np.argsort(g, kind="stable")[:k].astype(np.int32)
