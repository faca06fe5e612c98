"""CUDA-graph replay of the LPG meta-gradient step.

The reference jit-compiles the whole experiment into one XLA program (train.py:66-67); the eager B200 path enqueues
~125 launches per meta-step from Python through 113 C-ABI calls, which costs 5.6 ms of host time per step (DESIGN.md
section 7) -- more than the GPU needs once the agents are sharded over 8 GPUs.  ``GraphedMetaGradStep`` captures one call
of ``lpg_meta_grad_train_step`` (all chunk / side streams fork from and join the capturing stream, the NCCL all-reduce
included) and replays it: per meta-step the host then does one 8-byte key upload and one graph launch.

What makes the step capturable: every input lives in a fixed device buffer (the step key, the LPG parameters and Adam
state incl. its update count, the agents' tables / steps / env states / level records, the value-critic tables); the
per-agent outputs are written back into those same buffers; nothing enqueued depends on host state.  Consequences for
the caller (the reference's pytrees are immutable; here the buffers are *donated*):
  * the returned train state / agent states alias the static buffers and are overwritten by the next call;
  * the returned metric scalars are copies (one 9-float clone per step), so a history of metrics stays valid.
The first call runs eagerly (it loads the kernels and allocates the workspaces), the second call captures."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ..util.data import AgentState, TrainState
from ..environments.gridworld.gridworld import EnvState
from .train import LPGTrainState, lpg_meta_grad_train_step, _advance_host_step

_KEY_RING = 16


class GraphedMetaGradStep:
    def __init__(self, **bound):
        self.bound = bound                         # rollout_manager, num_mini_batches, gamma, gae_lambda, lpg_hypers, ...
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.sig = None
        self.calls = 0

    # ------------------------------------------------------------------------------------------------------------
    def _signature(self, ts, ag, vc, kw):
        return (tuple(ag.actor_state.params.shape), tuple(ag.env_state.packed.shape), str(ag.actor_state.params.device),
                ts.params.numel(), bool(ts.model.lifetime_conditioning), tuple(sorted((k, repr(v)) for k, v in kw.items())))

    def _alloc(self, ts, ag, vc):
        dev = ag.actor_state.params.device
        c = lambda t: t.detach().clone()
        self.key = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        self.key_host = [torch.zeros((1, 2), dtype=torch.int32).pin_memory() for _ in range(_KEY_RING)]
        self.key_ev = [None] * _KEY_RING
        self.params, self.mu, self.nu = c(ts.params), c(ts.opt_state["mu"]), c(ts.opt_state["nu"])
        self.count = torch.tensor([int(ts.opt_state["count"])], dtype=torch.int32, device=dev)
        self.actor, self.critic = c(ag.actor_state.params), c(ag.critic_state.params)
        self.step, self.state, self.obs = c(ag.actor_state.step), c(ag.env_state.packed), c(ag.env_obs)
        self.levels = c(ag.level.packed)
        self.vparams, self.vstep = c(vc.params), c(vc.step)
        self.src = {}                              # name -> the tensor object last loaded into the static buffer

    def _load(self, name, static, t):
        """Copy ``t`` into the static buffer unless it IS the buffer or the very tensor loaded last time."""
        if t is static or self.src.get(name) is t:
            return
        static.copy_(t)
        self.src[name] = t

    def _load_inputs(self, rng, ts, ag, vc):
        slot = self.calls % _KEY_RING
        if self.key_ev[slot] is not None:
            self.key_ev[slot].synchronize()        # the upload that last used this pinned slot has completed
        self.key_host[slot].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(rng, np.uint32).reshape(1, 2)).view(np.int32)))
        from .. import _lib
        _lib.h2d(self.key_host[slot])
        self.key.copy_(self.key_host[slot], non_blocking=True)
        self.key_ev[slot] = torch.cuda.Event()
        self.key_ev[slot].record()
        self._load("params", self.params, ts.params)
        self._load("mu", self.mu, ts.opt_state["mu"])
        self._load("nu", self.nu, ts.opt_state["nu"])
        if ts.opt_state.get("count_dev") is not self.count:
            self.count.fill_(int(ts.opt_state["count"]))
        self._load("actor", self.actor, ag.actor_state.params)
        self._load("critic", self.critic, ag.critic_state.params)
        self._load("step", self.step, ag.actor_state.step)
        self._load("state", self.state, ag.env_state.packed)
        self._load("obs", self.obs, ag.env_obs)
        self._load("levels", self.levels, ag.level.packed)
        self._load("vparams", self.vparams, vc.params)
        self._load("vstep", self.vstep, vc.step)

    def _static_states(self, ts, ag, vc):
        sts = LPGTrainState(ts.model, self.params, ts.tx, {"mu": self.mu, "nu": self.nu, "count": ts.opt_state["count"],
                                                          "count_dev": self.count}, ts.step)
        a, cst = ag.actor_state, ag.critic_state
        level = type(ag.level)(ag.level.env_params, ag.level.lifetime, ag.level.buffer_id, self.levels)
        sag = AgentState(a.replace(params=self.actor, step=self.step), cst.replace(params=self.critic, step=self.step),
                         level, self.obs, EnvState(self.state, ag.env_state.max_n_objs), ag.host_step)
        svc = vc.replace(params=self.vparams, step=self.vstep)
        return sts, sag, svc

    def _capture(self, ts, ag, vc, kw):
        sts, sag, svc = self._static_states(ts, ag, vc)
        static = {"out": (self.actor, self.critic, self.step, self.state, self.obs), "count_dev": self.count}
        from .. import _lib
        self.graph = torch.cuda.CUDAGraph()
        before, steps_before = _lib.kernel_launches(), _lib.ENV_STEPS[0]
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            _, _, _, metrics = lpg_meta_grad_train_step(self.key, sts, sag, svc, **self.bound, **kw, _static=static)
        # kernels of OUR library inside one replay (the capture went through the same counted C-ABI calls); the capture
        # itself launched nothing
        self.kernels_per_replay = _lib.kernel_launches() - before
        self.env_steps_per_replay = _lib.ENV_STEPS[0] - steps_before
        _lib.GRAPH_KERNELS[0] -= self.kernels_per_replay
        _lib.ENV_STEPS[0] -= self.env_steps_per_replay
        # (the capture enqueued nothing: the two subtractions take its counted calls back out of the totals)
        self.mvec = metrics["_mvec"]
        self.grad = metrics.get("_grad")

    def release(self):
        """Drop the captured graph (and its NCCL nodes).  Call before torch.distributed.destroy_process_group():
        tearing the communicator down while a captured graph still references it blocks."""
        self.graph = None
        self.sig = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------------------------------------------------
    def __call__(self, rng, lpg_train_state, agent_states, value_critic_states, **kw):
        ts, ag, vc = lpg_train_state, agent_states, value_critic_states
        self.calls += 1
        if self.calls == 1:                        # eager warm-up call: loads kernels, allocates workspaces
            return lpg_meta_grad_train_step(rng, ts, ag, vc, **self.bound, **kw)
        sig = self._signature(ts, ag, vc, kw)
        if self.graph is None or sig != self.sig:
            self._alloc(ts, ag, vc)
            self._load_inputs(rng, ts, ag, vc)
            self._capture(ts, ag, vc, kw)
            self.sig = sig
        else:
            self._load_inputs(rng, ts, ag, vc)
        self.graph.replay()
        from .. import _lib
        _lib.GRAPH_KERNELS[0] += self.kernels_per_replay
        _lib.ENV_STEPS[0] += self.env_steps_per_replay
        for name in ("params", "mu", "nu", "actor", "critic", "step", "state", "obs", "vstep"):
            self.src.pop(name, None)               # overwritten in place by the step: whatever was loaded is stale
        K = self.bound["lpg_hypers"].num_agent_updates
        count = int(ts.opt_state["count"]) + 1
        new_ts = LPGTrainState(ts.model, self.params, ts.tx, {"mu": self.mu, "nu": self.nu, "count": count,
                                                             "count_dev": self.count}, ts.step + 1)
        a, cst = ag.actor_state, ag.critic_state
        level = ag.level
        if level.packed is not self.levels:
            level = type(level)(level.env_params, level.lifetime, level.buffer_id, self.levels)
            self.src["levels"] = self.levels
        new_ag = AgentState(a.replace(params=self.actor, step=self.step), cst.replace(params=self.critic, step=self.step),
                            level, self.obs, EnvState(self.state, ag.env_state.max_n_objs),
                            _advance_host_step(ag.host_step, ag.level.lifetime, K))
        new_vc = vc.replace(params=self.vparams, step=self.vstep)
        mv = self.mvec.clone()                     # the only per-step copy: 9 floats
        metrics = {"lpg_loss": mv[0], "reg_lpg_loss": mv[1], "value_loss": mv[2],
                   "lpg_agent": {"policy_l2": mv[3], "policy_entropy": mv[4], "critic_loss": mv[5],
                                 "critic_l2": mv[6], "critic_entropy": mv[7]},
                   "lpg_agent_return": mv[8]}
        if self.grad is not None:
            metrics["_grad"] = self.grad.clone()
        return new_ts, new_ag, new_vc, metrics
