"""Shared test plumbing: build matching oracle / product inputs from one seed."""
import numpy as np
import torch

from oracle import prng, configs
from oracle.gridworld import GridWorld as OGrid
from oracle.rollout import RolloutWrapper as ORollout, Trajectory
from oracle.lpg import LPGLayout, init_lpg_params


class Case:
    """N agents on levels of ``mode``; identical inputs for the oracle and the CUDA path."""

    def __init__(self, mode="all_shortlife", n=6, w=64, L=20, seed=0, table_scale=None, cond=False,
                 lifetimes=None, steps=None):
        self.mode, self.n, self.w, self.L = mode, n, w, L
        self.keys = prng.split(prng.PRNGKey(seed), n)
        self.p, self.life = configs.reset_env_params(self.keys, mode)
        if lifetimes is not None:
            self.life = np.asarray(lifetimes, np.int32)
        self.kw, self.ep = configs.get_env_spec(mode)
        self.oenv = OGrid(**self.kw)
        self.D = self.oenv.obs_dim
        self.oro = ORollout(self.oenv, L, self.ep)
        rs = np.random.RandomState(seed + 1)
        sc = table_scale if table_scale is not None else 1.0 / np.sqrt(self.D)
        self.actor = (rs.randn(n, self.D, 5) * sc).astype(np.float32)
        self.critic = (rs.randn(n, self.D, 8) * sc).astype(np.float32)
        self.value = (rs.randn(n, self.D, 1) * sc).astype(np.float32)
        self.layout = LPGLayout(lifetime_conditioning=cond)
        self.lpg = init_lpg_params(self.layout, seed)
        # non-trivial biases so every parameter block is exercised
        self.lpg += (rs.randn(self.lpg.size) * 0.02).astype(np.float32)
        self.steps = np.zeros(n, np.int32) if steps is None else np.asarray(steps, np.int32)

    # ---- product-side objects (GPU) ----
    def device_levels(self):
        from to_ued_b200.environments.gridworld.gridworld import EnvParams, pack_levels, levels_to_device
        pp = EnvParams(**{k: getattr(self.p, k) for k in self.p.__dataclass_fields__})
        self.pparams = pp
        return levels_to_device(pack_levels(pp, self.life))

    @staticmethod
    def pad8(t):
        out = np.zeros(t.shape[:-1] + (8,), np.float32)
        out[..., : t.shape[-1]] = t
        return torch.from_numpy(out).cuda()

    def agent_state(self):
        from to_ued_b200.util.data import AgentState, TrainState, Level
        from to_ued_b200.environments.rollout import RolloutWrapper
        lv = self.device_levels()
        ro = RolloutWrapper("GridWorld-v0", self.L, self.ep, self.kw)
        obs0, st0 = ro.batch_reset(None, lv, self.w)
        step = torch.from_numpy(self.steps.copy()).cuda()
        a = TrainState(self.pad8(self.actor), step, 5, 4e1, 0.5)
        c = TrainState(self.pad8(self.critic), step.clone(), 8, 4e0, 0.5)
        level = Level(self.pparams, self.life, np.zeros(self.n, np.int32), lv)
        return AgentState(a, c, level, obs0, st0), ro


def to_oracle_traj(tr) -> Trajectory:
    o = tr.obs.cpu().numpy()
    return Trajectory((o & 0xFFFF).astype(np.int32), ((o >> 16) & 0xFFFF).astype(np.int32),
                      tr.action.cpu().numpy().astype(np.int32), tr.reward.cpu().numpy(),
                      tr.done.cpu().numpy().astype(bool))


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))
