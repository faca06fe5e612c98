"""Gridworld oracle (numpy, vectorised).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Restates environments/gridworld/gridworld.py:12-211 for ``tabular=True`` environments (every
in-scope env_mode) plus gymnax==0.0.6 ``Environment.step`` / ``reset``
([3P-recall]: ``key, key_reset = split(key)``; ``step_env``; ``reset_env(key_reset)``; select the
reset state/obs where ``done``).

Vectorisation: EnvParams carry a leading agent axis ``[N]``; EnvState carries ``[N, W]``
(agents x workers).  Observations are never materialised densely: an observation is the pair
``(idx, time)`` with ``obs[idx] = 1`` and ``obs[D-1] = time * 0.001`` (gridworld.py:184-205);
``dense_obs`` builds the reference's f32[D] vector for small tests.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
import numpy as np

from . import prng

F32 = np.float32


@dataclass
class EnvParams:
    """gridworld.py:22-35.  Every field has a leading agent axis [N]."""
    max_steps_in_episode: np.ndarray  # i32[N]
    random_respawn: np.ndarray        # bool[N]   (unused when tabular)
    auto_collect: np.ndarray          # bool[N]   (never read by the reference step)
    grid_size: np.ndarray             # i32[N]
    walls: np.ndarray                 # bool[N, G*G]
    start_pos: np.ndarray             # i32[N]
    n_objs: np.ndarray                # i32[N]
    obj_ids: np.ndarray               # i32[N, O]   (-1 = padding)
    static_obj_poss: np.ndarray       # i32[N, O]
    obj_rewards: np.ndarray           # f32[N, T]
    obj_p_terminate: np.ndarray       # f32[N, T]
    obj_p_respawn: np.ndarray         # f32[N, T]

    def __len__(self):
        return int(self.grid_size.shape[0])

    def index(self, ids):
        return EnvParams(**{k: np.asarray(v)[ids] for k, v in self.__dict__.items()})


@dataclass
class EnvState:
    """gridworld.py:12-18, batched [N, W] (obj arrays [N, W, O])."""
    time: np.ndarray
    pos: np.ndarray
    obj_poss: np.ndarray
    obj_existss: np.ndarray
    early_term: np.ndarray


class GridWorld:
    """gridworld.py:38-51 (tabular only)."""

    def __init__(self, max_grid_size=11, max_n_objs=4, max_n_obj_types=3, tabular=True):
        if not tabular:
            raise NotImplementedError("oracle restates tabular gridworlds only (SURVEY.md Q13)")
        self.max_grid_size = max_grid_size
        self.max_n_objs = max_n_objs
        self.max_n_obj_types = max_n_obj_types
        self.tabular = True

    # gridworld.py:219-222
    num_actions = 5

    @property
    def obs_dim(self):
        """gridworld.py:230-233"""
        return self.max_grid_size ** 2 * (2 ** self.max_n_objs) + 1

    @property
    def default_params(self) -> EnvParams:
        """gridworld.py:54-70 (N = 1)."""
        return EnvParams(
            max_steps_in_episode=np.array([500], np.int32),
            random_respawn=np.array([False]),
            auto_collect=np.array([True]),
            grid_size=np.array([11], np.int32),
            walls=np.zeros((1, 121), bool),
            start_pos=np.array([0], np.int32),
            n_objs=np.array([4], np.int32),
            obj_ids=np.array([[0, 0, 1, 2]], np.int32),
            static_obj_poss=np.array([[1 * 11 + 3, 3 * 11 + 7, 8 * 11 + 7, 9 * 11 + 2]], np.int32),
            obj_rewards=np.array([[1.0, -1.0, -1.0]], F32),
            obj_p_terminate=np.array([[0.0, 0.5, 0.0]], F32),
            obj_p_respawn=np.array([[0.05, 0.1, 0.5]], F32),
        )

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _take_type(table, obj_ids):
        """jnp.take(params.obj_X, params.obj_ids): negative ids wrap python-style
        ([3P-recall] jnp.take normalises negative indices before applying its mode)."""
        T = table.shape[-1]
        ids = np.where(obj_ids < 0, obj_ids + T, obj_ids)
        ids = np.clip(ids, 0, T - 1)
        return np.take_along_axis(table, ids, axis=-1)

    def _get_next_pos(self, pos, action, p: EnvParams):
        """gridworld.py:138-146.  pos, action: [N, W]"""
        g = p.grid_size[:, None]
        top, bottom = pos < g, pos >= g * (g - 1)
        left, right = (pos % g) == 0, (pos % g) == g - 1
        step = ((action == 0) * (1 - top) * -g + (action == 1) * (1 - bottom) * g
                + (action == 2) * (1 - left) * -1 + (action == 3) * (1 - right) * 1)
        nxt = pos + step
        blocked = np.take_along_axis(p.walls, nxt, axis=1)
        return np.where(blocked, pos, nxt).astype(np.int32)

    def _get_tabular_pos(self, pos, exists):
        """gridworld.py:201-205"""
        w = (1 << np.arange(self.max_n_objs)).astype(np.int32)
        return (pos + self.max_grid_size ** 2 * (exists * w).sum(-1)).astype(np.int32)

    def obs_index(self, s: EnvState):
        """(idx, time) form of get_obs, gridworld.py:184-199"""
        return self._get_tabular_pos(s.pos, s.obj_existss), s.time.astype(np.int32)

    def dense_obs(self, s: EnvState):
        idx, t = self.obs_index(s)
        obs = np.zeros(idx.shape + (self.obs_dim,), F32)
        np.put_along_axis(obs, idx[..., None], F32(1.0), axis=-1)
        obs[..., -1] = t.astype(F32) * F32(0.001)
        return obs

    # ------------------------------------------------------------------ reset / step
    def reset_env(self, key, p: EnvParams, W: int) -> EnvState:
        """gridworld.py:157-182.  key is unused in tabular mode (obj_key/pos_key are dead)."""
        N, O, G2 = len(p), self.max_n_objs, self.max_grid_size ** 2
        bc = lambda a: np.broadcast_to(a[:, None], (N, W)).copy()
        return EnvState(
            time=np.zeros((N, W), np.int32),
            pos=bc(p.start_pos.astype(np.int32)),
            obj_poss=np.broadcast_to((p.static_obj_poss + p.obj_ids * G2)[:, None, :], (N, W, O)).astype(np.int32).copy(),
            obj_existss=np.broadcast_to((np.arange(O)[None, :] < p.n_objs[:, None])[:, None, :], (N, W, O)).copy(),
            early_term=np.zeros((N, W), bool),
        )

    def step_env(self, key, s: EnvState, action, p: EnvParams):
        """gridworld.py:72-136.  key: uint32[N, W, 2]"""
        G2, O = self.max_grid_size ** 2, self.max_n_objs
        ks = prng.split(key, 3)
        term_key, respawn_key = ks[..., 0, :], ks[..., 1, :]      # obj_key unused (tabular)
        pos = self._get_next_pos(s.pos, action, p)
        old_obj_poss = s.obj_poss - (p.obj_ids * G2)[:, None, :]
        collected = s.obj_existss & (old_obj_poss == pos[..., None])
        p_resp = self._take_type(p.obj_p_respawn, p.obj_ids)[:, None, :]
        respawn = prng.uniform(respawn_key, (O,)) < p_resp
        exists = s.obj_existss | respawn
        obj_poss = old_obj_poss + (p.obj_ids * G2)[:, None, :]
        exists = exists & ~collected
        exists = exists & (np.arange(O)[None, None, :] < p.n_objs[:, None, None])
        p_term = self._take_type(p.obj_p_terminate, p.obj_ids)[:, None, :]
        # jnp.dot(padded_p_terminate, obj_collected): sequential f32 sum over objects
        pt = np.zeros(pos.shape, F32)
        rew = np.zeros(pos.shape, F32)
        r_obj = self._take_type(p.obj_rewards, p.obj_ids)[:, None, :]
        for i in range(O):
            c = collected[..., i].astype(F32)
            pt = (pt + (p_term[..., i] * c).astype(F32)).astype(F32)
            rew = (rew + (r_obj[..., i] * c).astype(F32)).astype(F32)
        term = (prng.uniform(term_key, ()) < pt) | s.early_term
        time = s.time + 1
        ns = EnvState(time.astype(np.int32), pos, obj_poss.astype(np.int32), exists, term)
        done = self.is_terminal(ns, p)
        return ns, rew, done

    def is_terminal(self, s: EnvState, p: EnvParams):
        """gridworld.py:207-211"""
        return (s.time >= p.max_steps_in_episode[:, None]) | s.early_term

    def step(self, key, s: EnvState, action, p: EnvParams):
        """gymnax Environment.step with auto-reset [3P-recall]."""
        ks = prng.split(key, 2)
        k_step, k_reset = ks[..., 0, :], ks[..., 1, :]
        st, rew, done = self.step_env(k_step, s, action, p)
        re = self.reset_env(k_reset, p, s.pos.shape[1])
        sel = lambda a, b: np.where(done if a.ndim == 2 else done[..., None], a, b)
        ns = EnvState(sel(re.time, st.time), sel(re.pos, st.pos), sel(re.obj_poss, st.obj_poss),
                      sel(re.obj_existss, st.obj_existss), sel(re.early_term, st.early_term))
        return ns, rew, done

    def reset(self, key, p: EnvParams, W: int):
        return self.reset_env(key, p, W)


# ---------------------------------------------------------------------- optimal_return (DP)
def optimal_return(env: GridWorld, p: EnvParams, max_rollout_len: int, agent: int = 0) -> float:
    """gridworld.py:253-323: exact finite-horizon DP over the tabular MDP of ONE level
    (float64; used as a known-answer upper bound on any policy's expected first-episode return).
    """
    G2, O = env.max_grid_size ** 2, env.max_n_objs
    g = int(p.grid_size[agent]); n_objs = int(p.n_objs[agent]); cap = int(p.max_steps_in_episode[agent])
    walls = p.walls[agent]
    one = p.index([agent])
    r_obj = env._take_type(one.obj_rewards, one.obj_ids)[0].astype(np.float64)
    p_t = env._take_type(one.obj_p_terminate, one.obj_ids)[0].astype(np.float64)
    p_r = env._take_type(one.obj_p_respawn, one.obj_ids)[0].astype(np.float64)
    spos = p.static_obj_poss[agent]
    nstate = 1 << O
    valid_mask = [m for m in range(nstate) if (m >> n_objs) == 0]
    v_next = np.zeros((G2, nstate))
    cells = [c for c in range(g * g) if not walls[c]]
    for time in reversed(range(max_rollout_len)):
        v = np.full((G2, nstate), -np.inf)
        if time >= cap:
            v_next = np.zeros((G2, nstate)); continue
        for c in cells:
            for m in valid_mask:
                best = -np.inf
                for a in range(5):
                    nxt = int(env._get_next_pos(np.array([[c]], np.int32), np.array([[a]]), one)[0, 0])
                    coll = [(m >> i) & 1 and spos[i] == nxt for i in range(O)]
                    r = sum(r_obj[i] for i in range(O) if coll[i])
                    pterm = sum(p_t[i] for i in range(O) if coll[i])
                    ev = 0.0
                    for m2 in valid_mask:
                        pr = 1.0
                        for i in range(n_objs):
                            b = (m2 >> i) & 1
                            if coll[i]:
                                pr *= (1 - b)
                            elif (m >> i) & 1:
                                pr *= b
                            else:
                                pr *= p_r[i] if b else (1 - p_r[i])
                        if pr > 0:
                            ev += pr * v_next[nxt, m2]
                    best = max(best, r + ev * (1 - pterm))
                v[c, m] = best
        v_next = np.where(np.isfinite(v), v, 0.0)
        v_fin = v
    start = int(p.start_pos[agent]); m0 = (1 << n_objs) - 1
    return float(v_fin[start, m0])
