"""Hand-worked transitions on the reference's default level (gridworld.py:54-70) and structural
properties of the oracle env / rollout / level generator."""
import numpy as np
import pytest

from oracle import prng, configs
from oracle.gridworld import GridWorld, EnvState, optimal_return
from oracle.rollout import RolloutWrapper, exp_portable, softmax_portable


def _state(env, p, pos, exists_bits, time=0):
    O = env.max_n_objs
    s = env.reset(None, p, 1)
    s.pos[:] = pos
    s.time[:] = time
    s.obj_existss[:] = [(exists_bits >> i) & 1 for i in range(O)]
    return s


def test_default_level_moves_and_clamps():
    env = GridWorld()
    p = env.default_params
    k = prng.split(prng.PRNGKey(0), 1)[None]
    for pos, a, want in [(0, 0, 0), (0, 2, 0), (0, 1, 11), (0, 3, 1), (10, 3, 10), (120, 1, 120),
                         (60, 4, 60), (60, 0, 49), (60, 2, 59), (110, 2, 110)]:
        s = _state(env, p, pos, 0)
        ns, r, d = env.step_env(k, s, np.array([[a]]), p)
        assert ns.pos[0, 0] == want and r[0, 0] == 0.0 and ns.time[0, 0] == 1


def test_default_level_collect_reward_and_removal():
    env = GridWorld()
    p = env.default_params
    p.obj_p_respawn[:] = 0.0
    k = prng.split(prng.PRNGKey(0), 1)[None]
    s = _state(env, p, 13, 0b1111)                       # one step left of object 0 at 1*11+3
    ns, r, d = env.step_env(k, s, np.array([[3]]), p)
    assert ns.pos[0, 0] == 14 and r[0, 0] == 1.0
    assert ns.obj_existss[0, 0].tolist() == [False, True, True, True]
    assert env.obs_index(ns)[0][0, 0] == 14 + 121 * 0b1110
    s = _state(env, p, 8 * 11 + 6, 0b1111)               # object 2 (type 1): reward -1, p_term .5
    ns, r, d = env.step_env(k, s, np.array([[3]]), p)
    assert r[0, 0] == -1.0


def test_walls_block():
    env = GridWorld()
    p = env.default_params
    p.walls[0, 1] = True
    k = prng.split(prng.PRNGKey(0), 1)[None]
    ns, _, _ = env.step_env(k, _state(env, p, 0, 0), np.array([[3]]), p)
    assert ns.pos[0, 0] == 0


def test_episode_cap_and_autoreset():
    env = GridWorld()
    p = env.default_params
    p.max_steps_in_episode[:] = 3
    k = prng.split(prng.PRNGKey(0), 1)[None]
    s = _state(env, p, 50, 0b0101, time=2)
    ns, r, d = env.step(k, s, np.array([[4]]), p)
    assert d[0, 0] and ns.time[0, 0] == 0 and ns.pos[0, 0] == 0 and ns.obj_existss[0, 0].all()
    assert not ns.early_term.any()


def test_padding_object_ids_wrap_and_never_exist():
    p, _ = configs.reset_env_params(prng.split(prng.PRNGKey(0), 4), "tabular")
    kw, ep = configs.get_env_spec("tabular")
    env = GridWorld(**kw)
    s = env.reset(None, p, 3)
    assert (s.obj_existss.sum(-1) == p.n_objs[:, None]).all()
    ro = RolloutWrapper(env, 30, ep)
    tab = np.zeros((4, env.obs_dim, 5), np.float32)
    traj, s2, _ = ro.batch_rollout(prng.split(prng.PRNGKey(1), 4), tab, p, s)
    unused = np.arange(kw["max_n_objs"])[None, None] >= p.n_objs[:, None, None]
    assert not (s2.obj_existss & unused).any()


def test_exp_portable_accuracy_and_softmax():
    x = -np.random.RandomState(0).rand(200000).astype(np.float32) * 80
    rel = np.abs(exp_portable(x) - np.exp(x.astype(np.float64))) / np.exp(x.astype(np.float64))
    assert rel.max() < 2.5 * 2 ** -24
    z = np.random.RandomState(1).randn(1000, 5).astype(np.float32) * 5
    p = softmax_portable(z)
    ref = np.exp(z - z.max(-1, keepdims=True)); ref /= ref.sum(-1, keepdims=True)
    np.testing.assert_allclose(p, ref, rtol=1e-6)


def test_rollout_first_episode_return_and_shapes():
    kw, ep = configs.get_env_spec("all_shortlife")
    env = GridWorld(**kw)
    keys = prng.split(prng.PRNGKey(0), 6)
    p, life = configs.reset_env_params(keys, "all_shortlife")
    ro = RolloutWrapper(env, 20, ep)
    tab = (np.random.RandomState(0).randn(6, env.obs_dim, 5) / np.sqrt(env.obs_dim)).astype(np.float32)
    s0 = ro.batch_reset(None, p, 8)
    traj, s1, ret = ro.batch_rollout(keys, tab, p, s0)
    assert traj.obs_idx.shape == (6, 21, 8) and traj.action.shape == (6, 20, 8)
    # first-episode return == sum of rewards up to and including the first done
    first = np.where(traj.done.any(1), traj.done.argmax(1), 19)
    mask = np.arange(20)[None, :, None] <= first[:, None, :]
    np.testing.assert_allclose(ret, (traj.reward * mask).sum(1), rtol=1e-6, atol=1e-6)
    # obs after a done is the reset observation
    full = (1 << p.n_objs) - 1
    reset_idx = p.start_pos + 100 * full
    nxt = traj.obs_idx[:, 1:]
    assert (nxt[traj.done] == np.broadcast_to(reset_idx[:, None, None], nxt.shape)[traj.done]).all()
    assert (traj.obs_time[:, 1:][traj.done] == 0).all()
    # continuing the rollout from the end state with the same key chain is a pure function
    traj2, _, _ = ro.batch_rollout(keys, tab, p, s0)
    assert (traj2.action == traj.action).all()


def test_forced_actions_are_replayed():
    kw, ep = configs.get_env_spec("debug")
    env = GridWorld(**kw)
    keys = prng.split(prng.PRNGKey(3), 3)
    p, _ = configs.reset_env_params(keys, "debug")
    ro = RolloutWrapper(env, 12, ep)
    fa = np.random.RandomState(0).randint(0, 5, (3, 12, 4))
    tab = np.zeros((3, env.obs_dim, 5), np.float32)
    traj, _, _ = ro.batch_rollout(keys, tab, p, ro.batch_reset(None, p, 4), forced_actions=fa)
    assert (traj.action == fa).all()


def test_optimal_return_bounds_policy_return():
    """gridworld.py:253-323 DP as a known-answer upper bound (debug-sized level)."""
    kw, ep = configs.get_env_spec("debug")
    env = GridWorld(**kw)
    keys = prng.split(prng.PRNGKey(11), 1)
    p, _ = configs.reset_env_params(keys, "debug")
    opt = optimal_return(env, p, ep)
    ro = RolloutWrapper(env, 20, ep)
    tab = np.zeros((1, env.obs_dim, 5), np.float32)
    s0 = ro.batch_reset(None, p, 128)
    rets = []
    for i in range(8):
        _, _, ret = ro.batch_rollout(prng.split(prng.PRNGKey(100 + i), 1), tab, p, s0, eval=True)
        rets.append(ret.mean())
    assert np.mean(rets) <= opt + 0.15
    assert opt >= 0.0
