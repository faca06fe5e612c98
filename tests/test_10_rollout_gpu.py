"""GPU parity: fused CUDA rollout (through the C ABI) vs the CPU oracle, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import prng, configs
from oracle.gridworld import GridWorld as OGrid
from oracle.rollout import RolloutWrapper as ORollout

pytestmark = pytest.mark.gpu


def _setup(mode, n, seed):
    from to_ued_b200.environments.gridworld.gridworld import EnvParams, pack_levels, levels_to_device
    keys = prng.split(prng.PRNGKey(seed), n)
    p, life = configs.reset_env_params(keys, mode)
    kw, ep = configs.get_env_spec(mode)
    oenv = OGrid(**kw)
    D = oenv.obs_dim
    tab = (np.random.RandomState(seed).randn(n, D, 5) * 2.0 / np.sqrt(1.0)).astype(np.float32)
    pp = EnvParams(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    lv = levels_to_device(pack_levels(pp, life))
    tab8 = np.zeros((n, D, 8), np.float32); tab8[..., :5] = tab
    return keys, p, kw, ep, oenv, tab, lv, torch.from_numpy(tab8).cuda()


def _unpack_obs(o):
    o = o.cpu().numpy()
    return o & 0xFFFF, (o >> 16) & 0xFFFF


@pytest.mark.parametrize("mode,n,w,L", [("all_shortlife", 24, 64, 20), ("debug", 7, 64, 20),
                                        ("tabular", 6, 64, 20), ("mazes", 5, 64, 20),
                                        ("all_shortlife", 40, 4, 100), ("small", 3, 32, 33)])
def test_rollout_bit_exact(built_lib, mode, n, w, L):
    from to_ued_b200.environments.rollout import RolloutWrapper
    keys, p, kw, ep, oenv, tab, lv, tab8 = _setup(mode, n, 5)
    oro = ORollout(oenv, L, ep)
    s0 = oro.batch_reset(None, p, w)
    otraj, os1, oret = oro.batch_rollout(keys, tab, p, s0)
    # continue a second rollout from the end state with fresh keys (carried state path)
    keys2 = prng.split(prng.PRNGKey(99), n)
    otraj2, os2, oret2 = oro.batch_rollout(keys2, tab, p, os1)

    ro = RolloutWrapper("GridWorld-v0", L, ep, kw)
    obs0, st0 = ro.batch_reset(None, lv, w)
    traj, end_obs, st1, ret = ro.batch_rollout(keys, tab8, lv, obs0, st0)
    traj2, _, st2, ret2 = ro.batch_rollout(keys2, tab8, lv, end_obs, st1)
    torch.cuda.synchronize()
    for t, o, r, orr in ((traj, otraj, ret, oret), (traj2, otraj2, ret2, oret2)):
        idx, tm = _unpack_obs(t.obs)
        np.testing.assert_array_equal(t.action.cpu().numpy(), o.action)
        np.testing.assert_array_equal(idx, o.obs_idx)
        np.testing.assert_array_equal(tm, o.obs_time)
        np.testing.assert_array_equal(t.done.cpu().numpy().astype(bool), o.done)
        np.testing.assert_array_equal(t.reward.cpu().numpy(), o.reward)          # bit-exact f32
        np.testing.assert_array_equal(r.cpu().numpy(), orr)
    np.testing.assert_array_equal(st2.pos.cpu().numpy(), os2.pos)
    np.testing.assert_array_equal(st2.time.cpu().numpy(), os2.time)
    np.testing.assert_array_equal(st2.obj_existss.cpu().numpy(), os2.obj_existss)


def test_forced_action_stream_and_eval_mode(built_lib):
    from to_ued_b200.environments.rollout import RolloutWrapper
    keys, p, kw, ep, oenv, tab, lv, tab8 = _setup("all_shortlife", 9, 1)
    fa = np.random.RandomState(0).randint(0, 5, (9, 20, 64)).astype(np.uint8)
    oro = ORollout(oenv, 20, ep)
    otraj, _, oret = oro.batch_rollout(keys, tab, p, oro.batch_reset(None, p, 64), forced_actions=fa)
    ro = RolloutWrapper("GridWorld-v0", 20, ep, kw)
    obs0, st0 = ro.batch_reset(None, lv, 64)
    traj, _, _, ret = ro.batch_rollout(keys, tab8, lv, obs0, st0, forced_actions=fa)
    np.testing.assert_array_equal(traj.action.cpu().numpy(), fa)
    np.testing.assert_array_equal(_unpack_obs(traj.obs)[0], otraj.obs_idx)
    np.testing.assert_array_equal(traj.reward.cpu().numpy(), otraj.reward)
    # eval rollout (max_rollout_len steps, 4 workers, returns only) == agents.py:98-106 eval_agent core
    o4 = oro.batch_reset(None, p, 4)
    _, _, oret4 = oro.batch_rollout(keys, tab, p, o4, eval=True)
    obs4, st4 = ro.batch_reset(None, lv, 4)
    _, _, _, ret4 = ro.batch_rollout(keys, tab8, lv, obs4, st4, eval=True, want_trajectory=False)
    np.testing.assert_array_equal(ret4.cpu().numpy(), oret4)


def test_gymnax_step_reset_api(built_lib):
    from to_ued_b200.environments.gridworld.gridworld import GridWorld
    keys, p, kw, ep, oenv, tab, lv, tab8 = _setup("all_shortlife", 11, 2)
    env = GridWorld(**kw)
    obs, st = env.reset(None, lv, 16)
    os_ = oenv.reset(None, p, 16)
    rs = np.random.RandomState(0)
    for it in range(30):
        k = prng.split(prng.split(prng.PRNGKey(it), 11), 16)          # [11, 16, 2]
        a = rs.randint(0, 5, (11, 16))
        obs, st, r, d, _ = env.step(k, st, a, lv)
        os_, orr, od = oenv.step(k, os_, a, p)
        np.testing.assert_array_equal(st.pos.cpu().numpy(), os_.pos)
        np.testing.assert_array_equal(st.time.cpu().numpy(), os_.time)
        np.testing.assert_array_equal(st.obj_existss.cpu().numpy(), os_.obj_existss)
        np.testing.assert_array_equal(r.cpu().numpy(), orr)
        np.testing.assert_array_equal(d.cpu().numpy(), od)
    dense = env.dense_obs(obs)
    np.testing.assert_array_equal(dense.cpu().numpy(), oenv.dense_obs(os_))


def test_full_size_properties(built_lib):
    """BASELINE configs[1] shape (512 agents x 64 workers x 20 steps): size-independent checks."""
    from to_ued_b200.environments.rollout import RolloutWrapper
    keys, p, kw, ep, oenv, tab, lv, tab8 = _setup("all_shortlife", 512, 0)
    ro = RolloutWrapper("GridWorld-v0", 20, ep, kw)
    obs0, st0 = ro.batch_reset(None, lv, 64)
    traj, end_obs, st1, ret = ro.batch_rollout(keys, tab8, lv, obs0, st0)
    traj_b, _, st1_b, ret_b = ro.batch_rollout(keys, tab8, lv, obs0, st0)
    assert torch.equal(traj.action, traj_b.action) and torch.equal(ret, ret_b)      # deterministic
    idx, tm = _unpack_obs(traj.obs)
    done = traj.done.cpu().numpy().astype(bool)
    rew = traj.reward.cpu().numpy()
    assert traj.action.max().item() <= 4
    # time advances by one or resets to zero exactly where done
    assert ((tm[:, 1:] == tm[:, :-1] + 1) | done).all() and (tm[:, 1:][done] == 0).all()
    # agent never stands on a wall or outside its grid
    pos = idx % 100
    assert (pos < (p.grid_size ** 2)[:, None, None]).all()
    assert not np.take_along_axis(p.walls, pos.reshape(512, -1), 1).any()
    # first-episode return identity
    first = np.where(done.any(1), done.argmax(1), 19)
    mask = np.arange(20)[None, :, None] <= first[:, None, :]
    np.testing.assert_allclose(ret.cpu().numpy(), (rew * mask).sum(1), rtol=1e-5, atol=1e-6)
    # a slice of the big batch equals the oracle on that slice
    sl = slice(100, 104)
    oro = ORollout(oenv, 20, ep)
    otraj, _, _ = oro.batch_rollout(keys[sl], tab[sl], p.index(sl), oro.batch_reset(None, p.index(sl), 64))
    np.testing.assert_array_equal(traj.action[sl].cpu().numpy(), otraj.action)
