"""N > 1 host logic on CPU with gloo (world_size 2), product code under test (SURVEY.md sections 4, 8e):

* ``LevelSampler._plan_sample`` / ``initial_sample`` key and level derivation: rank r's plan must equal rows
  [r n, (r+1) n) of the single-rank plan for the global batch, and the replicated level buffer must end up
  identical on every rank (``score_function`` random / frozen / alg_regret incl. Q3), over several lifetimes;
* ``train._shard`` slices every per-agent field consistently;
* the mean-reduce of the meta-gradient over agent shards reproduces the single-rank gradient (per-shard gradient from
  the oracle, reduction over gloo, keys from ``to_ued_b200.util.prng``);
* ES: the shard-wise gradient estimate all-reduced over ranks equals the single-rank ``tell`` input."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_GLOBAL = 8


def _setup(n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import Case
    return Case("debug", n=n, w=8, L=6, seed=2, table_scale=0.3)


def _shard_grad(c, sl, n_global, trajs, ev):
    from oracle import prng
    from oracle.agents import AgentTables
    from oracle.meta import lpg_meta_grad_train_step
    from oracle.rollout import Trajectory
    dt = torch.float64
    pick = lambda t: Trajectory(t.obs_idx[sl], t.obs_time[sl], t.action[sl], t.reward[sl], t.done[sl])
    ag = AgentTables(torch.tensor(c.actor[sl]).to(dt), torch.tensor(c.critic[sl]).to(dt),
                     torch.zeros(len(c.actor[sl]), dtype=torch.long))
    s0 = c.oro.batch_reset(None, c.p.index(sl), c.w)
    o = lpg_meta_grad_train_step(prng.PRNGKey(4), c.layout, torch.tensor(c.lpg).to(dt), ag,
                                 torch.tensor(c.value[sl]).to(dt), c.oro, c.p.index(sl), s0, c.life[sl],
                                 num_agent_updates=2, trajectories=[pick(t) for t in trajs], eval_trajectory=pick(ev),
                                 do_eval=False)
    n_local = len(c.actor[sl])
    return o["grad"] * n_local / n_global            # oracle returns the mean over its own agents


# ----------------------------------------------------------------------------------------------------------------
# sampler plan: a driver that exercises the host side of LevelSampler over several meta-steps without a GPU
def _sampler_args(score_function):
    from to_ued_b200.experiments.parse_args import parse_args
    return parse_args(["--env_mode", "all_vrandlife", "--num_agents", str(N_GLOBAL), "--score_function", score_function,
                       "--buffer_size", "64", "--num_mini_batches", "1"])


def _fake_score(keys, only, ids):
    """Stand-in for the device regret evaluation: a deterministic function of the agent's key and level id."""
    s = ((keys[:, 0] % 1000).astype(np.float32) / 1000.0 + (ids % 7).astype(np.float32)).astype(np.float32)
    return np.where(only, s, 0.0).astype(np.float32)


def _drive_sampler(score_function, rank, world, steps=12):
    """Run ``steps`` meta-steps of host-side sampler logic for this rank; lifetimes are shortened so that agents
    terminate at different times.  Returns a log of everything that must be rank-independent."""
    from to_ued_b200.environments.level_sampler import LevelSampler, _index_level
    from to_ued_b200.meta.train import _advance_host_step
    from to_ued_b200.util import prng, dist as udist
    args = _sampler_args(score_function)
    ls = LevelSampler(args, device="cpu")
    assert (ls.rank, ls.world) == (rank, world)
    rng = prng.PRNGKey(11)
    rng, brng = prng.split(rng, 2)
    buf = ls.initialize_buffer(brng)
    sl = udist.local_slice(N_GLOBAL)
    # initial levels exactly as initial_sample derives them (host part)
    rng, _rng = prng.split(rng, 2)
    if score_function == "random":
        _rng2 = prng.split(_rng, 2)[1]
        level = ls._sample_random_levels(_rng2, N_GLOBAL, sl)
    else:
        level = _index_level(buf.level, np.arange(N_GLOBAL)[sl])
        buf = buf.replace(active=np.arange(ls.buffer_size) < N_GLOBAL)
    level.lifetime = (level.lifetime % 23 + 3).astype(np.int32)          # short, different lifetimes
    host_step = np.zeros(len(level), np.int32)
    log = []
    for t in range(steps):
        host_step = _advance_host_step(host_step, level.lifetime, 5)
        rng, _rng = prng.split(rng, 2)
        score_fn = lambda keys, only: _fake_score(keys, only, level.buffer_id)
        buf, plan = ls._plan_sample(_rng, buf, host_step, level, True, score_fn)
        entry = {"t": t, "terminated": None}
        if plan is not None:
            term, new_levels, akeys, vkeys = plan
            new_levels.lifetime = np.where(term, new_levels.lifetime % 23 + 3, new_levels.lifetime).astype(np.int32)
            level = new_levels
            host_step = np.where(term, 0, host_step).astype(np.int32)
            entry.update(terminated=term.copy(), akeys=akeys.copy(), vkeys=vkeys.copy())
        entry.update(lifetime=level.lifetime.copy(), buffer_id=level.buffer_id.copy(),
                     walls=np.asarray(level.env_params.walls).copy(), start=np.asarray(level.env_params.start_pos).copy())
        if buf is not None:
            entry.update(score=buf.score.copy(), active=buf.active.copy(), new=buf.new.copy(),
                         buf_life=buf.level.lifetime.copy())
        log.append(entry)
    return log


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sys.path.insert(0, ROOT)
    # ---- (1) meta-gradient shard sums, keys from the product's prng ----
    c = _setup(4)
    trajs, ev = torch.load(out + ".traj", weights_only=False)
    n_local = 4 // world
    from to_ued_b200.util import prng as P
    from oracle import prng as O
    keys = P.split(P.PRNGKey(4), 4)[rank * n_local:(rank + 1) * n_local]
    np.testing.assert_array_equal(keys, O.split(O.PRNGKey(4), 4)[rank * n_local:(rank + 1) * n_local])
    g = _shard_grad(c, slice(rank * n_local, (rank + 1) * n_local), 4, trajs, ev)
    dist.all_reduce(g)                                # sum of 1/N_global-scaled shard sums
    # ---- (2) sampler plans under world_size 2 ----
    logs = {sf: _drive_sampler(sf, rank, world) for sf in ("random", "frozen", "alg_regret")}
    # ---- (3) host all-gather helper ----
    from to_ued_b200.util import dist as udist
    got = udist.all_gather_host(np.arange(3, dtype=np.int32) + 10 * rank)
    np.testing.assert_array_equal(got, np.concatenate([np.arange(3) + 10 * r for r in range(world)]))
    gb = udist.all_gather_host(np.array([rank == 0, True]))
    assert gb.dtype == np.bool_ and gb.tolist() == [True, True, False, True]
    torch.save({"g": g, "logs": logs}, f"{out}.rank{rank}")
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_match_single_rank(tmp_path):
    from oracle import prng
    from oracle.agents import AgentTables
    from oracle.meta import lpg_meta_grad_train_step
    c = _setup(4)
    dt = torch.float64
    ag = AgentTables(torch.tensor(c.actor).to(dt), torch.tensor(c.critic).to(dt), torch.zeros(4, dtype=torch.long))
    s0 = c.oro.batch_reset(None, c.p, c.w)
    full = lpg_meta_grad_train_step(prng.PRNGKey(4), c.layout, torch.tensor(c.lpg).to(dt), ag,
                                    torch.tensor(c.value).to(dt), c.oro, c.p, s0, c.life, num_agent_updates=2,
                                    do_eval=False)
    out = str(tmp_path / "g.pt")
    torch.save((full["trajectories"], full["eval_trajectory"]), out + ".traj")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = [torch.load(f"{out}.rank{r}", weights_only=False) for r in range(2)]
    g2 = res[0]["g"]
    assert float((g2 - full["grad"]).abs().max() / full["grad"].abs().max()) < 1e-12
    # ---- sampler: rank r's log == rows [r n, (r+1) n) of the single-rank log; buffers identical everywhere ----
    n = N_GLOBAL // 2
    for sf in ("random", "frozen", "alg_regret"):
        one = _drive_sampler(sf, 0, 1)
        n_term = 0
        for t, e1 in enumerate(one):
            for r in range(2):
                e2 = res[r]["logs"][sf][t]
                sl = slice(r * n, (r + 1) * n)
                assert (e1["terminated"] is None) == (e2["terminated"] is None), (sf, t)
                for k in ("lifetime", "buffer_id", "walls", "start"):
                    np.testing.assert_array_equal(e2[k], e1[k][sl], err_msg=f"{sf} step {t} rank {r} {k}")
                if e1["terminated"] is not None:
                    n_term += int(e1["terminated"][sl].sum())
                    for k in ("terminated", "akeys", "vkeys"):
                        np.testing.assert_array_equal(e2[k], e1[k][sl], err_msg=f"{sf} step {t} rank {r} {k}")
                for k in ("score", "active", "new", "buf_life"):
                    if k in e1:
                        np.testing.assert_array_equal(e2[k], e1[k], err_msg=f"{sf} step {t} rank {r} buffer {k}")
        assert n_term >= N_GLOBAL, f"{sf}: the driver must see at least one re-creation per agent ({n_term})"
    # two agents of different ranks never share a freshly drawn level (the round-1 defect: same key on every rank)
    e = res[0]["logs"]["random"][-1]["walls"], res[1]["logs"]["random"][-1]["walls"]
    assert not np.array_equal(e[0], e[1])


def test_shard_slices_every_field():
    """train._shard on CPU tensors: rank r gets rows [r n, (r+1) n) of every per-agent field."""
    sys.path.insert(0, ROOT)
    import train as train_mod
    from to_ued_b200.util.data import AgentState, TrainState, Level
    from to_ued_b200.environments.gridworld.gridworld import EnvState
    from oracle import prng, configs
    n, D, W = 6, 11, 4
    p, life = configs.reset_env_params(prng.split(prng.PRNGKey(0), n), "debug")
    from to_ued_b200.environments.gridworld.gridworld import EnvParams
    pp = EnvParams(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    ar = lambda *s: torch.arange(int(np.prod(s)), dtype=torch.float32).reshape(*s)
    ag = AgentState(TrainState(ar(n, D, 8), torch.arange(n, dtype=torch.int32), 5, 1.0, 0.5),
                    TrainState(ar(n, D, 8) + 0.5, torch.arange(n, dtype=torch.int32), 8, 1.0, 0.5),
                    Level(pp, life, np.arange(n, dtype=np.int32), torch.arange(n * 192, dtype=torch.uint8).reshape(n, 192)),
                    torch.arange(n * W, dtype=torch.int32).reshape(n, W), EnvState(torch.arange(n * W, dtype=torch.int32).reshape(n, W), 5),
                    np.arange(n, dtype=np.int32))
    vc = TrainState(ar(n, D, 8) + 0.25, torch.zeros(n, dtype=torch.int32), 1, 1.0, 0.5)
    for r in range(3):
        a, v = train_mod._shard(ag, vc, r, 2)
        sl = slice(2 * r, 2 * r + 2)
        assert torch.equal(a.actor_state.params, ag.actor_state.params[sl]) and torch.equal(a.critic_state.params, ag.critic_state.params[sl])
        assert torch.equal(a.actor_state.step, ag.actor_state.step[sl]) and torch.equal(a.env_obs, ag.env_obs[sl])
        assert torch.equal(a.env_state.packed, ag.env_state.packed[sl]) and torch.equal(a.level.packed, ag.level.packed[sl])
        np.testing.assert_array_equal(a.level.lifetime, life[sl]); np.testing.assert_array_equal(a.level.buffer_id, np.arange(n)[sl])
        np.testing.assert_array_equal(a.host_step, np.arange(n)[sl]); assert torch.equal(v.params, vc.params[sl])


def test_host_step_mirror_matches_masked_updates():
    from to_ued_b200.meta.train import _advance_host_step
    step = np.array([0, 3, 248, 250, 7], np.int32)
    life = np.array([250, 4, 250, 250, 5], np.int32)
    want = step.copy()
    for _ in range(5):
        want = np.where(want + 1 <= life, want + 1, want)
    np.testing.assert_array_equal(_advance_host_step(step, life, 5), want)
