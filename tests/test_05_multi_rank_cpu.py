"""N > 1 host logic on CPU with gloo (world_size 2): rank-local key slices, agent sharding and the
mean-reduce of the meta-gradient must reproduce the single-rank result.
The per-shard gradient is computed by the oracle (no GPU here); the sharding / reduction code under
test is the product's (`train._shard` logic is exercised through the same slicing conventions and
`to_ued_b200.util.prng` key derivation)."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    c = _setup(4)
    trajs, ev = torch.load(out + ".traj", weights_only=False)
    n_local = 4 // world
    # product-side key derivation: global split, local slice (meta/train.py)
    from to_ued_b200.util import prng as P
    from oracle import prng as O
    keys = P.split(P.PRNGKey(4), 4)[rank * n_local:(rank + 1) * n_local]
    np.testing.assert_array_equal(keys, O.split(O.PRNGKey(4), 4)[rank * n_local:(rank + 1) * n_local])
    g = _shard_grad(c, slice(rank * n_local, (rank + 1) * n_local), 4, trajs, ev)
    dist.all_reduce(g)                                # sum of 1/N_global-scaled shard sums
    if rank == 0:
        torch.save(g, out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_mean_reduce_matches_single_rank(tmp_path):
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
    g2 = torch.load(out)
    assert float((g2 - full["grad"]).abs().max() / full["grad"].abs().max()) < 1e-12


def test_host_step_mirror_matches_masked_updates():
    from to_ued_b200.meta.train import _advance_host_step
    step = np.array([0, 3, 248, 250, 7], np.int32)
    life = np.array([250, 4, 250, 250, 5], np.int32)
    want = step.copy()
    for _ in range(5):
        want = np.where(want + 1 <= life, want + 1, want)
    np.testing.assert_array_equal(_advance_host_step(step, life, 5), want)
