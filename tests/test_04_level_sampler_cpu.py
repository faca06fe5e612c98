"""PLR buffer index logic (environments/level_sampler.py:183-234, 331-408): the product's host code
against the oracle restatement — bit-exact ids / flags (incl. quirk Q3)."""
import numpy as np
import pytest

from oracle import prng
from oracle import level_sampler as O


def _sampler(buffer_size=64, mode="debug", extra=()):
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.level_sampler import LevelSampler
    args = parse_args(["--env_mode", mode, "--num_agents", "8", "--num_mini_batches", "1", "--score_function",
                       "alg_regret", "--buffer_size", str(buffer_size), *extra])
    return LevelSampler(args, device="cpu")


def _random_buffer(ls, seed):
    buf = ls.initialize_buffer(prng.PRNGKey(seed))
    rs = np.random.RandomState(seed)
    B = ls.buffer_size
    buf = buf.replace(score=rs.randn(B).astype(np.float32), active=rs.rand(B) < 0.2, new=rs.rand(B) < 0.4)
    return buf


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_reset_lowest_scoring_matches_oracle_including_q3(seed):
    ls = _sampler()
    buf = _random_buffer(ls, seed)
    out = ls._reset_lowest_scoring(prng.PRNGKey(seed + 10), buf, 8)
    ids, sc, ac, nw = O.reset_lowest_scoring(buf.score, buf.active, buf.new, 8)
    np.testing.assert_array_equal(out.score, sc)
    np.testing.assert_array_equal(out.active, ac)
    np.testing.assert_array_equal(out.new, nw)                 # Q3: rebuilt from `active`
    np.testing.assert_array_equal(out.level.buffer_id[ids], ids)
    # the reset slots hold freshly generated levels drawn from split(rng, minimum_new)
    from to_ued_b200.environments.environments import reset_env_params
    p, life = reset_env_params(prng.split(prng.PRNGKey(seed + 10), 8), "GridWorld-v0", "debug")
    np.testing.assert_array_equal(out.level.env_params.start_pos[ids], p.start_pos)
    np.testing.assert_array_equal(out.level.lifetime[ids], life)


@pytest.mark.parametrize("seed", [0, 3])
def test_replay_and_random_ids_match_oracle(seed):
    ls = _sampler()
    buf = _random_buffer(ls, seed)
    rep = ls._replay_from_buffer(prng.PRNGKey(5), buf, 8)
    np.testing.assert_array_equal(rep.buffer_id, O.replay_ids(buf.score, buf.active, buf.new, 8))
    rnd = ls._sample_random_from_buffer(prng.PRNGKey(6), buf, 8)
    want = O.random_ids(prng.PRNGKey(6), buf.active, buf.new, 8)
    np.testing.assert_array_equal(rnd.buffer_id, want)
    assert (buf.new[want] & ~buf.active[want]).all() and len(set(want.tolist())) == 8


def test_not_enough_replayable_levels_falls_back_to_uniform_scores():
    ls = _sampler(buffer_size=16)
    buf = _random_buffer(ls, 1)
    buf = buf.replace(new=np.ones(16, bool))                   # nothing evaluated yet
    rep = ls._replay_from_buffer(prng.PRNGKey(5), buf, 8)
    np.testing.assert_array_equal(rep.buffer_id, O.replay_ids(buf.score, buf.active, buf.new, 8))


def test_score_function_and_transform_validation():
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.level_sampler import LevelSampler
    with pytest.raises(ValueError):
        LevelSampler(parse_args(["--score_function", "positive_value_loss", "--num_mini_batches", "1"]), device="cpu")
    with pytest.raises(ValueError):
        LevelSampler(parse_args(["--score_transform", "bogus", "--num_mini_batches", "1"]), device="cpu")
    with pytest.raises(ValueError):
        parse_args(["--num_agents", "10", "--num_mini_batches", "3"])
