"""Device level generator (csrc/levelgen.cu; reference environments/gridworld/configs.py:12-57) against the oracle's
per-level numpy restatement: the packed LevelRec bytes must be IDENTICAL for every registered distribution -- integer
fields, wall masks, positions, and the float tables (log-uniform draws go through exp_portable on both sides)."""
import numpy as np
import pytest
import torch

from oracle import prng, configs as oconf

pytestmark = pytest.mark.gpu

MODES = ["all_shortlife", "all_vrandlife", "all_randlife", "small", "medium", "large", "debug", "tabular", "mazes",
         "dense", "longer", "long_dense"]


def _oracle_records(keys, mode, buffer_ids):
    from to_ued_b200.environments.gridworld.gridworld import EnvParams, pack_levels
    p, life = oconf.reset_env_params(keys, mode)
    pp = EnvParams(**{k: getattr(p, k) for k in p.__dataclass_fields__})
    return pack_levels(pp, life, buffer_ids), life


@pytest.mark.parametrize("mode", MODES)
def test_device_levels_are_bit_identical_to_the_oracle(built_lib, mode):
    from to_ued_b200.environments.gridworld.levelgen import generate_levels
    n = 96
    keys = prng.split(prng.PRNGKey(sum(map(ord, mode)) + 5), n)
    ids = (np.arange(n, dtype=np.int32) * 7) % 50
    want, life = _oracle_records(keys, mode, ids)
    got, glife = generate_levels(keys, mode, buffer_ids=ids, want_lifetimes=True)
    got = got.cpu().numpy().reshape(n, 192)
    want_b = want.view(np.uint8).reshape(n, 192)
    np.testing.assert_array_equal(glife.cpu().numpy(), life)
    if not np.array_equal(got, want_b):
        bad = np.nonzero((got != want_b).any(1))[0]
        i = int(bad[0])
        names = want.dtype.names
        rec = got[i].view(want.dtype)[0]
        diff = [f"{nm}: {rec[nm]} vs {want[i][nm]}" for nm in names if not np.array_equal(rec[nm], want[i][nm])]
        raise AssertionError(f"{mode}: {len(bad)} of {n} levels differ; level {i}: " + "; ".join(diff))


def test_sampler_device_levels_match_host_levels(built_lib):
    """LevelSampler with device-resident levels (default on CUDA) draws the same levels as the host path."""
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.level_sampler import LevelSampler
    from to_ued_b200.environments.gridworld.gridworld import pack_levels
    args = parse_args(["--env_mode", "all_vrandlife", "--num_agents", "16", "--num_mini_batches", "1"])
    dev_s, host_s = LevelSampler(args), LevelSampler(args)
    host_s.device_levels = False
    rng = prng.PRNGKey(9)
    a = dev_s._sample_random_levels(rng, 16)
    b = host_s._sample_random_levels(rng, 16)
    np.testing.assert_array_equal(a.lifetime, b.lifetime)
    want = pack_levels(b.env_params, b.lifetime, b.buffer_id).view(np.uint8).reshape(16, 192)
    np.testing.assert_array_equal(a.packed.cpu().numpy(), want)
    assert a.env_params is None
