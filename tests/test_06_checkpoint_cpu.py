"""Checkpoint save / restore round trip (experiments/logging.py; reference logging.py:25-46 saves, never restores)."""
import os
from types import SimpleNamespace

import numpy as np
import torch


def _train_state(n=257, seed=0):
    from to_ued_b200.meta.train import LPGTrainState
    from to_ued_b200.models.optim import Adam
    g = torch.Generator().manual_seed(seed)
    params = torch.randn(n, generator=g)
    ts = LPGTrainState(model=SimpleNamespace(lifetime_conditioning=False), params=params, tx=Adam(1e-4))
    ts.opt_state["mu"].copy_(torch.randn(n, generator=g))
    ts.opt_state["nu"].copy_(torch.rand(n, generator=g))
    ts.opt_state["count"] = 7
    return ts.replace(step=7)


def test_train_state_round_trip(tmp_path):
    from to_ued_b200.experiments import logging as tlog
    ts = _train_state()
    path = tlog.save_checkpoint(str(tmp_path), ts, 7)
    assert os.path.basename(path) == "checkpoint_7.npz"
    fresh = _train_state(seed=1)
    assert not torch.equal(fresh.params, ts.params)
    back = tlog.restore_checkpoint(path, fresh)
    assert torch.equal(back.params, ts.params) and back.step == 7 and back.opt_state["count"] == 7
    assert torch.equal(back.opt_state["mu"], ts.opt_state["mu"]) and torch.equal(back.opt_state["nu"], ts.opt_state["nu"])
    # keep=1: a newer checkpoint replaces the old one
    tlog.save_checkpoint(str(tmp_path), ts, 9)
    assert sorted(os.listdir(tmp_path)) == ["checkpoint_9.npz"]
    # a model of a different size is refused
    try:
        tlog.restore_checkpoint(os.path.join(tmp_path, "checkpoint_9.npz"), _train_state(n=100))
    except ValueError:
        pass
    else:
        raise AssertionError("size mismatch not detected")


def test_level_buffer_round_trip_and_log_results(tmp_path, monkeypatch):
    from oracle import prng as oprng
    from to_ued_b200.experiments import logging as tlog
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.environments import reset_env_params
    from to_ued_b200.environments.level_sampler import LevelBuffer
    args = parse_args(["--env_mode", "all_shortlife", "--num_agents", "4", "--num_mini_batches", "1", "--train_steps", "3"])
    keys = oprng.split(oprng.PRNGKey(3), 16)
    params, lifetimes = reset_env_params(keys, "GridWorld-v0", "all_shortlife")
    buf = LevelBuffer.create_buffer(params, lifetimes)
    buf = buf.replace(score=np.linspace(-1, 1, 16).astype(np.float32), active=np.arange(16) % 3 == 0)
    monkeypatch.setenv("TOUED_LOG_DIR", str(tmp_path))
    run = tlog.init_logger(args)
    metrics = [{"lpg_loss": torch.tensor(0.5 * i), "lpg_agent": {"policy_l2": 1.0}, "_grad": None} for i in range(3)]
    paths = tlog.log_results(args, metrics, _train_state(), buf)
    assert [os.path.basename(p) for p in paths] == ["checkpoint_3.npz", "buffer_3.npz"]
    lines = open(os.path.join(run, "metrics.jsonl")).read().strip().split("\n")
    assert len(lines) == 3 and '"lpg_loss": 1.0' in lines[2] and "_grad" not in lines[0]
    back = tlog.restore_buffer(paths[1])
    assert np.array_equal(back.score, buf.score) and np.array_equal(back.active, buf.active) and np.array_equal(back.new, buf.new)
    assert np.array_equal(back.level.lifetime, buf.level.lifetime)
    for f in ("walls", "obj_rewards", "start_pos", "grid_size"):
        assert np.array_equal(getattr(back.level.env_params, f), getattr(buf.level.env_params, f)), f
