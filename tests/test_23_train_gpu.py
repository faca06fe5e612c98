"""End-to-end drop-in path: train.py loop (reference train.py:14-82) incl. agent re-creation by
LevelSampler.sample, determinism, device table init and the LPG init."""
import numpy as np
import pytest
import torch

from oracle import prng

pytestmark = pytest.mark.gpu


def _run(extra=(), steps=6):
    import train
    from to_ued_b200.experiments.parse_args import parse_args
    args = parse_args(["--env_mode", "debug", "--num_agents", "4", "--num_mini_batches", "2", "--train_steps", str(steps),
                       "--num_agent_updates", "3", *extra])
    hist, ts, buf = train.make_train(args)(prng.PRNGKey(args.seed))
    torch.cuda.synchronize()
    return hist, ts, buf


def test_train_loop_runs_recreates_agents_and_is_deterministic(built_lib):
    h1, ts1, _ = _run()
    h2, ts2, _ = _run()
    assert len(h1) == 6
    for a, b in zip(h1, h2):
        assert float(a["lpg_loss"]) == float(b["lpg_loss"]) and float(a["lpg_agent_return"]) == float(b["lpg_agent_return"])
        assert np.isfinite(float(a["reg_lpg_loss"])) and np.isfinite(float(a["lpg_agent"]["policy_entropy"]))
    assert torch.equal(ts1.params, ts2.params)                 # bitwise reproducible training
    assert ts1.opt_state["count"] == 6


def test_cuda_graph_replay_matches_eager_enqueue(built_lib, monkeypatch):
    """meta/graph.py: the captured-and-replayed meta-step (default) against the eager launch-by-launch step over 8
    meta-steps that include agent re-creation by the sampler.  Same kernels, same order; the only arithmetic difference
    is Adam's bias correction computed on the device (powf) instead of on the host."""
    import to_ued_b200
    monkeypatch.setattr(to_ued_b200, "CUDA_GRAPH", True)
    hg, tg, _ = _run(steps=8)
    monkeypatch.setattr(to_ued_b200, "CUDA_GRAPH", False)
    he, te, _ = _run(steps=8)
    assert tg.opt_state["count"] == te.opt_state["count"] == 8
    for t, (a, b) in enumerate(zip(hg, he)):
        for k in ("lpg_loss", "reg_lpg_loss", "value_loss", "lpg_agent_return"):
            assert abs(float(a[k]) - float(b[k])) <= 1e-5 * max(1.0, abs(float(b[k]))), (t, k, float(a[k]), float(b[k]))
    d = float((tg.params - te.params).abs().max() / te.params.abs().max())
    assert d < 1e-6, f"graph vs eager LPG parameters differ by {d:.2e}"
    assert float(hg[0]["lpg_loss"]) == float(he[0]["lpg_loss"])       # first step: eager warm-up call, bitwise equal


def test_train_loop_frozen_and_es_paths(built_lib):
    h, ts, buf = _run(["--score_function", "frozen", "--buffer_size", "8"], steps=3)
    assert len(buf) == 8 and np.isfinite(float(h[-1]["lpg_loss"]))
    h, es, _ = _run(["--use_es", "--lifetime_conditioning", "--lpg_learning_rate", "0.01"], steps=2)
    assert np.isfinite(float(h[-1]["fitness"]["mean"])) and es.es_state["gen_counter"] == 2


def test_device_table_init_matches_host_lecun_normal(built_lib):
    from to_ued_b200.models.agent import init_tables, lecun_normal
    keys = prng.split(prng.PRNGKey(3), 5)
    D = 801
    t = init_tables(keys, D, 5)
    want = lecun_normal(keys, (D, 5), D)
    np.testing.assert_allclose(t[..., :5].cpu().numpy(), want, rtol=0, atol=2e-7)
    assert (t[..., 5:] == 0).all()
    std = t[..., :5].std().item()
    assert abs(std - 1 / np.sqrt(D)) / (1 / np.sqrt(D)) < 0.05      # lecun-normal variance
    # masked re-init only touches masked agents
    t2 = init_tables(prng.split(prng.PRNGKey(4), 5), D, 5, out=t.clone(), mask=np.array([1, 0, 0, 1, 0], np.uint8))
    assert torch.equal(t2[1], t[1]) and not torch.equal(t2[0], t[0])


def test_lpg_init_shapes_and_orthogonality(built_lib):
    from to_ued_b200.models.lpg import LPG
    m = LPG(lifetime_conditioning=True)
    flat = m.init(prng.PRNGKey(0))
    assert flat.numel() == 205482 and LPG().size == 203946          # SURVEY §8(a) parameter counts
    named = m.named_params(flat)
    hr = named["LPGGRU_0"]["GRUCell_0"]["hr"]["kernel"].cpu().double()
    np.testing.assert_allclose((hr.T @ hr).numpy(), np.eye(256), atol=1e-4)   # orthogonal recurrent init
    assert named["LPGGRU_0"]["GRUCell_0"]["ir"]["kernel"].shape == (7, 256)
    assert float(named["MLP_0"]["Dense_0"]["bias"].abs().sum()) == 0.0


def test_log_writes_checkpoints_and_restore_resumes_bitwise(built_lib, tmp_path, monkeypatch):
    """--log (reference train.py:63-69 + logging.py) and the restore the reference lacks: a restored LPG state
    continues exactly like the original (same Adam moments and step count)."""
    import os
    import train
    from to_ued_b200.experiments import logging as tlog
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.meta.meta import create_lpg_train_state
    monkeypatch.setenv("TOUED_LOG_DIR", str(tmp_path))
    argv = ["--env_mode", "debug", "--num_agents", "4", "--num_mini_batches", "2", "--train_steps", "3",
            "--num_agent_updates", "3", "--score_function", "frozen", "--buffer_size", "8", "--log", "--wandb_group", "t"]
    hist, ts, buf = train.main(argv)
    run = [d for d in os.listdir(tmp_path) if d.startswith("t-")][0]
    ck = os.path.join(tmp_path, run, "checkpoints")
    assert sorted(os.listdir(ck)) == ["buffer_3.npz", "checkpoint_3.npz"]
    args = parse_args(argv)
    fresh = create_lpg_train_state(prng.split(prng.PRNGKey(123), 3)[1], args)
    back = tlog.restore_checkpoint(os.path.join(ck, "checkpoint_3.npz"), fresh)
    assert torch.equal(back.params, ts.params) and back.opt_state["count"] == 3 and back.step == ts.step
    assert torch.equal(back.opt_state["mu"], ts.opt_state["mu"])
    b2 = tlog.restore_buffer(os.path.join(ck, "buffer_3.npz"))
    assert np.array_equal(b2.score, buf.score) and torch.equal(b2.level.packed.cpu(), buf.level.packed.cpu())   # device-resident records
    # one more Adam step from the restored and from the original state gives identical parameters
    g = torch.randn_like(ts.params)
    p1, p2 = ts.params.clone(), back.params.clone()
    s1 = ts.tx.update_(p1, g, {**ts.opt_state, "mu": ts.opt_state["mu"].clone(), "nu": ts.opt_state["nu"].clone()})
    s2 = back.tx.update_(p2, g, {**back.opt_state, "mu": back.opt_state["mu"].clone(), "nu": back.opt_state["nu"].clone()})
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and s1["count"] == s2["count"] == 4


def test_cli_entry_points_print_metrics(built_lib, capsys):
    """train.py / train_do.py main(): the printed metric trees (ES metrics hold nested dicts of device scalars)."""
    import train
    import train_do
    train.main(["--env_mode", "debug", "--num_agents", "4", "--num_mini_batches", "1", "--train_steps", "1",
                "--num_agent_updates", "2", "--use_es", "--lifetime_conditioning"])
    out = capsys.readouterr().out
    assert "'fitness': {'mean':" in out and "tensor(" not in out
    train_do.main(["--env_mode", "debug", "--num_agents", "2", "--num_mini_batches", "1", "--buffer_size", "3", "-br", "2",
                   "--train_steps", "1", "--num_agent_updates", "2", "--score_function", "alg_regret"])
    out = capsys.readouterr().out
    assert "'eval_regret':" in out and "tensor(" not in out
