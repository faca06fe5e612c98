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
