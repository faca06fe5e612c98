"""LPG optimisers on the device (reference models/optim.py:5-17): ``--lpg_opt SGD`` = clip_by_global_norm -> scale(lr) ->
scale(-1) (toued_sgd_clip) against the oracle, below and above the clipping threshold, and through the product's
train-state / train-step plumbing (eager and captured-graph step)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scale", [1e-3, 5.0])
def test_sgd_clip_matches_oracle(built_lib, scale):
    from oracle.meta import SGDClip
    from to_ued_b200.models.optim import create_optimizer
    rs = np.random.RandomState(3)
    n = 204_049
    p0 = rs.randn(n).astype(np.float32)
    g = (rs.randn(n) * scale / np.sqrt(n)).astype(np.float32)            # |g| ~ scale: below / above max_norm = 0.5
    tx = create_optimizer("SGD", 1e-2, 0.5)
    params, grad = torch.from_numpy(p0).cuda(), torch.from_numpy(g).cuda()
    st = tx.init(params)
    st = tx.update_(params, grad, st)
    want = SGDClip(1e-2, 0.5).step(torch.from_numpy(p0).double(), torch.from_numpy(g).double()).numpy()
    got = params.cpu().numpy()
    assert st["count"] == 1
    assert abs(float(st["mu"][0]) - float((g.astype(np.float64) ** 2).sum())) < 1e-5 * scale * scale
    assert np.abs(got - want).max() < 2e-7 * max(1.0, np.abs(want).max())
    clipped = np.linalg.norm(g) >= 0.5
    assert clipped == (scale > 1.0)


@pytest.mark.parametrize("graph", [False, True])
def test_train_step_with_sgd_optimizer(built_lib, monkeypatch, graph):
    """Three meta-steps with --lpg_opt SGD (eager steps / captured-graph replay): every step moves the parameters by at most
    lr * max_norm in L2."""
    import to_ued_b200
    import train
    from oracle.meta import SGDClip
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.meta.meta import create_lpg_train_state
    from to_ued_b200.util import prng as P
    monkeypatch.setattr(to_ued_b200, "CUDA_GRAPH", graph)
    args = parse_args(["--env_mode", "all_shortlife", "--num_agents", "4", "--num_mini_batches", "1", "--train_steps", "3",
                       "--lpg_opt", "SGD", "--lpg_learning_rate", "0.05", "--lpg_max_grad_norm", "0.001"])
    ts0 = create_lpg_train_state(P.split(P.PRNGKey(0), 3)[1], args)
    assert ts0.tx.name == "SGD"
    p0 = ts0.params.clone()
    hist, ts, _ = train.make_train(args)(P.PRNGKey(0))
    torch.cuda.synchronize()
    d = (ts.params - p0).double().cpu().numpy()
    assert np.isfinite(d).all() and np.abs(d).max() > 0
    assert np.linalg.norm(d) <= 3 * 0.05 * 0.001 * (1 + 1e-5)
    assert np.isfinite([float(h["lpg_loss"]) for h in hist]).all()
