"""Double-oracle pieces: simplex projection and the Nash solver kernel vs the numpy oracle
(util/projection.py:9-38, nash_sampler.py:39-58), and a tiny end-to-end train_do run."""
import numpy as np
import pytest
import torch

from oracle import nash as ON

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nz", [(7, 7), (16, 5), (100, 37), (1, 1), (1000, 999)])
def test_projection_simplex(built_lib, n, nz):
    from to_ued_b200.util.projection import projection_simplex
    x = np.random.RandomState(n).randn(n).astype(np.float32)
    got = projection_simplex(torch.from_numpy(x).cuda(), nz).cpu().numpy()
    want = ON.projection_simplex(x.astype(np.float64), nz)
    np.testing.assert_allclose(got, want, atol=2e-6)
    assert abs(got.sum() - 1) < 1e-5 and (got[nz:] == 0).all() and (got >= 0).all()


@pytest.mark.parametrize("n,xnz,ynz", [(6, 6, 6), (12, 5, 9)])
def test_get_nash_matches_oracle(built_lib, n, xnz, ynz):
    from to_ued_b200.environments.nash_sampler import Game, get_nash
    rs = np.random.RandomState(n)
    G = rs.randn(n, n).astype(np.float32)
    x0 = ON.projection_simplex(np.where(np.arange(n) < xnz, rs.rand(n), 0), xnz).astype(np.float32)
    y0 = ON.projection_simplex(np.where(np.arange(n) < ynz, rs.rand(n), 0), ynz).astype(np.float32)
    xo, yo = get_nash(Game(torch.from_numpy(G).cuda(), torch.from_numpy(x0).cuda(), torch.from_numpy(y0).cuda()), xnz, ynz, num_iters=300)
    wx, wy = ON.get_nash(G.astype(np.float64), x0.astype(np.float64), y0.astype(np.float64), xnz, ynz, num_iters=300)
    np.testing.assert_allclose(xo.cpu().numpy(), wx, atol=2e-4)
    np.testing.assert_allclose(yo.cpu().numpy(), wy, atol=2e-4)


def test_matching_pennies_nash(built_lib):
    from to_ued_b200.environments.nash_sampler import Game, get_nash
    G = torch.tensor([[1.0, -1.0], [-1.0, 1.0]]).cuda()
    x, y = get_nash(Game(G, torch.tensor([0.9, 0.1]).cuda(), torch.tensor([0.2, 0.8]).cuda()), 2, 2)
    np.testing.assert_allclose(x.cpu().numpy(), [0.5, 0.5], atol=0.03)
    np.testing.assert_allclose(y.cpu().numpy(), [0.5, 0.5], atol=0.03)


def test_train_do_tiny_end_to_end(built_lib):
    """train_do.py on the 'debug' distribution: 3-slot buffers, 2 agents, br=2, 1 meta-step per LPG fit."""
    import train_do
    from to_ued_b200.experiments.parse_args import parse_args
    args = parse_args(["--env_mode", "debug", "--num_agents", "2", "--num_mini_batches", "1", "--buffer_size", "3",
                       "-br", "2", "--train_steps", "1", "--num_agent_updates", "2", "--score_function", "alg_regret"])
    hist, ts, buf = train_do.make_train(args)(np.array([0, 5], np.uint32))
    torch.cuda.synchronize()
    assert len(hist) == 2 and buf.active.all()
    assert all(np.isfinite(m["GT"]["eval_regret"]) for m in hist)
    assert torch.isfinite(ts.params).all()
