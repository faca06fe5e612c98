"""Precision of the PRODUCTION path (tcgen05 GRU: fp16 forward operands, 16-bit reverse operands, fp32 accumulation),
pinned at BASELINE size and over many meta-steps (reference meta/train.py:119-129: the meta-gradient is what Adam sees).

(a) BASELINE-size slice: one 512-agent step (2 chunks, the bench's launch geometry) and one 508-agent step; their
    difference is the contribution of agents 508..511 computed INSIDE full-size launches (tile indices, strides and
    stream plan of the real workload).  It is compared with the fp64 autograd oracle on those four agents and the CUDA
    trajectories; pi_hat / y_hat of all K updates and the updated tables of two slices are compared as well.
(b) The stated tolerance is derived, not assumed: the oracle's own fp32-vs-fp64 drift on the same slice is measured and
    the tensor-core error is reported as a multiple of it.
(c) 20 meta-steps of training with the tensor-core path against the exact-fp32 path: the parameter displacement,
    lpg_loss and lpg_agent_return must agree to within the run-to-run variability of the fp32 path itself (two seeds)."""
import numpy as np
import pytest
import torch

from helpers import Case, to_oracle_traj, rel_err
from oracle import prng
from oracle.agents import AgentTables
from oracle.meta import lpg_meta_grad_train_step as o_step

pytestmark = pytest.mark.gpu

# stated tolerances of the production path (measured values are printed by the tests)
SLICE_GRAD_L2 = 1e-2        # relative L2 of the slice's meta-gradient contribution vs the fp64 oracle
SLICE_GRAD_BLOCK = 2e-2     # every parameter block, relative to the block's max |g|
PI_HAT_TOL, Y_HAT_TOL, TABLE_TOL = 3e-3, 3e-3, 5e-3


def _step(c, K, sl, n_global, mini_batches, snapshot=None):
    """Product step on the agents ``sl`` of case ``c`` with the keys / scaling of the global batch."""
    from to_ued_b200.meta.train import lpg_meta_grad_train_step, LPGTrainState, _WS_CACHE
    from to_ued_b200.models.lpg import LPG
    from to_ued_b200.models.optim import Adam
    from to_ued_b200.util.data import LpgHyperparams, TrainState
    import train as train_mod
    ag, ro = c.agent_state()
    vc = TrainState(Case.pad8(c.value), torch.zeros(c.n, dtype=torch.int32, device="cuda"), 1, 4e0, 0.5)
    if sl.stop - sl.start != c.n:
        assert sl.start == 0
        ag, vc = train_mod._shard(ag, vc, 0, sl.stop)
    ts = LPGTrainState(LPG(), torch.from_numpy(c.lpg).cuda(), Adam(1e-4))
    hy = LpgHyperparams(K, 0.5, 5e-2, 1e-3, 5e-3, 1e-3)
    out = lpg_meta_grad_train_step(prng.PRNGKey(21), ts, ag, vc, ro, mini_batches, 0.99, 0.95, hy, return_grad=True,
                                   global_num_agents=n_global, global_agent_offset=sl.start)
    torch.cuda.synchronize()
    return out, list(_WS_CACHE.values())


def _oracle_slice(c, sl, trajs, ev, K, dt):
    n = sl.stop - sl.start
    oag = AgentTables(torch.tensor(c.actor[sl]).to(dt), torch.tensor(c.critic[sl]).to(dt), torch.zeros(n, dtype=torch.long))
    p = c.p.index(sl)
    s0 = c.oro.batch_reset(None, p, c.w)
    return o_step(prng.PRNGKey(21), c.layout, torch.tensor(c.lpg).to(dt), oag, torch.tensor(c.value[sl]).to(dt), c.oro, p,
                  s0, c.life[sl], num_agent_updates=K, trajectories=trajs, eval_trajectory=ev, do_eval=False)


def _slice_traj(tape, k, a0, a1):
    from to_ued_b200.util.data import Transition
    return to_oracle_traj(Transition(tape.obs[k][a0:a1], tape.action[k][a0:a1], tape.reward[k][a0:a1], tape.done[k][a0:a1]))


def test_baseline_size_slice_against_fp64_oracle(built_lib):
    K, n, W, L = 5, 512, 64, 20
    c = Case("all_shortlife", n=n, seed=3)
    (_, ag_full, _, m_full), wss = _step(c, K, slice(0, n), n, 2)
    g_full = m_full["_grad"].double().cpu().numpy()
    # ---- forward quantities of two slices of the full-size run (chunk 0: agents 100..103, chunk 1: agents 508..511)
    nb = n // 2
    checks = []
    for a0 in (100, 508):
        ws = wss[a0 // nb]
        tape, l0 = ws.tape, a0 % nb
        trajs = [_slice_traj(tape, k, l0, l0 + 4) for k in range(K)]
        ev = _slice_traj(tape, K, l0, l0 + 4)
        o = _oracle_slice(c, slice(a0, a0 + 4), trajs, ev, K, torch.float64)
        rows = slice(l0 * W, (l0 + 4) * W)
        for k in range(K):
            # oracle debug tensors are [N, L, W]; the tape is time-major [L, R]
            pi_o = o["debug"][k]["pi_hat"].detach().numpy().transpose(1, 0, 2).reshape(L, 4 * W)
            y_o = o["debug"][k]["y_hat"].detach().numpy().transpose(1, 0, 2, 3).reshape(L, 4 * W, 8)
            e_pi = rel_err(tape.pi_hat[k][:, rows].cpu().numpy(), pi_o)
            e_y = rel_err(tape.y_hat[k][:, rows].cpu().numpy(), y_o)
            assert e_pi < PI_HAT_TOL and e_y < Y_HAT_TOL, f"agents {a0}.. update {k}: pi_hat {e_pi:.2e} y_hat {e_y:.2e}"
        e_t = rel_err(ag_full.actor_state.params[a0:a0 + 4, :, :5].cpu().numpy(), o["agents"].actor.detach().numpy())
        assert e_t < TABLE_TOL, f"agents {a0}..: updated actor tables {e_t:.2e}"
        checks.append((a0, e_pi, e_y, e_t))
        if a0 == 508:
            o64, trajs508, ev508 = o, trajs, ev
    print("forward slices (agents, pi_hat, y_hat of the last update, tables):", [(a, f"{x:.1e}", f"{y:.1e}", f"{z:.1e}") for a, x, y, z in checks])
    # ---- meta-gradient contribution of agents 508..511 inside full-size launches: g(512) - g(first 508)
    (_, _, _, m_508), _ = _step(c, K, slice(0, 508), n, 2)
    g_slice = g_full - m_508["_grad"].double().cpu().numpy()
    og = o64["grad"].numpy() * (4.0 / n)                       # the oracle returns the mean over ITS agents
    l2 = np.linalg.norm(g_slice - og) / np.linalg.norm(og)
    worst = 0.0
    for name, (off, cnt, shp) in c.layout.offsets.items():
        e = rel_err(g_slice[off:off + cnt], og[off:off + cnt])
        worst = max(worst, e)
        assert e < SLICE_GRAD_BLOCK, f"slice meta-gradient block {name}: {e:.3e}"
    # ---- (b) the oracle's own fp32-vs-fp64 drift on the same slice sets the scale
    o32 = _oracle_slice(c, slice(508, 512), trajs508, ev508, K, torch.float32)
    d32 = np.linalg.norm(o32["grad"].double().numpy() * (4.0 / n) - og) / np.linalg.norm(og)
    # summation-order noise of the subtraction itself: |g_full| * 1e-6 relative to the slice's norm
    noise = 2e-6 * np.linalg.norm(g_full) / np.linalg.norm(og)
    print(f"BASELINE-size slice (agents 508..511 of 512): meta-gradient rel L2 {l2:.2e}, worst block {worst:.2e}; "
          f"oracle fp32-vs-fp64 drift {d32:.2e} -> tensor-core error = {l2 / max(d32, 1e-12):.0f} x fp32 drift; "
          f"subtraction noise floor {noise:.1e}")
    assert l2 < SLICE_GRAD_L2
    assert d32 < 1e-4, "the fp32 oracle itself drifted more than expected"


def _train(steps, n_agents, precision, seed, monkeypatch):
    import to_ued_b200
    import train
    from to_ued_b200.experiments.parse_args import parse_args
    monkeypatch.setattr(to_ued_b200, "GRU_PRECISION", precision)
    monkeypatch.setattr(to_ued_b200, "CUDA_GRAPH", False)
    args = parse_args(["--env_mode", "all_shortlife", "--num_agents", str(n_agents), "--num_mini_batches", "1",
                       "--train_steps", str(steps), "--seed", str(seed)])
    from to_ued_b200.meta.meta import create_lpg_train_state
    from to_ued_b200.util import prng as P
    p0 = create_lpg_train_state(P.split(P.PRNGKey(seed), 3)[1], args).params.clone()
    hist, ts, _ = train.make_train(args)(P.PRNGKey(seed))
    torch.cuda.synchronize()
    loss = np.array([float(h["lpg_loss"]) for h in hist])
    ret = np.array([float(h["lpg_agent_return"]) for h in hist])
    return (ts.params - p0).double().cpu().numpy(), loss, ret


def test_twenty_meta_steps_tc_tracks_fp32_path(built_lib, monkeypatch):
    """(c) drift: the same 20 meta-steps (64 agents) on the tensor-core path and on the exact-fp32 path.  Sampled actions
    make training chaotic at the trajectory level, so the yardstick is the fp32 path's own sensitivity: the same run
    from a different seed."""
    steps, n = 20, 64
    d_tc, loss_tc, ret_tc = _train(steps, n, "tc", 0, monkeypatch)
    d_32, loss_32, ret_32 = _train(steps, n, "fp32", 0, monkeypatch)
    d_s1, loss_s1, ret_s1 = _train(steps, n, "fp32", 1, monkeypatch)
    cos = float(d_tc @ d_32 / (np.linalg.norm(d_tc) * np.linalg.norm(d_32)))
    rel = float(np.linalg.norm(d_tc - d_32) / np.linalg.norm(d_32))
    cos_seed = float(d_s1 @ d_32 / (np.linalg.norm(d_s1) * np.linalg.norm(d_32)))
    dl, dls = np.abs(loss_tc - loss_32).max(), np.abs(loss_s1 - loss_32).max()
    dr, drs = np.abs(ret_tc - ret_32).max(), np.abs(ret_s1 - ret_32).max()
    print(f"20 meta-steps, {n} agents: parameter displacement tc vs fp32: cosine {cos:.4f}, relative L2 {rel:.3f} "
          f"(fp32 seed 0 vs seed 1: cosine {cos_seed:.4f}); max |lpg_loss| diff {dl:.2e} (seeds: {dls:.2e}); "
          f"max |return| diff {dr:.3f} (seeds: {drs:.3f}); first-step loss diff {abs(loss_tc[0] - loss_32[0]):.2e}")
    assert abs(loss_tc[0] - loss_32[0]) < 1e-5                  # same rollouts in the first meta-step: loss is fp32-exact
    assert cos > 0.9 and cos > cos_seed - 0.02, "tensor-core training direction deviates from the fp32 path"
    assert dl <= max(2.0 * dls, 1e-3) and dr <= max(2.0 * drs, 0.05)
