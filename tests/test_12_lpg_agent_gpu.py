"""GPU parity of the LPG forward pass and the LPG-driven agent update against the oracle
(agents/lpg_agent.py:31-140, models/lpg.py:11-85).  Float tolerance: the CUDA path is fp32; the
oracle is evaluated in fp64 on the *same trajectories* (a 1-ulp table difference must not change
the sampled data), and separately each CUDA rollout is re-sampled by the oracle from the CUDA
tables and must match bit for bit."""
import numpy as np
import pytest
import torch

from oracle import prng
from oracle.agents import AgentTables, Hypers, train_lpg_agent as o_train
from helpers import Case, to_oracle_traj, rel_err

pytestmark = pytest.mark.gpu

RTOL_TABLE = 2e-5     # updated tables (values O(1..40) after lr=40 steps), relative to max |entry|
RTOL_LPG = 2e-4       # pi_hat / y_hat after a 20-step fp32 GRU


@pytest.mark.parametrize("mode,cond,K", [("all_shortlife", False, 3), ("all_vrandlife", True, 3), ("mazes", False, 2)])
def test_train_lpg_agent_matches_oracle(built_lib, fp32_gru, mode, cond, K):
    from to_ued_b200.agents.lpg_agent import train_lpg_agent, Tape
    n = 5
    # agent 1 reaches its lifetime during the K updates (mask path), agent 2 starts beyond it
    c = Case(mode, n=n, seed=3, cond=cond, table_scale=0.5,
             lifetimes=[250, 2, 7, 250, 250], steps=[0, 0, 9, 40, 249])
    ag, ro = c.agent_state()
    lpg = torch.from_numpy(c.lpg).cuda()

    class LS:                       # minimal lpg_train_state
        params = lpg
        class model: lifetime_conditioning = cond
    tape = Tape(n, c.w, c.L, c.D, K, "cuda", precision="fp32")
    rng = prng.split(prng.PRNGKey(11), n)
    ag2, rollouts, met = train_lpg_agent(rng, LS, ag, ro, K, 0.5, tape=tape)
    torch.cuda.synchronize()

    # ---- oracle on the same trajectories, fp64 ----
    trajs = [to_oracle_traj(r) for r in rollouts]
    dt = torch.float64
    oag = AgentTables(torch.tensor(c.actor).to(dt).requires_grad_(True), torch.tensor(c.critic).to(dt).requires_grad_(True),
                      torch.tensor(c.steps.astype(np.int64)))
    s0 = c.oro.batch_reset(None, c.p, c.w)
    oagK, _, _, omet, dbg = o_train(rng, c.layout, torch.tensor(c.lpg).to(dt), oag, c.oro, c.p, s0, c.life, K, 0.5,
                                    Hypers(), trajectories=trajs)
    for k in range(K):
        seq = lambda a: a.detach().numpy()            # oracle debug tensors are [N, L, W(, 8)]
        ph = tape.pi_hat[k].view(c.L, n, c.w).permute(1, 0, 2).cpu().numpy()
        yh = tape.y_hat[k].view(c.L, n, c.w, 8).permute(1, 0, 2, 3).cpu().numpy()
        assert rel_err(ph, seq(dbg[k]["pi_hat"])) < RTOL_LPG, f"pi_hat k={k}"
        assert rel_err(yh, seq(dbg[k]["y_hat"])) < RTOL_LPG, f"y_hat k={k}"
        assert rel_err(tape.actor[k + 1][..., :5].cpu().numpy(), seq(dbg[k]["actor"])) < RTOL_TABLE, f"actor k={k}"
        assert rel_err(tape.critic[k + 1].cpu().numpy(), seq(dbg[k]["critic"])) < RTOL_TABLE, f"critic k={k}"
        gn = tape.scalars[k, :, 0].cpu().numpy()
        np.testing.assert_allclose(gn, dbg[k]["ga"].detach().flatten(1).norm(dim=1).numpy(), rtol=1e-4)
        np.testing.assert_array_equal(tape.step_in[k].cpu().numpy() + tape.scalars[k, :, 2].cpu().numpy().astype(np.int32),
                                      dbg[k]["step"].numpy())
    np.testing.assert_array_equal(ag2.actor_state.step.cpu().numpy(), oagK.step.numpy())
    for name in ("policy_l2", "policy_entropy", "critic_loss", "critic_l2", "critic_entropy"):
        np.testing.assert_allclose(getattr(met, name).cpu().numpy(), omet[name].detach().numpy(), rtol=2e-4, atol=1e-7,
                                   err_msg=name)

    # ---- every CUDA rollout re-sampled by the oracle from the CUDA tables: bit-exact ----
    r = rng
    st = s0
    for k in range(K):
        ks = prng.split(r, 2); r, rk = ks[:, 0, :], ks[:, 1, :]
        otraj, st, _ = c.oro.batch_rollout(rk, tape.actor[k][..., :5].cpu().numpy(), c.p, st)
        np.testing.assert_array_equal(otraj.action, trajs[k].action)
        np.testing.assert_array_equal(otraj.obs_idx, trajs[k].obs_idx)
        np.testing.assert_array_equal(otraj.reward, trajs[k].reward)
    np.testing.assert_array_equal(ag2.env_state.pos.cpu().numpy(), st.pos)


def test_sort_tokens_is_stable_row_sort(built_lib):
    from to_ued_b200 import _lib
    c = Case("all_shortlife", n=4, seed=1)
    ag, ro = c.agent_state()
    traj, _, _, _ = ro.batch_rollout(c.keys, ag.actor_state, ag.level.packed, ag.env_obs, ag.env_state)
    T = c.w * c.L
    out = torch.empty((4, T), dtype=torch.int16, device="cuda")
    _lib.call("toued_sort_tokens", _lib.ptr(traj.obs), _lib.ptr(out), 4, c.w, c.L, _lib.stream_ptr())
    rows = (traj.obs[:, :c.L].reshape(4, T) & 0xFFFF).cpu().numpy()
    want = np.argsort(rows, axis=1, kind="stable")
    np.testing.assert_array_equal(out.cpu().numpy().astype(np.int64), want)


@pytest.mark.parametrize("w,L", [(8, 5), (64, 6), (64, 8), (64, 16), (64, 20), (128, 25), (128, 32)])
def test_sort_tokens_all_sizes(built_lib, w, L):
    """every variant of the sort (shared-memory network below 512 keys; 2 / 4 / 8 / 16 keys per thread above) against
    numpy's stable argsort, on random rows with many repeats"""
    from to_ued_b200 import _lib
    g = np.random.default_rng(w * 100 + L)
    n, T = 5, w * L
    rows = g.integers(0, 40 if T < 1000 else 3201, size=(n, L + 1, w)).astype(np.int32)
    times = g.integers(0, 50, size=(n, L + 1, w)).astype(np.int32)
    obs = torch.from_numpy(rows | (times << 16)).cuda()
    out = torch.full((n, T), -1, dtype=torch.int16, device="cuda")
    _lib.call("toued_sort_tokens", _lib.ptr(obs), _lib.ptr(out), n, w, L, _lib.stream_ptr())
    want = np.argsort(rows[:, :L].reshape(n, T), axis=1, kind="stable")
    np.testing.assert_array_equal(out.cpu().numpy().astype(np.int64), want)
