"""ES / TA-LPG step (meta/train.py:133-227, evosax OpenES restated): CUDA vs oracle."""
import numpy as np
import pytest
import torch

from oracle import prng
from oracle import es as OES
from oracle.agents import AgentTables
from helpers import Case, to_oracle_traj, rel_err

pytestmark = pytest.mark.gpu


def test_es_ask_tell_match_oracle(built_lib):
    from to_ued_b200 import _lib
    P, pop = 1000, 8
    rs = np.random.RandomState(0)
    mean = torch.tensor(rs.randn(P).astype(np.float32)).cuda()
    key = prng.PRNGKey(5)
    kd = torch.from_numpy(key.view(np.int32)).cuda()
    cand = torch.empty((pop, P), device="cuda")
    _lib.call("toued_es_ask", _lib.ptr(kd), _lib.ptr(mean), 0.1, _lib.ptr(cand), pop, P, P, _lib.stream_ptr())
    st = OES.ESState(mean.cpu().double(), 0.1, torch.zeros(P, dtype=torch.float64), torch.zeros(P, dtype=torch.float64), 1e-2, 3)
    x = OES.es_ask(key, st, pop)
    idx = np.stack([np.arange(pop // 2), np.arange(pop // 2) + pop // 2], 1).reshape(-1)
    np.testing.assert_allclose(cand.cpu().numpy(), x[idx].numpy(), rtol=0, atol=2e-6)       # erfinv rounding
    fit = torch.tensor([1., 0., 0., 1., 1., 0., 0., 1.])
    m = torch.tensor(rs.randn(P).astype(np.float32) * 1e-2).cuda(); v = torch.tensor(rs.rand(P).astype(np.float32) * 1e-3).cuda()
    st.m, st.v = m.cpu().double(), v.cpu().double()
    want = OES.es_tell(cand.cpu().double(), fit, st, pop)
    mean2 = mean.clone()
    _lib.call("toued_es_tell", _lib.ptr(cand), _lib.ptr(fit.cuda()), _lib.ptr(mean2), _lib.ptr(m), _lib.ptr(v), pop, P, P,
              0.1, 1e-2, 0.99, 0.999, 1e-8, 3, 0.0, _lib.stream_ptr())
    np.testing.assert_allclose(mean2.cpu().numpy(), want.mean.numpy(), rtol=0, atol=5e-6)
    np.testing.assert_allclose(m.cpu().numpy(), want.m.numpy(), rtol=1e-4, atol=1e-7)


def test_lpg_es_train_step_matches_oracle(built_lib, fp32_gru):
    from to_ued_b200.meta.es import lpg_es_train_step, create_es_train_state
    from to_ued_b200.meta.train import LPGTrainState
    from to_ued_b200.models.lpg import LPG
    from to_ued_b200.models.optim import Adam
    from to_ued_b200.util.data import LpgHyperparams
    from to_ued_b200.experiments.parse_args import parse_args
    n, K = 3, 3
    c = Case("all_vrandlife", n=n, seed=8, cond=True, table_scale=0.3, lifetimes=[250, 2, 250])
    ag, ro = c.agent_state()
    args = parse_args(["--env_mode", "all_vrandlife", "--num_agents", str(n), "--num_mini_batches", "1", "--use_es",
                       "--lifetime_conditioning", "--lpg_learning_rate", "0.01"])
    ts = LPGTrainState(LPG(lifetime_conditioning=True), torch.from_numpy(c.lpg).cuda(), Adam(1e-2))
    es = create_es_train_state(prng.PRNGKey(0), args, ts)
    es.es_state["mean"] = torch.from_numpy(c.lpg).cuda()          # a non-degenerate mean (evosax starts from zeros)
    hy = LpgHyperparams(K, 0.5, 5e-2, 1e-3, 5e-3, 1e-3)
    rng = prng.PRNGKey(31)
    es2, ag2, _, met = lpg_es_train_step(rng, es, ag, None, ro, 1, hy)
    torch.cuda.synchronize()
    cand = met["_candidates"].cpu()
    fit = met["_fitness"].cpu().numpy()
    # ---- oracle with the CUDA candidates (erfinv rounding aside they are the oracle's own) ----
    dt = torch.float64
    ost = OES.ESState(torch.tensor(c.lpg).to(dt), 0.1, torch.zeros(c.layout.size, dtype=dt), torch.zeros(c.layout.size, dtype=dt), 1e-2)
    x = OES.es_ask(prng.split(rng, 2)[1], ost, 2 * n)
    idx = np.stack([np.arange(n), np.arange(n) + n], 1).reshape(-1)
    np.testing.assert_allclose(cand.numpy(), x[idx].numpy(), rtol=0, atol=2e-6)
    oag = AgentTables(torch.tensor(c.actor).to(dt), torch.tensor(c.critic).to(dt), torch.tensor(c.steps.astype(np.int64)))
    s0 = c.oro.batch_reset(None, c.p, c.w)
    o = OES.lpg_es_train_step(rng, c.layout, ost, oag, c.oro, c.p, s0, c.life, num_agent_updates=K,
                              candidates=cand.to(dt))
    # fitness comes from bit-exact rollouts of tables that agree to ~1e-6: compare with tolerance, ranks exactly
    np.testing.assert_allclose(fit, o["fitness"], rtol=0, atol=0.08)
    assert rel_err(ag2.actor_state.params[..., :5].cpu().numpy(), o["agents"].actor.numpy()) < 5e-4 or \
        not np.array_equal(fit[::2] > fit[1::2], o["first_greater"])
    want = OES.es_tell(cand.to(dt), torch.tensor((np.repeat(fit[::2] > fit[1::2], 2) ^ np.tile([False, True], n)).astype(np.float32)), ost, 2 * n)
    # first Adam step is sign-like (lr * g / (|g| + eps)): an element whose gradient is ~0 is ill-conditioned
    diff = np.abs(es2.es_state["mean"].cpu().numpy() - want.mean.numpy())
    assert (diff > 5e-6).mean() < 1e-4 and diff.max() < 2e-2
    assert es2.es_state["gen_counter"] == 1 and abs(es2.es_state["lrate"] - 1e-2 * 0.999) < 1e-12


def test_es_per_candidate_forward_on_tensor_cores(built_lib, monkeypatch):
    """TOUED_ES_PRECISION=tc: the per-candidate LPG forward on tcgen05 (one parameter set and one set of pass
    images per CTA) against the exact-fp32 per-candidate kernel.  Stated tolerance: agent tables after K updates
    within 5e-3 (relative to max |value|), candidate fitness within 0.15 (fitness comes from sampled rollouts of
    the trained tables, so it is compared loosely)."""
    import to_ued_b200
    from to_ued_b200.agents.lpg_agent import train_lpg_agent
    n, K = 6, 3
    c = Case("all_vrandlife", n=n, seed=11, cond=True, table_scale=0.4)
    rs = np.random.RandomState(5)
    P = c.lpg.size
    Pp = (P + 3) // 4 * 4
    cand = np.zeros((n, Pp), np.float32)
    cand[:, :P] = c.lpg[None] + 0.05 * rs.randn(n, P).astype(np.float32)
    cand_d = torch.from_numpy(cand).cuda()

    class _Cand:
        params = cand_d
        model = type("M", (), {"lifetime_conditioning": True})()
    out = {}
    for prec in ("fp32", "tc"):
        monkeypatch.setattr(to_ued_b200, "ES_PRECISION", prec)
        ag, ro = c.agent_state()
        ag2, _, m = train_lpg_agent(c.keys, _Cand, ag, ro, K, 0.5, lpg_stride=Pp)
        torch.cuda.synchronize()
        out[prec] = (ag2.actor_state.params.cpu().numpy(), ag2.critic_state.params.cpu().numpy(), m.policy_entropy.cpu().numpy())
    for i, name in enumerate(("actor", "critic", "policy_entropy")):
        e = rel_err(out["tc"][i], out["fp32"][i])
        print(f"ES tc vs fp32 {name}: rel err {e:.2e}")
        assert e < 5e-3, (name, e)
