"""GPU parity of the full LPG meta-gradient step (meta/train.py:14-130) against the autograd oracle.

The CUDA path is fp32 with hand-written reverse-mode kernels; the oracle differentiates the same
computation with torch.autograd in fp64 on the *same trajectories*.  Stated tolerance: every
parameter block of the meta-gradient within 2e-4 of the oracle, relative to the block's max |g|
(measured fp32 error is ~1e-6..2e-5; the oracle's own fp32-vs-fp64 drift is ~1e-6)."""
import numpy as np
import pytest
import torch

from oracle import prng
from oracle.agents import AgentTables, Hypers
from oracle.meta import lpg_meta_grad_train_step as o_step, Adam as OAdam
from helpers import Case, to_oracle_traj, rel_err

pytestmark = pytest.mark.gpu
GRAD_RTOL = 2e-4


def _run(c, K, quirk=True, mini_batches=1, n_global=None, offset=0):
    from to_ued_b200.meta.train import lpg_meta_grad_train_step, LPGTrainState, _WS_CACHE
    from to_ued_b200.models.lpg import LPG
    from to_ued_b200.models.optim import Adam
    from to_ued_b200.util.data import LpgHyperparams, TrainState
    ag, ro = c.agent_state()
    model = LPG(lifetime_conditioning=c.layout.lifetime_conditioning)
    lpg = torch.from_numpy(c.lpg).cuda()
    ts = LPGTrainState(model, lpg, Adam(1e-4))
    vc = TrainState(Case.pad8(c.value), torch.zeros(c.n, dtype=torch.int32, device="cuda"), 1, 4e0, 0.5)
    hy = LpgHyperparams(K, 0.5, 5e-2, 1e-3, 5e-3, 1e-3)
    out = lpg_meta_grad_train_step(prng.PRNGKey(21), ts, ag, vc, ro, mini_batches, 0.99, 0.95, hy,
                                   outer_product_quirk=quirk, return_grad=True,
                                   global_num_agents=n_global, global_agent_offset=offset)
    torch.cuda.synchronize()
    return out, list(_WS_CACHE.values())[0]


@pytest.mark.parametrize("mode,cond,quirk", [("all_shortlife", False, True), ("all_vrandlife", True, True),
                                             ("all_shortlife", False, False)])
def test_meta_gradient_matches_autograd_oracle(built_lib, fp32_gru, mode, cond, quirk):
    K, n = 5, 4
    c = Case(mode, n=n, seed=7, cond=cond, table_scale=0.3, lifetimes=[250, 3, 250, 250], steps=[0, 0, 17, 246])
    (new_ts, ag2, vc2, met), ws = _run(c, K, quirk)
    tape = ws.tape
    trajs = [to_oracle_traj(tape.transition(k)) for k in range(K)]
    ev = to_oracle_traj(tape.transition(K))
    dt = torch.float64
    oag = AgentTables(torch.tensor(c.actor).to(dt), torch.tensor(c.critic).to(dt), torch.tensor(c.steps.astype(np.int64)))
    s0 = c.oro.batch_reset(None, c.p, c.w)
    o = o_step(prng.PRNGKey(21), c.layout, torch.tensor(c.lpg).to(dt), oag, torch.tensor(c.value).to(dt), c.oro, c.p,
               s0, c.life, num_agent_updates=K, trajectories=trajs, eval_trajectory=ev, outer_product_quirk=quirk,
               do_eval=False)
    g = met["_grad"].cpu().numpy().astype(np.float64)
    og = o["grad"].numpy()
    worst = 0.0
    for name, (off, cnt, shp) in c.layout.offsets.items():
        e = rel_err(g[off:off + cnt], og[off:off + cnt])
        worst = max(worst, e)
        assert e < GRAD_RTOL, f"meta-gradient block {name}: rel err {e:.3e}"
    print(f"[{mode} cond={cond} quirk={quirk}] worst block rel err {worst:.2e}; |g|={np.linalg.norm(og):.3e}")
    # metrics
    om = o["metrics"]
    for k_ in ("lpg_loss", "reg_lpg_loss", "value_loss"):
        np.testing.assert_allclose(float(met[k_]), float(om[k_]), rtol=5e-4, atol=1e-6, err_msg=k_)
    for k_, v in om["lpg_agent"].items():
        np.testing.assert_allclose(float(met["lpg_agent"][k_]), float(v), rtol=5e-4, atol=1e-7, err_msg=k_)
    # Adam step on the LPG parameters (models/optim.py:12-17)
    oadam = OAdam(c.layout.size, 1e-4, dtype=dt)
    want = oadam.step(torch.tensor(c.lpg).to(dt), torch.tensor(g))
    np.testing.assert_allclose(new_ts.params.cpu().numpy(), want.numpy(), rtol=0, atol=2e-7)
    # agents after the step; value critic untouched except its step counter (Q2)
    np.testing.assert_array_equal(ag2.actor_state.step.cpu().numpy(), o["agents"].step.numpy())
    assert rel_err(ag2.actor_state.params[..., :5].cpu().numpy(), o["agents"].actor.numpy()) < 5e-5
    assert torch.equal(vc2.params, Case.pad8(c.value)) and int(vc2.step[0]) == K + 1


def test_eval_rollout_and_return_metric_bit_exact(built_lib, fp32_gru):
    """The eval rollout (meta/train.py:46-58) and eval_agent's 4-worker return (Q11) re-sampled by the
    oracle from the CUDA tables."""
    from oracle.agents import eval_agent as o_eval
    K, n = 2, 3
    c = Case("all_shortlife", n=n, seed=9, table_scale=0.3)
    (new_ts, ag2, vc2, met), ws = _run(c, K)
    rngs = prng.split(prng.PRNGKey(21), n)
    ks = prng.split(rngs, 2); rngs = ks[:, 0, :]
    ks = prng.split(rngs, 2); rngs = ks[:, 0, :]
    ks = prng.split(rngs, 2)
    table = ws.tape.actor[K][..., :5].cpu().numpy()
    ret = o_eval(ks[:, 1, :], c.oro, c.p, table, 4)
    np.testing.assert_allclose(float(met["lpg_agent_return"]), float(ret.mean()), rtol=1e-6)


def test_mini_batches_and_sharding_are_exact_partitions(built_lib, fp32_gru):
    """num_mini_batches (util/jax.py:25-41) and the N-way agent partition used for data parallelism
    must not change the result: chunked == full, and sum of per-shard gradients == full."""
    K, n = 2, 4
    c = Case("all_shortlife", n=n, seed=5, table_scale=0.3)
    (_, _, _, m1), _ = _run(c, K, mini_batches=1)
    g1 = m1["_grad"].clone()
    (_, _, _, m2), _ = _run(c, K, mini_batches=2)
    assert rel_err(m2["_grad"].cpu().numpy(), g1.cpu().numpy()) < 2e-6
    # two "ranks" of 2 agents each, emulated sequentially: gradients are pre-scaled by 1/n_global
    gs = []
    for r in range(2):
        cr = Case("all_shortlife", n=n, seed=5, table_scale=0.3)
        for f in ("actor", "critic", "value", "life", "steps", "keys"):
            setattr(cr, f, getattr(cr, f)[2 * r:2 * r + 2])
        cr.p = cr.p.index(slice(2 * r, 2 * r + 2)); cr.n = 2
        (_, _, _, mr), _ = _run(cr, K, n_global=4, offset=2 * r)
        gs.append(mr["_grad"].clone())
    assert rel_err((gs[0] + gs[1]).cpu().numpy(), g1.cpu().numpy()) < 2e-6


def test_side_stream_plan_equals_serial_enqueue(built_lib, monkeypatch):
    """The production stream plan (token sort, the dense theta_k -> theta_{k+1} table copy, the agent adjoint and the
    embedding gradient on side streams; ``tables_precopied``) against everything enqueued on one stream with the copy
    inside toued_agent_update: bit-identical gradient, tables, step counters and metrics."""
    import to_ued_b200
    K, n = 5, 6
    res = {}
    for side in (True, False):
        monkeypatch.setattr(to_ued_b200, "SIDE_STREAMS", side)
        c = Case("all_shortlife", n=n, seed=11, table_scale=0.3, lifetimes=[250, 3, 250, 250, 1, 250], steps=[0, 0, 17, 246, 0, 5])
        (new_ts, ag2, vc2, met), ws = _run(c, K)
        res[side] = (met["_grad"].cpu().numpy(), ag2.actor_state.params.cpu().numpy(), ag2.critic_state.params.cpu().numpy(),
                     ag2.actor_state.step.cpu().numpy(), float(met["lpg_loss"]), float(met["lpg_agent"]["policy_entropy"]))
    for a, b in zip(res[True], res[False]):
        np.testing.assert_array_equal(a, b)
