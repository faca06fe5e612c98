"""tcgen05 / TMEM building blocks on real hardware: a 128x48x256 fp16 GEMM tile vs an exact product of
the fp16-rounded operands (the only rounding is fp32 accumulation order)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_tcgen05_gemm_tile(built_lib):
    from to_ued_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(0)
    A = torch.randn(128, 256, generator=g)
    B = torch.randn(48, 256, generator=g)
    Ad, Bd = A.cuda(), B.cuda()
    img = torch.zeros(48 * 256, dtype=torch.float16, device="cuda")
    D = torch.zeros(128, 48, device="cuda")
    _lib.call("toued_tc_gemm_test", _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(img), _lib.ptr(D), _lib.stream_ptr())
    torch.cuda.synchronize()
    want = A.half().double() @ B.half().double().T
    err = (D.cpu().double() - want).abs().max().item()
    assert err < 2e-4, f"tcgen05 tile mismatch: max abs err {err}"


@pytest.mark.parametrize("cond", [False, True])
def test_gru_forward_tc_matches_fp32_kernel_and_oracle(built_lib, cond):
    """Tensor-core GRU forward vs the exact-fp32 SIMT kernel on identical inputs.  Stated tolerance:
    pi_hat / y_hat within 3e-3 of the fp32 kernel relative to max |value| (fp16 operands, hidden state
    quantised to fp16 once per step, 20 recurrent steps)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import Case, rel_err
    from oracle import prng
    from to_ued_b200 import _lib
    from to_ued_b200.agents.lpg_agent import Tape
    n, K = 6, 1
    c = Case("all_shortlife", n=n, seed=4, cond=cond, table_scale=0.5)
    ag, ro = c.agent_state()
    lpg = torch.from_numpy(c.lpg).cuda()
    p, s = _lib.ptr, _lib.stream_ptr()
    out = {}
    traj, _, _, _ = ro.batch_rollout(c.keys, ag.actor_state, ag.level.packed, ag.env_obs, ag.env_state)
    for prec in ("fp32", "tc"):
        tape = Tape(n, c.w, c.L, c.D, K, "cuda", precision=prec)
        tape.obs[0].copy_(traj.obs); tape.action[0].copy_(traj.action)
        tape.reward[0].copy_(traj.reward); tape.done[0].copy_(traj.done)
        _lib.call("toued_lpg_prepare", p(tape.obs[0]), p(tape.action[0]), p(tape.reward[0]), p(tape.done[0]),
                  p(ag.actor_state.params), p(ag.critic_state.params), p(lpg), p(ag.actor_state.step),
                  p(ag.level.packed), p(tape.x[0]), n, c.w, c.L, c.D, int(cond), s)
        if prec == "tc":
            _lib.call("toued_pack_wh_forward", p(lpg), p(tape.wh_img), s)
            _lib.call("toued_gru_forward_tc", p(tape.x[0]), p(tape.done[0]), p(lpg), p(tape.wh_img), p(tape.h16[0]),
                      p(tape.g16[0]), p(tape.pi_hat[0]), p(tape.y_hat[0]), n, c.w, c.L, int(cond), s)
            out[prec] = (tape.pi_hat[0].clone(), tape.y_hat[0].clone(), tape.h16[0].float(), tape.g16[0].float())
        else:
            _lib.call("toued_gru_forward", p(tape.x[0]), p(tape.done[0]), p(lpg), p(tape.h[0]), p(tape.gates[0]),
                      p(tape.pi_hat[0]), p(tape.y_hat[0]), n, c.w, c.L, int(cond), s)
            out[prec] = (tape.pi_hat[0].clone(), tape.y_hat[0].clone(), tape.h[0].clone(), tape.gates[0].clone())
    torch.cuda.synchronize()
    names = ("pi_hat", "y_hat", "h", "gates")
    for nm, a, b in zip(names, out["tc"], out["fp32"]):
        e = rel_err(a.cpu().numpy(), b.cpu().numpy())
        print(f"tc vs fp32 {nm}: rel err {e:.2e}")
        assert e < 3e-3, f"{nm}: {e}"
