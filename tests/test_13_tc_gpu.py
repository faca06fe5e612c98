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


def test_tcgen05_mixed_swizzle_k_blocks(built_lib):
    """K = 256 in SW128 blocks + 16 in a no-swizzle K-major block within one accumulation chain."""
    from to_ued_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(2)
    A = torch.randn(128, 272, generator=g)
    B = torch.randn(48, 272, generator=g)
    Ad, Bd = A.cuda(), B.cuda()
    img = torch.zeros(32768, dtype=torch.uint8, device="cuda")
    D = torch.zeros(128, 48, device="cuda")
    _lib.call("toued_tc_gemm_mixed_test", _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(img), _lib.ptr(D), _lib.stream_ptr())
    torch.cuda.synchronize()
    want = A.half().double() @ B.half().double().T
    err = (D.cpu().double() - want).abs().max().item()
    assert err < 2e-4, f"mixed-swizzle tcgen05 tile mismatch: max abs err {err}"


def test_tcgen05_mn_major_gemm_tile(built_lib):
    from to_ued_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randn(128, 128, generator=g)          # [k][m]
    B = torch.randn(128, 128, generator=g)          # [k][n]
    img = torch.zeros(65536, dtype=torch.uint8, device="cuda")
    D = torch.zeros(128, 128, device="cuda")
    Ad, Bd = A.cuda(), B.cuda()                      # keep alive across the asynchronous launch
    # MN-major SW128 descriptor: LBO = 8192 B between 64-element MN groups, SBO = 1024 B between 8-row
    # k groups, +2048 B per K=16 step
    _lib.call("toued_tc_gemm_mn_test", _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(img), _lib.ptr(D), 8192, 1024, 2048,
              _lib.stream_ptr())
    torch.cuda.synchronize()
    want = A.bfloat16().double().T @ B.bfloat16().double()
    err = (D.cpu().double() - want).abs().max().item()
    assert err < 2e-4, f"MN-major tcgen05 tile mismatch: max abs err {err}"


def test_tcgen05_cta_pair_mn_major_gemm_tile(built_lib):
    """cta_group::2: a 2-CTA cluster shares M256 N256 K16 MMAs (each CTA holds its 128 rows of A and half of B's N)."""
    from to_ued_b200 import _lib
    g = torch.Generator(device="cpu").manual_seed(2)
    A = torch.randn(128, 256, generator=g)          # [k][m]
    B = torch.randn(128, 256, generator=g)          # [k][n]
    img = torch.zeros(131072, dtype=torch.uint8, device="cuda")
    D = torch.zeros(256, 256, device="cuda")
    Ad, Bd = A.cuda(), B.cuda()
    _lib.call("toued_tc_gemm_mn2_test", _lib.ptr(Ad), _lib.ptr(Bd), _lib.ptr(img), _lib.ptr(D), _lib.stream_ptr())
    torch.cuda.synchronize()
    want = A.bfloat16().double().T @ B.bfloat16().double()
    err = (D.cpu().double() - want).abs().max().item()
    assert err < 4e-4, f"CTA-pair MN-major tcgen05 tile mismatch: max abs err {err}"


@pytest.mark.parametrize("cond,n", [(False, 6), (True, 6), (False, 5), (True, 3)])
def test_gru_forward_tc_matches_fp32_kernel_and_oracle(built_lib, cond, n):
    """Tensor-core GRU forward vs the exact-fp32 SIMT kernel on identical inputs.  Stated tolerance:
    pi_hat / y_hat within 3e-3 of the fp32 kernel relative to max |value| (fp16 operands, hidden state
    quantised to fp16 once per step, 20 recurrent steps)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import Case, rel_err
    from oracle import prng
    from to_ued_b200 import _lib
    from to_ued_b200.agents.lpg_agent import Tape
    K = 1           # n = 5 / 3: 320 / 192 sequences, i.e. a ragged last 128-row tile (64 valid rows)
    c = Case("all_shortlife", n=n, seed=4, cond=cond, table_scale=0.5)
    ag, ro = c.agent_state()
    lpg = torch.from_numpy(c.lpg).cuda()
    p, s = _lib.ptr, _lib.stream_ptr()
    out = {}
    global tape_tc_hpimg
    tape_tc_hpimg = lambda o: o["hpimg"]
    traj, _, _, _ = ro.batch_rollout(c.keys, ag.actor_state, ag.level.packed, ag.env_obs, ag.env_state)
    for prec in ("fp32", "tc"):
        tape = Tape(n, c.w, c.L, c.D, K, "cuda", precision=prec)
        tape.obs[0].copy_(traj.obs); tape.action[0].copy_(traj.action)
        tape.reward[0].copy_(traj.reward); tape.done[0].copy_(traj.done)
        _lib.call("toued_lpg_prepare", p(tape.obs[0]), p(tape.action[0]), p(tape.reward[0]), p(tape.done[0]),
                  p(ag.actor_state.params), p(ag.critic_state.params), p(lpg), p(ag.actor_state.step),
                  p(ag.level.packed), p(tape.x[0]), None, n, c.w, c.L, c.D, int(cond), 0, s)
        if prec == "tc":
            _lib.call("toued_pack_wh_forward", p(lpg), p(tape.wh_img), int(cond), s)
            _lib.call("toued_gru_forward_tc", p(tape.x[0]), p(tape.done[0]), p(lpg), p(tape.wh_img), p(tape.h16[0]),
                      p(tape.fac[0]), p(tape.hpimg[0]), p(tape.pi_hat[0]), p(tape.y_hat[0]), n, c.w, c.L, int(cond), s)
            R = n * c.w
            planes = [_fac_plane(tape.fac[0], i, c.L, R) for i in range(4)]
            # the sign bits of the saved z plane carry relu'(h_t) for the reverse pass: set <=> fp16 h_t <= 0
            h16_ = _from_rb32(tape.h16[0], c.L, R)
            assert torch.equal(torch.signbit(planes[1]), h16_ <= 0), "z-plane sign bits != (h16 <= 0)"
            planes[1] = planes[1].abs()
            out[prec] = (tape.pi_hat[0].clone(), tape.y_hat[0].clone(), h16_.float(), torch.stack(planes).float())
            out["hpimg"] = tape.hpimg[0].clone()
        else:
            _lib.call("toued_gru_forward", p(tape.x[0]), p(tape.done[0]), p(lpg), p(tape.h[0]), p(tape.gates[0]),
                      p(tape.pi_hat[0]), p(tape.y_hat[0]), n, c.w, c.L, int(cond), 0, s)
            r_, z_, n_, hn_ = tape.gates[0]
            h_ = tape.h[0]
            nd = (1 - tape.done[0].float()).permute(1, 0, 2).reshape(c.L, n * c.w, 1)     # [L][R][1]
            hp_ = torch.cat([h_[1:], torch.zeros_like(h_[:1])], 0) * nd                  # masked carry
            fac_ref = torch.stack([r_, z_, n_, hn_])                 # the saved gate planes
            out[prec] = (tape.pi_hat[0].clone(), tape.y_hat[0].clone(), h_.clone(), fac_ref)
            out["hp_ref"] = hp_
    torch.cuda.synchronize()
    # masked carry image vs reference
    from_img = _unpack_tile_img(tape_tc_hpimg(out), c.L * n * c.w, 256)
    e = rel_err(from_img.float().cpu().numpy(), out["hp_ref"].reshape(-1, 256).cpu().numpy())
    print(f"hp image: rel err {e:.2e}")
    assert e < 5e-3
    names = ("pi_hat", "y_hat", "h", "gates")
    for nm, a, b in zip(names, out["tc"], out["fp32"]):
        e = rel_err(a.cpu().numpy(), b.cpu().numpy())
        print(f"tc vs fp32 {nm}: rel err {e:.2e}")
        assert e < 3e-3, f"{nm}: {e}"


def _unpack_tile_img(img_u8, n_tok, C):
    """token tile image (uint8 tensor) -> fp16 [n_tok][C] (inverse of tile_img_offset in csrc/tc.cuh)."""
    tok = torch.arange(n_tok, device=img_u8.device)
    col = torch.arange(C, device=img_u8.device)
    tb, r = tok // 64, tok % 64
    cg, cin = col // 64, col % 64
    off = ((tb[:, None] * (C // 64) + cg[None, :]) << 13) + r[:, None] * 128 + (((cin[None, :] >> 3) ^ (r[:, None] & 7)) << 4) + ((cin[None, :] & 7) << 1)
    flat = img_u8.view(torch.float16)
    return flat[(off // 2).reshape(-1)].reshape(n_tok, C)


@pytest.mark.parametrize("mode,cond,n,w", [("all_shortlife", False, 4, 64), ("all_vrandlife", True, 4, 64),
                                           ("all_shortlife", False, 3, 64), ("all_shortlife", False, 3, 8)])
def test_meta_gradient_tensor_core_path(built_lib, mode, cond, n, w, monkeypatch):
    """Full LPG meta-gradient with the tensor-core GRU (fp16 forward, bf16 reverse operands, fp32
    accumulation in TMEM) against the fp64 autograd oracle on the same trajectories.
    Stated tolerance: every parameter block within 2e-2 of the oracle relative to the block's max |g|,
    and the whole gradient within 1e-2 in relative L2 norm."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import to_ued_b200
    monkeypatch.setattr(to_ued_b200, "GRU_PRECISION", "tc")
    import numpy as np
    from helpers import Case, to_oracle_traj, rel_err
    from oracle import prng
    from oracle.agents import AgentTables
    from oracle.meta import lpg_meta_grad_train_step as o_step
    from test_14_meta_grad_gpu import _run
    K = 5           # n = 3: ragged last tile through the whole tensor-core chain (forward, BPTT, weight gradients);
                    # w = 8: 24 sequences = one partial tile, not a multiple of 32 (streaming head-gradient kernel)
    c = Case(mode, n=n, w=w, seed=7, cond=cond, table_scale=0.3, lifetimes=[250, 3, 250, 250][:n], steps=[0, 0, 17, 246][:n])
    (new_ts, ag2, vc2, met), ws = _run(c, K)
    assert ws.tape.precision == "tc"
    tape = ws.tape
    trajs = [to_oracle_traj(tape.transition(k)) for k in range(K)]
    ev = to_oracle_traj(tape.transition(K))
    dt = torch.float64
    oag = AgentTables(torch.tensor(c.actor).to(dt), torch.tensor(c.critic).to(dt), torch.tensor(c.steps.astype(np.int64)))
    s0 = c.oro.batch_reset(None, c.p, c.w)
    o = o_step(prng.PRNGKey(21), c.layout, torch.tensor(c.lpg).to(dt), oag, torch.tensor(c.value).to(dt), c.oro, c.p,
               s0, c.life, num_agent_updates=K, trajectories=trajs, eval_trajectory=ev, do_eval=False)
    g = met["_grad"].cpu().numpy().astype(np.float64)
    og = o["grad"].numpy()
    for name, (off, cnt, shp) in c.layout.offsets.items():
        e = rel_err(g[off:off + cnt], og[off:off + cnt])
        print(f"  block {name:5s} rel err {e:.2e}")
        assert e < 2e-2, f"meta-gradient block {name}: rel err {e:.3e}"
    l2 = np.linalg.norm(g - og) / np.linalg.norm(og)
    print(f"[tc {mode}] relative L2 error of the meta-gradient {l2:.2e}")
    assert l2 < 1e-2


def _fac_plane(x, i, L, R):
    """Gate plane i of the saved-gate tensor [L][R/32][4 planes][32 chunks][32 rows][8] (csrc/tc.cuh::fac_index) -> [L][R][256]."""
    R32 = (R + 31) // 32
    v = x.reshape(L, R32, 4, 32, 32, 8)[:, :, i].permute(0, 1, 3, 2, 4).reshape(L, R32 * 32, 256)
    return v[:, :R]


def _from_rb32(x, L, R):
    """RB32 layout [L][R/32][32 chunks][32 rows][8] -> [L][R][256] (inverse of csrc/tc.cuh::rb32_index)."""
    R32 = (R + 31) // 32
    v = x.reshape(L, R32, 32, 32, 8).permute(0, 1, 3, 2, 4).reshape(L, R32 * 32, 256)
    return v[:, :R]
