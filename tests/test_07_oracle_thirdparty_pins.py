"""Independent pins of the oracle's [3P-recall] pieces.  The reference's dependencies (flax 0.6.11 GRUCell, optax 0.1.5
adam / clip_by_global_norm, jax.nn.softmax) cannot be installed here, so the oracle restates their published algorithms
(DESIGN.md section 4).  PyTorch ships independent implementations of the SAME published definitions; these tests hold the
oracle against them on random inputs:

  * flax.linen.GRUCell (models/lpg.py:11-30 uses it) has the cuDNN / PyTorch gate structure
    r, z = sigmoid(.), n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) n + z h        -> torch.nn.GRUCell
  * optax.scale_by_adam(b1=.9, b2=.999, eps=1e-8, eps_root=0) -> scale(lr) -> scale(-1)          -> torch.optim.Adam
  * optax.clip_by_global_norm(max_norm) -> scale(lr) -> scale(-1)                                 -> clip_grad_norm_ + SGD
  * the reverse-scan LPGGRU with carry resets at terminal states                                  -> a torch.nn.GRUCell loop

This does not pin the oracle on the JAX reference itself ("parity unpinned" stands) but it removes the possibility that
the recalled third-party semantics are mis-stated in a way that only the reference would reveal."""
import numpy as np
import torch

from oracle.lpg import LPGLayout, init_lpg_params, lpg_forward
from oracle.meta import Adam as OAdam, SGDClip
from oracle.agents import clip_sgd


def _torch_cell(P, X, H):
    """torch.nn.GRUCell carrying the oracle's parameters (torch packs gates as (r, z, n) rows, like flax's (ir, iz, in))."""
    cell = torch.nn.GRUCell(X, H, bias=True, dtype=torch.float64)
    with torch.no_grad():
        cell.weight_ih.copy_(P["Wi"].T)                       # [3H, X]
        cell.weight_hh.copy_(P["Wh"].T)                       # [3H, H]
        cell.bias_ih.copy_(P["bi"])
        bhh = torch.zeros(3 * H, dtype=torch.float64)
        bhh[2 * H:] = P["bhn"]                                # flax: only the hn projection carries a hidden-side bias
        cell.bias_hh.copy_(bhh)
    return cell


def test_lpg_gru_matches_torch_grucell_reverse_scan():
    for cond in (False, True):
        layout = LPGLayout(lifetime_conditioning=cond)
        rs = np.random.RandomState(5)
        flat = torch.tensor(init_lpg_params(layout, 3) + 0.05 * rs.randn(layout.size).astype(np.float32), dtype=torch.float64)
        P = layout.unpack(flat)
        B, L, H, Y = 6, 9, layout.H, layout.Y
        r = torch.tensor(rs.randn(B, L)); d = torch.tensor((rs.rand(B, L) < 0.2).astype(np.float64))
        pi = torch.tensor(rs.rand(B, L)); yt = torch.tensor(rs.rand(B, L, Y)); yt1 = torch.tensor(rs.rand(B, L, Y))
        step = torch.tensor(rs.randint(0, 200, B)); life = torch.tensor(rs.randint(1, 250, B))
        pi_hat, y_hat = lpg_forward(layout, flat, r, d, pi, yt, yt1, step, life)
        # independent forward: embedding MLP, torch.nn.GRUCell scanned in reverse with carry reset, heads
        embed = lambda y: torch.relu(y @ P["e_w0"] + P["e_b0"]) @ P["e_w1"] + P["e_b1"]
        cols = [r, d, pi, embed(yt), embed(yt1) * (1 - d)]
        if cond:
            cols += [step.double()[:, None].expand(B, L), life.double()[:, None].expand(B, L)]
        x = torch.stack(cols, -1)
        cell = _torch_cell(P, layout.X, H)
        h = torch.zeros(B, H, dtype=torch.float64)
        hs = [None] * L
        with torch.no_grad():
            for t in reversed(range(L)):
                h = cell(x[:, t], h * (1 - d[:, t:t + 1]))
                hs[t] = h
        yy = torch.relu(torch.stack(hs, 1))
        np.testing.assert_allclose(pi_hat.detach().numpy(), (yy @ P["w_pi"] + P["b_pi"]).detach().numpy(), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(y_hat.detach().numpy(), torch.softmax(yy @ P["W_y"] + P["b_y"], -1).detach().numpy(), rtol=1e-12, atol=1e-12)


def test_oracle_adam_matches_torch_adam():
    rs = np.random.RandomState(1)
    p0 = rs.randn(257)
    opt = OAdam(257, 3e-3, dtype=torch.float64)
    p = torch.tensor(p0)
    tp = torch.nn.Parameter(torch.tensor(p0))
    topt = torch.optim.Adam([tp], lr=3e-3, betas=(0.9, 0.999), eps=1e-8)
    for i in range(7):
        g = torch.tensor(rs.randn(257) * (10.0 ** rs.randint(-4, 2)))
        p = opt.step(p, g)
        tp.grad = g.clone()
        topt.step()
        np.testing.assert_allclose(p.numpy(), tp.detach().numpy(), rtol=1e-12, atol=1e-14)


def test_oracle_clip_sgd_matches_torch_clip_grad_norm():
    rs = np.random.RandomState(2)
    for scale in (1e-3, 1.0, 30.0):                           # below, near and far above max_norm
        p0, g = rs.randn(3, 11, 5), rs.randn(3, 11, 5) * scale
        got = clip_sgd(torch.tensor(p0), torch.tensor(g), 0.7, 0.5).numpy()
        flat = SGDClip(0.7, 0.5)
        for a in range(3):                                    # per agent: clip_grad_norm_ + plain SGD
            tp = torch.nn.Parameter(torch.tensor(p0[a]))
            tp.grad = torch.tensor(g[a])
            torch.nn.utils.clip_grad_norm_([tp], 0.5)
            # torch divides by (norm + 1e-6); optax by norm: identical to 1e-6 relative, so compare at that level
            torch.optim.SGD([tp], lr=0.7).step()
            np.testing.assert_allclose(got[a], tp.detach().numpy(), rtol=3e-6, atol=1e-9)
            np.testing.assert_allclose(flat.step(torch.tensor(p0[a]).flatten(), torch.tensor(g[a]).flatten()).numpy(),
                                       got[a].reshape(-1), rtol=1e-12, atol=1e-14)
