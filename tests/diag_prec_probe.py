"""Which quantisation explains the tensor-core path's meta-gradient error?  CPU experiment on the fp64 oracle with
straight-through fp16 rounding injected into the GRU forward (a diagnostic next to the tests, like tests/diag_probe.py: only tests/, smoke() and bench.py may import oracle/)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.dirname(__file__))
import numpy as np, torch
import oracle.lpg as olpg
import oracle.agents as oag_mod
from oracle import prng
from oracle.agents import AgentTables
from oracle.meta import lpg_meta_grad_train_step as o_step
from helpers import Case

def q16(x):   # straight-through fp16 rounding
    return x + (x.to(torch.float16).to(x.dtype) - x).detach()

MODE = {"wh": False, "h": False, "gates": False, "hhl": False}

def lpg_forward_q(layout, flat, r, d, pi, yt, yt1, step, lifetime):
    P = layout.unpack(flat); H = layout.H; dt = flat.dtype; d = d.to(dt)
    def embed(y): return torch.relu(y @ P["e_w0"] + P["e_b0"]) @ P["e_w1"] + P["e_b1"]
    pyt = embed(yt); pyt1 = embed(yt1) * (1.0 - d)
    cols = [r.to(dt), d, pi, pyt, pyt1]
    if layout.lifetime_conditioning:
        cols += [step.to(dt)[:, None].expand_as(d), lifetime.to(dt)[:, None].expand_as(d)]
    x = torch.stack(cols, dim=-1)
    gi = x @ P["Wi"] + P["bi"]
    Wh = q16(P["Wh"]) if MODE["wh"] else P["Wh"]
    B, L = d.shape
    h = torch.zeros(B, H, dtype=dt); outs = [None] * L
    for t in reversed(range(L)):
        h = h * (1.0 - d[:, t:t + 1])
        gh = h @ Wh
        rg = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
        zg = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        hn = gh[:, 2 * H:] + P["bhn"]
        ng = torch.tanh(gi[:, t, 2 * H:] + rg * hn)
        h = (1.0 - zg) * ng + zg * h
        if MODE["h"]: h = q16(h)
        outs[t] = h
    y = torch.relu(torch.stack(outs, dim=1))
    pi_hat = y @ P["w_pi"] + P["b_pi"]
    y_hat = torch.softmax(y @ P["W_y"] + P["b_y"], dim=-1)
    return pi_hat, y_hat

def run(c, K, trajs=None, ev=None):
    dt = torch.float64
    oag = AgentTables(torch.tensor(c.actor).to(dt), torch.tensor(c.critic).to(dt), torch.tensor(c.steps.astype(np.int64)))
    s0 = c.oro.batch_reset(None, c.p, c.w)
    return o_step(prng.PRNGKey(21), c.layout, torch.tensor(c.lpg).to(dt), oag, torch.tensor(c.value).to(dt), c.oro, c.p,
                  s0, c.life, num_agent_updates=K, trajectories=trajs, eval_trajectory=ev, do_eval=False)

def main():
    K = 5
    c = Case("all_shortlife", n=4, w=64, seed=7, table_scale=0.3, lifetimes=[250, 3, 250, 250], steps=[0, 0, 17, 246])
    ref = run(c, K)
    trajs, ev = ref["trajectories"], ref["eval_trajectory"]
    og = ref["grad"].numpy()
    orig = olpg.lpg_forward
    import oracle.agents, oracle.meta
    mods = [m for m in sys.modules.values() if m and getattr(m, "lpg_forward", None) is orig]
    for m in mods: m.lpg_forward = lpg_forward_q
    for cfg in ({}, {"wh": True}, {"h": True}, {"wh": True, "h": True}):
        for k in MODE: MODE[k] = cfg.get(k, False)
        o = run(c, K, trajs, ev)
        g = o["grad"].numpy()
        l2 = np.linalg.norm(g - og) / np.linalg.norm(og)
        blocks = {n: np.abs(g[o_:o_ + cnt] - og[o_:o_ + cnt]).max() / np.abs(og[o_:o_ + cnt]).max() for n, (o_, cnt, s) in c.layout.offsets.items()}
        print(cfg, f"rel L2 {l2:.2e}", {k: f"{v:.1e}" for k, v in blocks.items() if k in ("Wh", "Wi", "bi", "bhn", "w_pi", "W_y", "e_w0")})

main()
