"""Regenerate tests/golden/oracle_v1.npz: small known-answer vectors of the ORACLE (the reference itself
ships no fixtures and cannot be run here).  They pin the RNG contract, the level
generator, one rollout and one LPG meta-gradient so that any later change of the oracle is visible.

    python tests/golden/make_golden.py
"""
import os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build():
    from oracle import prng, configs
    from oracle.agents import AgentTables
    from oracle.meta import lpg_meta_grad_train_step
    from helpers import Case
    torch.set_num_threads(1)
    out = {}
    k = prng.PRNGKey(1234)
    out["split"] = prng.split(k, 3)
    out["uniform"] = prng.uniform(prng.split(k, 2), (5,), -1.0, 2.0)
    out["randint"] = prng.randint(k, (6,), 3, 11)
    out["shuffle"] = prng.shuffle(k, 20)
    p, life = configs.reset_env_params(prng.split(prng.PRNGKey(7), 4), "all_vrandlife")
    out["lvl_grid"], out["lvl_nobj"], out["lvl_steps"] = p.grid_size, p.n_objs, p.max_steps_in_episode
    out["lvl_start"], out["lvl_objpos"], out["lvl_life"] = p.start_pos, p.static_obj_poss, life
    out["lvl_walls"] = np.packbits(p.walls, axis=1)
    c = Case("debug", n=3, w=8, L=6, seed=2, table_scale=0.3)
    s0 = c.oro.batch_reset(None, c.p, c.w)
    traj, s1, ret = c.oro.batch_rollout(c.keys, c.actor, c.p, s0)
    out["ro_action"], out["ro_obs"], out["ro_done"] = traj.action.astype(np.int8), traj.obs_idx.astype(np.int16), traj.done
    out["ro_reward"], out["ro_return"] = traj.reward, ret
    dt = torch.float64
    ag = AgentTables(torch.tensor(c.actor).to(dt), torch.tensor(c.critic).to(dt), torch.zeros(3, dtype=torch.long))
    o = lpg_meta_grad_train_step(prng.PRNGKey(4), c.layout, torch.tensor(c.lpg).to(dt), ag, torch.tensor(c.value).to(dt),
                                 c.oro, c.p, s0, c.life, num_agent_updates=2, do_eval=False)
    g = o["grad"].numpy()
    out["grad_norm"] = np.array([np.linalg.norm(g)])
    out["grad_head"] = g[:16]
    out["grad_block_norms"] = np.array([np.linalg.norm(g[off:off + n]) for _, (off, n, _) in c.layout.offsets.items()])
    out["lpg_loss"] = np.array([float(o["metrics"]["lpg_loss"])])
    return out


if __name__ == "__main__":
    d = build()
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz"), **d)
    print("wrote", sorted(d))
