"""Worker of tests/test_30_multi_gpu.py (not a pytest file): runs a few meta-steps of the train.py loop and saves what
must not depend on the number of ranks.  Launched once as a plain process (world 1) and once under torchrun (world 2).

    python tests/mgpu_worker.py <metagrad|groove|es> <out.pt>"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N_AGENTS = 8


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("TOUED_WORKER_TIMEOUT", "240")), exit=True)   # a hung rank prints its stack
    mode, out = sys.argv[1], sys.argv[2]
    import train as train_mod
    from to_ued_b200.util import prng, dist as udist
    from to_ued_b200.environments.level_sampler import LevelSampler
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.experiments.logging import to_host
    from to_ued_b200.meta.meta import create_lpg_train_state, make_lpg_train_step
    rank, world = train_mod._init_distributed()
    if world == 1:
        torch.cuda.set_device(0)
    flags = ["--env_mode", "all_vrandlife", "--num_agents", str(N_AGENTS), "--num_mini_batches", "1", "--seed", "3"]
    steps = 14
    if mode == "groove":
        flags += ["--score_function", "alg_regret", "--buffer_size", "64"]
        steps = 8
    if mode == "es":
        flags += ["--use_es", "--lifetime_conditioning"]
        steps = 2
    args = parse_args(flags)
    rng = prng.PRNGKey(args.seed)
    rng, lpg_rng, buffer_rng = prng.split(rng, 3)
    train_state = create_lpg_train_state(lpg_rng, args)
    sampler = LevelSampler(args)
    buf = sampler.initialize_buffer(buffer_rng)
    rng, _rng = prng.split(rng, 2)
    buf, agents, vcs = sampler.initial_sample(_rng, buf, args.num_agents, not args.use_es)
    step_fn = make_lpg_train_step(args, sampler)
    hist = []
    for t in range(steps):
        rng, _rng = prng.split(rng, 2)
        train_state, agents, vcs, metrics = step_fn(rng=_rng, lpg_train_state=train_state, agent_states=agents,
                                                    value_critic_states=vcs)
        rng, _rng = prng.split(rng, 2)
        buf, agents, vcs = sampler.sample(_rng, buf, agents, vcs)
        m = to_host({k: v for k, v in metrics.items() if not k.startswith("_")})
        rec = {"metrics": m,
               "lifetime": udist.all_gather_host(agents.level.lifetime.astype(np.int32)),
               "buffer_id": udist.all_gather_host(agents.level.buffer_id.astype(np.int32)),
               "walls": udist.all_gather_device(agents.level.packed).cpu().numpy(),     # the whole LevelRec of every agent
               "host_step": udist.all_gather_host(agents.host_step.astype(np.int32)),
               "step": udist.all_gather_device(agents.actor_state.step).cpu().numpy(),
               "actor": udist.all_gather_device(agents.actor_state.params).cpu().numpy(),
               "env_state": udist.all_gather_device(agents.env_state.packed).cpu().numpy()}
        if t == 0:
            rec["lpg"] = (train_state.mean if args.use_es else train_state.params).cpu().numpy()
        if buf is not None:
            rec.update(score=buf.score.copy(), active=buf.active.copy(), new=buf.new.copy())
        if "_fitness" in metrics:
            rec["fitness"] = metrics["_fitness"].cpu().numpy()
        hist.append(rec)
    params = train_state.mean if args.use_es else train_state.params
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"hist": hist, "lpg": params.cpu().numpy(), "world": world}, out)
    udist.shutdown(step_fn)


if __name__ == "__main__":
    main()
