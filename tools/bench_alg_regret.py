#!/usr/bin/env python
"""Secondary measurement: PLR algorithmic-regret scoring of a full batch of agents
(LevelSampler._compute_algorithmic_regret: an A2C antagonist trained for the level's whole lifetime + two evaluations).

    python tools/bench_alg_regret.py [env_mode] [num_agents]
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from to_ued_b200.util import prng  # noqa: E402
from to_ued_b200.experiments.parse_args import parse_args  # noqa: E402
from to_ued_b200.environments.level_sampler import LevelSampler  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "mazes"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    args = parse_args(["--env_mode", mode, "--num_agents", str(n), "--num_mini_batches", "1", "--score_function", "alg_regret",
                       "--buffer_size", str(4 * n)])
    sampler = LevelSampler(args)
    buf = sampler.initialize_buffer(prng.PRNGKey(1))
    buf, agents, _ = sampler.initial_sample(prng.PRNGKey(2), buf, n, True)
    keys = prng.split(prng.PRNGKey(3), n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    score = sampler._compute_algorithmic_regret(keys, agents)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    K, W, L = sampler.max_lifetime, args.env_workers, args.train_rollout_len
    env_steps = n * (K * W * L + 2 * W * sampler.max_rollout_len)
    print(f"{mode}, {n} agents: A2C antagonist {K} updates x {W} workers x {L} steps + 2 evaluations = {env_steps / 1e6:.1f} M "
          f"env-steps in {dt:.2f} s -> {env_steps / dt / 1e6:.1f} M env-steps/s; mean score {score.mean():.3f}")


if __name__ == "__main__":
    main()
