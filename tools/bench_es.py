#!/usr/bin/env python
"""Secondary measurement: one TA-LPG / ES meta-step (BASELINE config 4: env_mode=all_vrandlife,
lifetime conditioning, population = 2 x num_agents), timed as the difference between a 4-step and a 2-step run.

    python tools/bench_es.py [num_agents]          # TOUED_ES_PRECISION=fp32 selects the exact SIMT per-candidate kernel
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import train  # noqa: E402
import to_ued_b200  # noqa: E402
from to_ued_b200.util import prng  # noqa: E402
from to_ued_b200.experiments.parse_args import parse_args  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    args = parse_args(["--env_mode", "all_vrandlife", "--num_agents", str(n), "--num_mini_batches", "1", "--use_es",
                       "--lifetime_conditioning", "--lpg_learning_rate", "0.01", "--train_steps", "1"])
    train.make_train(args)(prng.PRNGKey(0))
    torch.cuda.synchronize()
    times = {}
    for steps in (2, 4):
        args.train_steps = steps
        t0 = time.perf_counter()
        hist, _, _ = train.make_train(args)(prng.PRNGKey(0))
        torch.cuda.synchronize()
        times[steps] = time.perf_counter() - t0
    per = (times[4] - times[2]) / 2
    env_steps = 2 * n * (250 * 64 * 20 + 64 * 750)          # lifetime 250, 64 workers, eval cap 750
    print(f"ES[{to_ued_b200.ES_PRECISION}] meta-step, population {2 * n}: {per:.3f} s, {env_steps / 1e6:.0f} M env-steps -> "
          f"{env_steps / per / 1e6:.0f} M env-steps/s; fitness mean {float(hist[-1]['fitness']['mean']):.3f}")


if __name__ == "__main__":
    main()
