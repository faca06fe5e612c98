#!/usr/bin/env python
"""Device-side timeline of one free-running meta-step (bench.py's workload): CUDA events recorded at the phase
boundaries of `lpg_meta_grad_train_step` (to_ued_b200.PHASE_EVENTS), per agent chunk, in ms since the start of
the step.  Shows how the two chunks' chains interleave and where the step's wall time goes.

    python tools/phase_timeline.py [steps]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import to_ued_b200  # noqa: E402
from to_ued_b200.util import prng  # noqa: E402
from to_ued_b200.experiments.parse_args import parse_args  # noqa: E402
from to_ued_b200.environments.level_sampler import LevelSampler  # noqa: E402
from to_ued_b200.meta.meta import create_lpg_train_state, make_lpg_train_step  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    to_ued_b200.PHASE_FINE = os.environ.get("FINE", "0") != "0"
    args = parse_args(["--env_mode", "all_shortlife", "--num_agents", os.environ.get("TOUED_AGENTS", "512"), "--num_mini_batches",
                       os.environ.get("TOUED_BENCH_MINI_BATCHES", "2")])
    rng = prng.PRNGKey(args.seed)
    rng, lpg_rng, buffer_rng = prng.split(rng, 3)
    train_state = create_lpg_train_state(lpg_rng, args)
    sampler = LevelSampler(args)
    buf = sampler.initialize_buffer(buffer_rng)
    rng, _rng = prng.split(rng, 2)
    buf, agents, vcs = sampler.initial_sample(_rng, buf, args.num_agents, True)
    step_fn = make_lpg_train_step(args, sampler)

    def one_step(rng, train_state, agents, vcs, buf):
        rng, _rng = prng.split(rng, 2)
        train_state, agents, vcs, metrics = step_fn(rng=_rng, lpg_train_state=train_state, agent_states=agents,
                                                    value_critic_states=vcs)
        rng, _rng = prng.split(rng, 2)
        buf, agents, vcs = sampler.sample(_rng, buf, agents, vcs)
        return rng, train_state, agents, vcs, buf

    state = (rng, train_state, agents, vcs, buf)
    for _ in range(4):
        state = one_step(*state)
    torch.cuda.synchronize()
    per_step = []
    for _ in range(steps):                                # free-running: no synchronisation between the steps
        to_ued_b200.PHASE_EVENTS = []
        state = one_step(*state)
        per_step.append(to_ued_b200.PHASE_EVENTS)
    to_ued_b200.PHASE_EVENTS = None
    torch.cuda.synchronize()
    t_prev_end = None
    for i, evs in enumerate(per_step):
        start = evs[0][2]
        end = evs[-1][2]
        gap = "" if t_prev_end is None else f" (starts {t_prev_end.elapsed_time(start):+.3f} ms after the previous step's end)"
        print(f"step {i}: {start.elapsed_time(end):.3f} ms{gap}")
        t_prev_end = end
    evs = per_step[-1]
    start = evs[0][2]
    fine = [(lab, e) for lab, c, e in evs if c == -2]
    if fine:                                              # FINE=1: inside the agent updates (one chunk)
        prev = start
        print("fine: " + "  ".join(f"{lab} +{(lambda d: d)(prev.elapsed_time(e)):.3f}" + ("" if (prev := e) is None else "") for lab, e in fine))
    chunks = sorted({c for _, c, _ in evs if c >= 0})
    for c in chunks:
        print(f"chunk {c}: " + "  ".join(f"{lab} {start.elapsed_time(e):.2f}" for lab, cc, e in evs if cc == c))
    print(f"end {start.elapsed_time(evs[-1][2]):.2f}")


if __name__ == "__main__":
    main()
