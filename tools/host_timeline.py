#!/usr/bin/env python
"""Host-side enqueue timeline of one end-to-end meta-step (bench.py's `e2e` leg: host key in, metrics out every
step).  Stamps every C-ABI call with the host clock relative to the start of the step and reports where the host
spends the time the GPU waits for at the start of a synchronous step.

    python tools/host_timeline.py [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from to_ued_b200 import _lib  # noqa: E402
from to_ued_b200.util import prng  # noqa: E402
from to_ued_b200.experiments.parse_args import parse_args  # noqa: E402
from to_ued_b200.environments.level_sampler import LevelSampler  # noqa: E402
from to_ued_b200.meta.meta import create_lpg_train_state, make_lpg_train_step  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    args = parse_args(["--env_mode", "all_shortlife", "--num_agents", "512", "--num_mini_batches", "2"])
    rng = prng.PRNGKey(args.seed)
    rng, lpg_rng, buffer_rng = prng.split(rng, 3)
    train_state = create_lpg_train_state(lpg_rng, args)
    sampler = LevelSampler(args)
    buf = sampler.initialize_buffer(buffer_rng)
    rng, _rng = prng.split(rng, 2)
    buf, agents, vcs = sampler.initial_sample(_rng, buf, args.num_agents, True)
    step_fn = make_lpg_train_step(args, sampler)

    stamps = []
    raw_call = _lib.call

    def stamped(name, *a):
        t = time.perf_counter()
        r = raw_call(name, *a)
        stamps.append((name, t, time.perf_counter()))
        return r

    def one_step(rng, train_state, agents, vcs, buf, marks):
        marks["t0"] = time.perf_counter()
        rng, _rng = prng.split(rng, 2)
        train_state, agents, vcs, metrics = step_fn(rng=_rng, lpg_train_state=train_state, agent_states=agents,
                                                    value_critic_states=vcs)
        marks["step_fn"] = time.perf_counter()
        rng, _rng = prng.split(rng, 2)
        buf, agents, vcs = sampler.sample(_rng, buf, agents, vcs)
        marks["sample"] = time.perf_counter()
        flat = [v for k, v in metrics.items() if not isinstance(v, dict)] + \
               [vv for v in metrics.values() if isinstance(v, dict) for vv in v.values()]
        torch.stack([f.reshape(()) for f in flat]).cpu()
        marks["read"] = time.perf_counter()
        return rng, train_state, agents, vcs, buf

    state = (rng, train_state, agents, vcs, buf)
    for _ in range(4):
        state = one_step(*state, {})
    torch.cuda.synchronize()
    _lib.call = stamped
    rows = []
    for i in range(steps):
        stamps.clear()
        marks = {}
        state = one_step(*state, marks)
        t0 = marks["t0"]
        first = {}
        for name, a, b in stamps:
            first.setdefault(name, (a - t0, b - a))
        rows.append((marks, list(stamps), first))
    _lib.call = raw_call
    marks, st, first = rows[-1]
    t0 = marks["t0"]
    print(f"step wall {1e3 * (marks['read'] - t0):.3f} ms: step_fn returns at {1e3 * (marks['step_fn'] - t0):.3f}, "
          f"sampler.sample at {1e3 * (marks['sample'] - t0):.3f}, metrics read at {1e3 * (marks['read'] - t0):.3f}")
    print(f"{len(st)} C-ABI calls; host time inside them {1e3 * sum(b - a for _, a, b in st):.3f} ms")
    print("first 40 calls of the step (ms since step start, call duration us):")
    for name, a, b in st[:40]:
        print(f"  {1e3 * (a - t0):8.3f}  {1e6 * (b - a):7.1f}  {name}")
    walls = [1e3 * (r[0]["read"] - r[0]["t0"]) for r in rows]
    firsts = [1e3 * r[2].get("toued_rollout", (0, 0))[0] for r in rows]
    print("per-step wall ms:", " ".join(f"{w:.2f}" for w in walls))
    print("first toued_rollout enqueued at ms:", " ".join(f"{w:.3f}" for w in firsts))
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    state = one_step(*state, {})
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
