"""Double-oracle (minimax-regret UED) meta-training entry point (reference train_do.py:13-102),
B200-native, with the intended semantics of SURVEY.md §3.4 (the reference script does not run, Q7)."""
import sys

import numpy as np
import torch

from to_ued_b200.util import prng
from to_ued_b200.environments.nash_sampler import NashSampler
from to_ued_b200.experiments.parse_args import parse_args
from to_ued_b200.meta.meta import make_lpg_train_step, create_lpg_train_state
from to_ued_b200.environments.level_sampler import _set_rows
from to_ued_b200.environments.gridworld.gridworld import EnvParams
from to_ued_b200.util.data import Level


def _write_level(buffer, t, level: Level):
    lv = buffer.level
    params = EnvParams(**{f: _set_rows(getattr(lv.env_params, f), [t], getattr(level.env_params, f))
                          for f in lv.env_params.__dataclass_fields__})
    new_level = Level(params, _set_rows(lv.lifetime, [t], level.lifetime), lv.buffer_id)
    return buffer.replace(level=new_level, active=_set_rows(buffer.active, [t], True))


def make_train(args):
    def _train_fn(rng):
        B = args.buffer_size
        dev = "cuda"
        train_nash = torch.zeros(B, device=dev); train_nash[0] = 1
        eval_nash = torch.zeros(B, device=dev); eval_nash[0] = 1
        sampler = NashSampler(args)
        rng, buffer_rng, train_rng = prng.split(rng, 3)
        train_buffer, eval_buffer = sampler.initialize_buffers(buffer_rng)
        train_state = create_lpg_train_state(train_rng, args)
        step_fn = make_lpg_train_step(args, sampler)
        history = []
        for t in range(1, B):                                  # train_do.py:75-77
            rng, _rng = prng.split(rng, 2)
            agents, vcs = sampler.get_training_levels(_rng, train_buffer, train_nash, create_value_critic=not args.use_es)
            rng, _rng = prng.split(rng, 2)
            train_state, agents, vcs, metrics = step_fn(rng=_rng, lpg_train_state=train_state, agent_states=agents,
                                                        value_critic_states=vcs)
            rng, br_train_rng, br_eval_rng, nash_rng = prng.split(rng, 4)
            new_train = sampler.get_train_br(br_train_rng, train_state, eval_nash, eval_buffer)
            new_eval, eval_regret = sampler.get_eval_br(br_eval_rng, train_state)
            train_buffer = _write_level(train_buffer, t, new_train)
            eval_buffer = _write_level(eval_buffer, t, new_eval)
            train_nash, eval_nash, game = sampler.compute_nash(nash_rng, train_state, train_buffer, eval_buffer)
            metrics["GT"] = {"eval_regret": float(eval_regret)}
            history.append(metrics)
        return history, train_state, train_buffer
    return _train_fn


def run_training_experiment(args):
    if args.log:                                     # train_do.py:89-95 (local run directory instead of WandB)
        from to_ued_b200.experiments.logging import init_logger
        print(f"[to_ued_b200] --log: run directory {init_logger(args)}")
    metrics, train_state, level_buffer = make_train(args)(prng.PRNGKey(args.seed))
    torch.cuda.synchronize()
    from to_ued_b200.experiments.logging import to_host
    print([to_host(m) for m in metrics])
    if args.log:
        from to_ued_b200.experiments.logging import log_results
        print("[to_ued_b200] checkpoints:", log_results(args, metrics, train_state, level_buffer))
    return metrics, train_state, level_buffer


def main(cmd_args=sys.argv[1:]):
    return run_training_experiment(parse_args(cmd_args))


if __name__ == "__main__":
    main()
