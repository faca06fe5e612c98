"""LPG train-state construction and train-step binding (reference meta/meta.py:10-52)."""
from __future__ import annotations

from functools import partial

import numpy as np

from ..util.data import LpgHyperparams
from ..models.lpg import LPG
from ..models.optim import create_optimizer
from .train import LPGTrainState, lpg_meta_grad_train_step


def create_lpg_train_state(rng, args, single_env=False, device="cuda"):
    """meta/meta.py:10-30.  TrainState for meta-gradients; ESTrainState when ``args.use_es``."""
    lpg_model = LPG(embedding_net_width=args.lpg_embedding_net_width, gru_width=args.lpg_gru_width,
                    target_width=args.lpg_target_width, lifetime_conditioning=args.lifetime_conditioning)
    params = lpg_model.init(rng, device=device)
    tx = create_optimizer(args.lpg_opt, args.lpg_learning_rate, args.lpg_max_grad_norm)
    train_state = LPGTrainState(lpg_model, params, tx)
    if not args.use_es or single_env:
        return train_state
    from .es import create_es_train_state
    return create_es_train_state(rng, args, train_state)


def make_lpg_train_step(args, level_sampler, cuda_graph=None):
    """meta/meta.py:33-52: bind rollout manager, mini-batches and hyper-parameters.  ``cuda_graph``: None = the package
    default (to_ued_b200.CUDA_GRAPH), False = eager launch-by-launch step (value semantics, fresh output tensors)."""
    lpg_hypers = LpgHyperparams.from_run_args(args)
    if args.use_es:
        from .es import lpg_es_train_step
        lpg_hypers = lpg_hypers.replace(num_agent_updates=level_sampler.max_lifetime)
        return partial(lpg_es_train_step, rollout_manager=level_sampler.rollout_manager,
                       num_mini_batches=args.num_mini_batches, lpg_hypers=lpg_hypers)
    bound = dict(rollout_manager=level_sampler.rollout_manager, num_mini_batches=args.num_mini_batches,
                 gamma=args.gamma, gae_lambda=args.gae_lambda, lpg_hypers=lpg_hypers)
    import to_ued_b200
    if to_ued_b200.CUDA_GRAPH if cuda_graph is None else cuda_graph:
        # one captured CUDA graph per meta-step instead of ~125 launches enqueued from Python (meta/graph.py);
        # buffers are donated: the returned states alias static device buffers
        from .graph import GraphedMetaGradStep
        return GraphedMetaGradStep(**bound)
    return partial(lpg_meta_grad_train_step, **bound)
