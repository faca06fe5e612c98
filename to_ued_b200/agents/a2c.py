"""A2C antagonist (reference agents/a2c.py:12-125), used by the algorithmic-regret level score."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class A2CHyperparams:
    """agents/a2c.py:12-16"""
    gamma: float
    gae_lambda: float
    entropy_coeff: float


def train_a2c_agent(rng, agent_state, rollout_manager, num_train_steps, hypers: A2CHyperparams):
    """agents/a2c.py:79-125 — CUDA kernels for the A2C update are the next widening step."""
    raise NotImplementedError("A2C antagonist kernels are not built yet (score_function=alg_regret)")
