"""Optimisers (reference models/optim.py:5-34).

``create_optimizer("SGD", ...)`` describes optax.chain(clip_by_global_norm, scale(lr), scale(-1));
for the tabular agents this is fused into toued_agent_update.  ``"Adam"`` is
optax.chain(scale_by_adam(), scale(lr), scale(-1)) (no clipping on this branch, Q9) and runs in
toued_adam on the flat LPG parameter vector."""
from __future__ import annotations

import torch

from .. import _lib


class SGD:
    """optax.chain(clip_by_global_norm(max_grad_norm), scale(lr), scale(-1)) (models/optim.py:6-11).  For the tabular
    agents the same chain is fused into toued_agent_update; as the LPG optimiser (``--lpg_opt SGD``) it runs in
    toued_sgd_clip on the flat parameter vector.  The optimiser state keeps Adam's keys so that the captured-graph step
    and the checkpoint code treat both optimisers alike: ``mu`` is the f32[1] scratch that receives |g|^2."""
    name = "SGD"

    def __init__(self, learning_rate: float, max_grad_norm: float):
        self.learning_rate, self.max_grad_norm = learning_rate, max_grad_norm

    def init(self, params: torch.Tensor):
        return {"mu": torch.zeros(1, dtype=torch.float32, device=params.device),
                "nu": torch.zeros(1, dtype=torch.float32, device=params.device), "count": 0}

    def update_(self, params: torch.Tensor, grad: torch.Tensor, state: dict) -> dict:
        _lib.call("toued_sgd_clip", _lib.ptr(params), _lib.ptr(grad), _lib.ptr(state["mu"]), params.numel(),
                  float(self.learning_rate), float(self.max_grad_norm), _lib.stream_ptr())
        return {"mu": state["mu"], "nu": state["nu"], "count": state["count"] + 1}

    def update_dev_(self, params, grad, mu, nu, count_dev) -> None:
        _lib.call("toued_sgd_clip", _lib.ptr(params), _lib.ptr(grad), _lib.ptr(mu), params.numel(),
                  float(self.learning_rate), float(self.max_grad_norm), _lib.stream_ptr())


class Adam:
    name = "Adam"

    def __init__(self, learning_rate: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8):
        self.learning_rate, self.b1, self.b2, self.eps = learning_rate, b1, b2, eps

    def init(self, params: torch.Tensor):
        return {"mu": torch.zeros_like(params), "nu": torch.zeros_like(params), "count": 0}

    def update_(self, params: torch.Tensor, grad: torch.Tensor, state: dict) -> dict:
        """In-place Adam step on the GPU; returns the new optimizer state."""
        count = state["count"] + 1
        _lib.call("toued_adam", _lib.ptr(params), _lib.ptr(grad), _lib.ptr(state["mu"]), _lib.ptr(state["nu"]),
                  params.numel(), count, float(self.learning_rate), float(self.b1), float(self.b2), float(self.eps),
                  _lib.stream_ptr())
        return {"mu": state["mu"], "nu": state["nu"], "count": count}

    def update_dev_(self, params: torch.Tensor, grad: torch.Tensor, mu: torch.Tensor, nu: torch.Tensor,
                    count_dev: torch.Tensor) -> None:
        """The same step with the update count in device memory (int32[1], incremented by the call): capturable in a
        CUDA graph (meta/graph.py)."""
        _lib.call("toued_adam_dev", _lib.ptr(params), _lib.ptr(grad), _lib.ptr(mu), _lib.ptr(nu), _lib.ptr(count_dev),
                  params.numel(), float(self.learning_rate), float(self.b1), float(self.b2), float(self.eps),
                  _lib.stream_ptr())


def create_optimizer(optimizer: str, learning_rate: float, max_grad_norm: float):
    if optimizer == "SGD":
        return SGD(learning_rate, max_grad_norm)
    elif optimizer == "Adam":
        return Adam(learning_rate)
    raise ValueError(f"Unknown optimizer: {optimizer}")
