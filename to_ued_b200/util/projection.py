"""Simplex projection (reference util/projection.py:9-38) on the GPU."""
import torch

from .. import _lib


def projection_simplex(x: torch.Tensor, max_nz: int) -> torch.Tensor:
    """Projection onto the unit simplex keeping only the first ``max_nz`` coordinates non-zero."""
    out = x.detach().to(torch.float32).contiguous().clone()
    _lib.call("toued_projection_simplex", _lib.ptr(out), out.numel(), int(max_nz), _lib.stream_ptr())
    return out
