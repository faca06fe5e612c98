"""Rank plumbing for the data-parallel agent axis.

Rank r of a world of G holds agents [r * n, (r + 1) * n) of the global batch.  Three kinds of exchange exist:
  * the LPG meta-gradient / ES gradient: one all-reduce on the device (meta/train.py, meta/es.py);
  * per-agent device results that every rank needs (ES fitness, PLR regret scores): ``all_gather_device``;
  * small host-side metadata of the level sampler (terminated flags, buffer ids): ``all_gather_host`` over a gloo
    group, so the sampler never synchronises with the CUDA stream just to learn who terminated.
Everything degenerates to the identity when torch.distributed is not initialised."""
from __future__ import annotations

import numpy as np
import torch

_HOST_GROUP = None


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def rank_world():
    d = _dist()
    if d is None:
        return 0, 1
    return d.get_rank(), d.get_world_size()


def local_slice(n_global: int) -> slice:
    r, w = rank_world()
    if n_global % w != 0:
        raise ValueError(f"global batch ({n_global}) must be divisible by the number of ranks ({w})")
    n = n_global // w
    return slice(r * n, (r + 1) * n)


def _host_group():
    """gloo group for host arrays (created collectively on first use; the default group when it already is gloo)."""
    global _HOST_GROUP
    d = _dist()
    if d.get_backend() == "gloo":
        return None
    if _HOST_GROUP is None:
        _HOST_GROUP = d.new_group(backend="gloo")
    return _HOST_GROUP


def all_gather_host(arr: np.ndarray) -> np.ndarray:
    """[n_local, ...] numpy -> [n_global, ...] numpy in rank order; no CUDA work, no stream synchronisation."""
    d = _dist()
    arr = np.ascontiguousarray(arr)
    if d is None or d.get_world_size() == 1:
        return arr
    is_bool = arr.dtype == np.bool_
    t = torch.from_numpy(arr.astype(np.uint8) if is_bool else arr)
    out = [torch.empty_like(t) for _ in range(d.get_world_size())]
    d.all_gather(out, t, group=_host_group())
    res = torch.cat(out, 0).numpy()
    return res.astype(np.bool_) if is_bool else res


def all_gather_device(t: torch.Tensor) -> torch.Tensor:
    """[n_local, ...] tensor -> [n_global, ...] in rank order with one collective on the tensor's device."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return t
    t = t.contiguous()
    out = torch.empty((t.shape[0] * d.get_world_size(),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    d.all_gather_into_tensor(out, t)
    return out


def all_reduce_sum(t: torch.Tensor) -> torch.Tensor:
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.all_reduce(t)
    return t


def shutdown(*graphed_steps, grace_s: float = 30.0):
    """Orderly end of a multi-rank run: release captured CUDA graphs (they hold NCCL nodes; destroying the process group
    underneath them blocks), then destroy the process group.  A watchdog ends the process if the teardown still does
    not return within ``grace_s`` seconds -- all results have been written by then."""
    d = _dist()
    for g in graphed_steps:
        if hasattr(g, "release"):
            g.release()
    if d is None:
        return
    import os
    import sys
    import threading
    torch.cuda.synchronize()
    d.barrier()

    def _bail():
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    t = threading.Timer(grace_s, _bail)
    t.daemon = True
    t.start()
    d.destroy_process_group()
    t.cancel()
