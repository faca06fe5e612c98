"""Run logging and checkpoints (reference experiments/logging.py:11-46), without WandB.

The reference logs every step's metrics to WandB and saves two flax checkpoints (the LPG ``TrainState`` and the
level buffer, ``prefix="buffer_"``) into the run directory; it has **no restore** (README: "coming soon").  Here
``init_logger`` creates a local run directory, ``log_results`` writes ``metrics.jsonl`` plus the two checkpoints
as ``.npz`` archives (plain arrays, no pickle), and ``restore_checkpoint`` / ``restore_buffer`` read them back —
onto whatever device the target state lives on.  File names follow flax's ``<prefix><step>`` convention."""
from __future__ import annotations

import json
import os
import time
from dataclasses import fields

import numpy as np
import torch

CKPT_DIR = "checkpoints"
_RUN = {"dir": None}


def init_logger(args):
    """logging.py:11-22.  Creates ``<log_dir>/<group>-<timestamp>/checkpoints`` (``--wandb_group`` names the
    group as in the reference; ``TOUED_LOG_DIR`` overrides the root, default ``runs``)."""
    root = os.environ.get("TOUED_LOG_DIR", "runs")
    run = os.path.join(root, f"{getattr(args, 'wandb_group', None) or 'debug'}-{time.strftime('%Y%m%d-%H%M%S')}-{os.getpid()}")
    os.makedirs(os.path.join(run, CKPT_DIR), exist_ok=True)
    with open(os.path.join(run, "config.json"), "w") as f:
        json.dump({k: v for k, v in vars(args).items() if isinstance(v, (int, float, str, bool, type(None)))}, f, indent=1)
    _RUN["dir"] = run
    return run


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def to_host(m):
    """Metrics tree -> plain Python (floats / lists), dropping private ``_``-keys; one value at a time (this is the
    logging path, not the step)."""
    if isinstance(m, dict):
        return {k: to_host(v) for k, v in m.items() if not str(k).startswith("_")}
    if isinstance(m, torch.Tensor):
        m = m.detach().cpu().numpy()
    if isinstance(m, np.ndarray):
        return float(m) if m.size == 1 else m.astype(float).tolist()
    if isinstance(m, (np.floating, np.integer)):
        return m.item()
    return m


_floats = to_host


def _train_state_arrays(train_state) -> dict:
    """Flat array dict of an LPGTrainState, or of an ESTrainState (which wraps one, util/data.py:62-68)."""
    out = {}
    inner = getattr(train_state, "train_state", None)
    if inner is not None:                                   # ESTrainState
        for k, v in train_state.es_state.items():
            out[f"es_state/{k}"] = _np(v)
        train_state = inner
    out["params"] = _np(train_state.params)
    out["step"] = np.asarray(int(train_state.step))
    for k, v in train_state.opt_state.items():
        if not k.endswith("_dev"):                          # device mirror of `count` kept by the graphed step
            out[f"opt_state/{k}"] = _np(v)
    return out


def save_checkpoint(ckpt_dir: str, target, step: int, prefix: str = "checkpoint_") -> str:
    """flax.training.checkpoints.save_checkpoint(keep=1) for the two targets the reference saves."""
    os.makedirs(ckpt_dir, exist_ok=True)
    arrays = _buffer_arrays(target) if hasattr(target, "score") and hasattr(target, "level") else _train_state_arrays(target)
    path = os.path.join(ckpt_dir, f"{prefix}{int(step)}.npz")
    tmp = path + ".tmp.npz"
    np.savez(tmp, **arrays)
    os.replace(tmp, path)
    for f in os.listdir(ckpt_dir):                          # keep=1
        if f.startswith(prefix) and f.endswith(".npz") and os.path.join(ckpt_dir, f) != path:
            os.remove(os.path.join(ckpt_dir, f))
    return path


def log_results(args, metrics, train_state, level_buffer):
    """logging.py:25-46: per-step metrics, then the train-state and level-buffer checkpoints."""
    run = _RUN["dir"] or init_logger(args)
    with open(os.path.join(run, "metrics.jsonl"), "w") as f:
        for step, m in enumerate(metrics):
            f.write(json.dumps({"step": step, **_floats(m)}) + "\n")
    paths = [save_checkpoint(os.path.join(run, CKPT_DIR), train_state, args.train_steps)]
    if level_buffer is not None:
        paths.append(save_checkpoint(os.path.join(run, CKPT_DIR), level_buffer, args.train_steps, prefix="buffer_"))
    return paths


def _assign(dst, src):
    if isinstance(dst, torch.Tensor):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(src)).to(dst.dtype))
        return dst
    return type(dst)(src) if isinstance(dst, (int, float)) else np.asarray(src)


def restore_checkpoint(path: str, train_state):
    """Load a ``checkpoint_<step>.npz`` into ``train_state`` (same model / optimizer as when it was saved).
    Tensors are overwritten in place on their device; returns the restored state."""
    z = np.load(path)
    inner = getattr(train_state, "train_state", None)
    if inner is not None:
        es = dict(train_state.es_state)
        for k in es:
            es[k] = _assign(es[k], z[f"es_state/{k}"])
        return train_state.replace(train_state=restore_checkpoint(path, inner), es_state=es)
    if z["params"].shape != tuple(train_state.params.shape):
        raise ValueError(f"checkpoint holds {z['params'].shape[0]} LPG parameters, the model has {train_state.params.numel()}")
    _assign(train_state.params, z["params"])
    opt = dict(train_state.opt_state)
    for k in opt:
        opt[k] = _assign(opt[k], z[f"opt_state/{k}"]) if isinstance(opt[k], torch.Tensor) else int(z[f"opt_state/{k}"])
    return train_state.replace(opt_state=opt, step=int(z["step"]))


def _buffer_arrays(buf) -> dict:
    out = {"score": _np(buf.score), "active": _np(buf.active), "new": _np(buf.new),
           "level/lifetime": _np(buf.level.lifetime), "level/buffer_id": _np(buf.level.buffer_id)}
    if buf.level.env_params is None:                      # device-resident records (csrc/levelgen.cu): the packed LevelRec bytes
        out["level/packed"] = _np(buf.level.packed)
    else:
        for f in fields(buf.level.env_params):
            out[f"level/env_params/{f.name}"] = _np(getattr(buf.level.env_params, f.name))
    return out


def restore_buffer(path: str):
    """Load a ``buffer_<step>.npz`` into a LevelBuffer (level_sampler.py:29-54)."""
    from ..environments.gridworld.gridworld import EnvParams
    from ..environments.level_sampler import LevelBuffer
    from ..util.data import Level
    z = np.load(path)
    if "level/packed" in z.files:
        import torch
        packed = torch.from_numpy(z["level/packed"])
        packed = packed.cuda() if torch.cuda.is_available() else packed
        return LevelBuffer(Level(None, z["level/lifetime"], z["level/buffer_id"], packed), z["score"], z["active"], z["new"])
    params = EnvParams(**{f.name: z[f"level/env_params/{f.name}"] for f in fields(EnvParams)})
    return LevelBuffer(Level(params, z["level/lifetime"], z["level/buffer_id"]), z["score"], z["active"], z["new"])
