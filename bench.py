#!/usr/bin/env python
"""Benchmark of the LPG meta-training hot path (BASELINE.json metric: gridworld agent env-steps/s
including the LPG update, and meta-steps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full meta-step of the reference's train loop (train.py:32-52):
lpg_meta_grad_train_step (K agent updates on 64 workers x 20 steps, eval rollout, meta-loss, the
meta-gradient, Adam) + level_sampler.sample, on the BASELINE configs[1] workload: env_mode
all_shortlife, 512 agents per GPU.  Prints ONE JSON line (rank 0).

--impl reference times the CPU restatement of the reference (oracle/, torch CPU + numpy; the JAX
reference itself cannot be installed in this image) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

AGENTS_PER_GPU = 512
ENV_MODE = "all_shortlife"
CPU_SAMPLE_AGENTS = 8


def env_steps_per_agent(args, max_rollout_len):
    """K*W*L (train) + W*L (eval rollout) + 4*cap (eval_agent), SURVEY.md §8(d)."""
    W, L, K = args.env_workers, args.train_rollout_len, args.num_agent_updates
    return K * W * L + W * L + 4 * max_rollout_len


# ------------------------------------------------------------------------------------------------
def cpu_reference_step(n_agents, seed=0, threads=None):
    """One meta-step of the CPU oracle on n_agents all_shortlife agents; returns seconds."""
    import numpy as np
    import torch
    from oracle import prng, configs
    from oracle.gridworld import GridWorld
    from oracle.rollout import RolloutWrapper
    from oracle.lpg import LPGLayout, init_lpg_params
    from oracle.agents import AgentTables
    from oracle.meta import lpg_meta_grad_train_step, Adam
    if threads:
        torch.set_num_threads(threads)
    kw, ep = configs.get_env_spec(ENV_MODE)
    env = GridWorld(**kw)
    ro = RolloutWrapper(env, 20, ep)
    keys = prng.split(prng.PRNGKey(seed), n_agents)
    p, life = configs.reset_env_params(keys, ENV_MODE)
    D = env.obs_dim
    rs = np.random.RandomState(seed)
    f = lambda c: torch.tensor((rs.randn(n_agents, D, c) / np.sqrt(D)).astype(np.float32))
    ag = AgentTables(f(5), f(8), torch.zeros(n_agents, dtype=torch.long))
    value = f(1)
    lay = LPGLayout()
    flat = torch.tensor(init_lpg_params(lay, seed))
    adam = Adam(lay.size, 1e-4)
    s0 = ro.batch_reset(None, p, 64)
    t0 = time.perf_counter()
    out = lpg_meta_grad_train_step(prng.PRNGKey(seed + 1), lay, flat, ag, value, ro, p, s0, life)
    adam.step(flat, out["grad"])
    return time.perf_counter() - t0, ep


def run_reference(a):
    """Reference arm: the CPU restatement on the host cores (rank 0 only)."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from to_ued_b200.experiments.parse_args import parse_args
    args = parse_args(["--env_mode", ENV_MODE])
    torch.set_num_threads(os.cpu_count() or 1)          # torchrun pins OMP_NUM_THREADS=1; use every host core
    cores = torch.get_num_threads()
    for _ in range(a.warmup):
        cpu_reference_step(2)
    times = []
    for i in range(a.steps):
        dt, ep = cpu_reference_step(CPU_SAMPLE_AGENTS, seed=i)
        times.append(dt)
    per_step = sum(times) / len(times)
    spa = env_steps_per_agent(args, ep)
    val = CPU_SAMPLE_AGENTS * spa / per_step
    sample = f"{CPU_SAMPLE_AGENTS} of {AGENTS_PER_GPU} agents, one full meta-step each step (oracle: torch CPU + numpy)"
    line = {
        "impl": "reference", "metric": "gridworld agent env-steps/sec incl. LPG update", "value": val,
        "unit": "env-steps/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LPG meta-gradient step, env_mode={ENV_MODE}, CPU restatement of the reference (not JAX)",
                   "agents": CPU_SAMPLE_AGENTS, "env_workers": 64, "train_rollout_len": 20, "num_agent_updates": 5},
        "meta_steps_per_s": 1.0 / per_step * CPU_SAMPLE_AGENTS / AGENTS_PER_GPU,
        "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import statistics
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from to_ued_b200 import _lib
    from to_ued_b200.util import prng
    from to_ued_b200.experiments.parse_args import parse_args
    from to_ued_b200.environments.level_sampler import LevelSampler
    from to_ued_b200.meta.meta import create_lpg_train_state, make_lpg_train_step

    import to_ued_b200
    precision = to_ued_b200.GRU_PRECISION
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_global = AGENTS_PER_GPU * world
    mini_batches = int(os.environ.get("TOUED_BENCH_MINI_BATCHES", "2"))
    args = parse_args(["--env_mode", ENV_MODE, "--num_agents", str(n_global), "--num_mini_batches", str(mini_batches)])
    rng = prng.PRNGKey(args.seed)
    rng, lpg_rng, buffer_rng = prng.split(rng, 3)
    train_state = create_lpg_train_state(lpg_rng, args)
    sampler = LevelSampler(args)
    buf = sampler.initialize_buffer(buffer_rng)
    rng, _rng = prng.split(rng, 2)
    # each rank builds only its own agents: keys of the global batch, local slice
    import train as train_mod
    buf, agents, vcs = sampler.initial_sample(_rng, buf, n_global, True) if world == 1 else \
        _initial_sample_sharded(sampler, _rng, buf, n_global, rank, AGENTS_PER_GPU)
    step_fn = make_lpg_train_step(args, sampler)
    spa = env_steps_per_agent(args, sampler.max_rollout_len)

    def one_step(rng, train_state, agents, vcs, buf):
        rng, _rng = prng.split(rng, 2)
        train_state, agents, vcs, metrics = step_fn(rng=_rng, lpg_train_state=train_state, agent_states=agents,
                                                    value_critic_states=vcs)
        rng, _rng = prng.split(rng, 2)
        buf, agents, vcs = sampler.sample(_rng, buf, agents, vcs)
        return rng, train_state, agents, vcs, buf, metrics

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = (rng, train_state, agents, vcs, buf)
    for _ in range(a.warmup):
        *state, metrics = one_step(*state)
    barrier()

    # ---- timed region 1: device-timed throughput, inputs resident, no host reads of results ----
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    _lib.reset_counters(profile=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        *state, metrics = one_step(*state)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = _lib.kernel_launches()

    # ---- per-kernel CUDA-event times for the roofline / shares: same steps, but the mini-batches run one
    #      after the other on one stream so that every launch is timed without a concurrent neighbour ----
    def one_step_serial(rng, train_state, agents, vcs, buf):
        rng, _rng = prng.split(rng, 2)
        train_state, agents, vcs, metrics = step_fn(rng=_rng, lpg_train_state=train_state, agent_states=agents,
                                                    value_critic_states=vcs, num_streams=1)
        rng, _rng = prng.split(rng, 2)
        buf, agents, vcs = sampler.sample(_rng, buf, agents, vcs)
        return rng, train_state, agents, vcs, buf, metrics
    PROF_STEPS = 2
    to_ued_b200.SIDE_STREAMS = False                  # no side streams either: every launch is timed alone
    *state, metrics = one_step_serial(*state)
    barrier()
    _lib.reset_counters(profile=True)
    for _ in range(PROF_STEPS):
        *state, metrics = one_step_serial(*state)
    barrier()
    prof = _lib.profile_ms()
    _lib.reset_counters(profile=False)
    to_ued_b200.SIDE_STREAMS = True

    # ---- timed region 2: end to end through the public API, host key in / metrics out each step ----
    barrier()
    _lib.reset_counters(profile=False)
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(a.steps):
        *state, metrics = one_step(*state)
        flat = [v for k, v in metrics.items() if not isinstance(v, dict)] + \
               [vv for v in metrics.values() if isinstance(v, dict) for vv in v.values()]
        host = torch.stack([f.reshape(()) for f in flat]).cpu()      # ONE D2H read of the step's 9 result scalars
        d2h = host.numel() * host.element_size()
    barrier()
    e2e_s = time.perf_counter() - t0
    if clocks:
        clocks.stop_flag = True
        clocks.join(timeout=2)

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    ms_per_step = dev_ms / a.steps
    total_env_steps = n_global * spa
    value = total_env_steps / (ms_per_step * 1e-3)
    e2e_val = total_env_steps / (e2e_ms / a.steps * 1e-3)
    K, W, L = args.num_agent_updates, args.env_workers, args.train_rollout_len
    # host->device bytes per step, counted where the package copies: the 8-byte step key, the lifetime mask
    # and the records / init keys of the levels the sampler replaced
    h2d = _lib.H2D_BYTES[0] / a.steps

    # ---- roofline of the dominant kernel (live CUDA-event times of the timed region) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    per_launch_agents = AGENTS_PER_GPU // mini_batches
    tokens = per_launch_agents * W * L
    flops = {   # algorithmic FLOPs per launch (2*256*768 per token for each GRU contraction + heads)
        "toued_gru_forward": tokens * 400896.0,
        "toued_gru_backward": per_launch_agents * W * (L - 1) * 2.0 * 768 * 256 + tokens * 2.0 * 256 * 9,
        "toued_lpg_wgrad": tokens * (2.0 * 256 * 768 + 2.0 * 8 * 768 + 2.0 * 256 * 9),
        "toued_gru_forward_tc": tokens * 400896.0,
        "toued_gru_backward_tc": per_launch_agents * W * (L - 1) * 2.0 * 768 * 256 + tokens * 2.0 * 256 * 9,
        "toued_lpg_wgrad_tc": tokens * (2.0 * 256 * 768 + 2.0 * 8 * 768 + 2.0 * 256 * 9),
    }
    # algorithmic HBM bytes per token the tensor-core kernels must move (fp16 gate planes + h,
    # bf16 dG / h' images) -- reported next to the tensor roofline because the activations do not fit on chip
    hbm_bytes = {"toued_gru_forward_tc": tokens * (32.0 + 4 * 512 + 512 + 512 + 36),
                 "toued_gru_backward_tc": tokens * (4 * 512 + 512 + 72 + 2048 + 40.0),
                 "toued_lpg_wgrad_tc": tokens * (2048 + 512 + 512 + 128 + 36.0)}
    total_prof = sum(ms for _, ms in prof.values()) or 1.0
    shares = {k: {"calls": c, "ms_per_step": ms / PROF_STEPS, "share": ms / total_prof} for k, (c, ms) in prof.items()}
    dom = max((k for k in prof if k in flops), key=lambda k: prof[k][1])
    calls, ms = prof[dom]
    achieved = flops[dom] / (ms / calls * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/), scaled to this
    # launch size (the capture records bytes per token)
    traffic = None
    try:
        ncu = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_summary.json")))
        traffic = ncu["kernels"][dom]["dram_bytes_per_token"] * tokens
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "agents_per_launch": per_launch_agents,
                "ms_per_launch": ms / calls,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback",
                "note": ("tcgen05 path (fp16/bf16 operands, fp32 accumulate in TMEM)" if precision == "tc" else
                         "exact-fp32 SIMT GRU path") + "; FLOPs = algorithmic GRU/head/weight-gradient FLOPs per launch"}
    if dom in hbm_bytes:
        gbs = hbm_bytes[dom] / (ms / calls * 1e-3) / 1e9
        roofline["hbm"] = {"achieved": gbs, "peak": peaks.get("hbm_gbs", 6556.8), "unit": "GB/s",
                           "frac": gbs / peaks.get("hbm_gbs", 6556.8),
                           "note": "algorithmic activation bytes of the same kernel (saved gates / dG image round trips)"}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) ----
    cpu = None
    if world == 1:
        dt, _ = cpu_reference_step(CPU_SAMPLE_AGENTS)
        cpu = {"value": CPU_SAMPLE_AGENTS * spa / dt, "unit": "env-steps/s", "cores": torch.get_num_threads(),
               "kind": "port",
               "sample": f"one meta-step of {CPU_SAMPLE_AGENTS} of the {AGENTS_PER_GPU} agents ({dt:.1f} s); "
                         "CPU restatement of the reference (torch CPU + numpy), not JAX"}
    line = {
        "metric": "gridworld agent env-steps/sec incl. LPG update", "value": value, "unit": "env-steps/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16/bf16 operands, f32 accumulate (GRU); f32 elsewhere" if precision == "tc" else "f32", "data": "synthetic",
        "config": {"workload": f"LPG meta-gradient step (train.py loop body), env_mode={ENV_MODE}, "
                               f"{AGENTS_PER_GPU} agents/GPU x {W} workers x {L} steps x K={K} updates, "
                               f"num_mini_batches={mini_batches} run concurrently on CUDA streams (the reference README uses 16 "
                               "sequential mini-batches as a memory device; results identical)",
                   "gru_precision": precision, "agents_per_gpu": AGENTS_PER_GPU, "global_agents": n_global, "env_steps_per_meta_step": total_env_steps,
                   "l2_policy": "working set per step (>15 GB of activations) exceeds L2; no explicit flush"},
        "meta_steps_per_s": 1e3 / ms_per_step,
        "e2e": {"value": e2e_val, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / a.steps},
        "gpu_launches": launches, "roofline": roofline, "kernel_shares": shares,
        "clocks": clocks.summary() if clocks else None,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _initial_sample_sharded(sampler, rng, buf, n_global, rank, n_local):
    """initial_sample for the global batch restricted to this rank's slice of agents."""
    import train as train_mod
    buf, agents, vcs = sampler.initial_sample(rng, buf, n_global, True)
    agents, vcs = train_mod._shard(agents, vcs, rank, n_local)
    return buf, agents, vcs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
