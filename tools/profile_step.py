#!/usr/bin/env python
"""One eagerly enqueued meta-step of the bench workload between cudaProfilerStart / Stop, for ncu
(`--profile-from-start off`):

    python tools/profile_step.py [agents] [serial]      # serial: one stream, no side streams (per-kernel view)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import to_ued_b200  # noqa: E402


def main():
    agents = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    serial = len(sys.argv) > 2 and sys.argv[2] == "serial"
    torch.cuda.set_device(0)
    wl = bench.Workload(bench.DEFAULT_CONFIG, agents, 1, cuda_graph=False)
    kw = {}
    if serial:
        to_ued_b200.SIDE_STREAMS = False
        kw = {"num_streams": 1}
    for _ in range(3):
        wl.one_step(wl.eager_fn, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    wl.one_step(wl.eager_fn, **kw)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one meta-step:", agents, "agents", "serial" if serial else "stream plan")


if __name__ == "__main__":
    main()
