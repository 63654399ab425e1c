#!/usr/bin/env python
"""Probe of the host-buffer (e2e) path: PCIe copy rates alone / concurrent, NUMA binding, and the
HostPipeline chunking.  Prints JSON lines.  Not part of the driver's bench."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from continuum_robot_b200.sharding import bind_to_gpu_numa_node

    bound = bind_to_gpu_numa_node(0) if os.environ.get("PROBE_BIND", "0") == "1" else False
    dev = torch.device("cuda", 0)
    nbytes = 65536 * 192 * 8
    h_in = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    h_out = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    h_in.normal_()
    d_a = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    d_b = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        for s in (s1, s2):
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def h2d():
        d_a.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_b, non_blocking=True)

    def both():
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    out = {"numa_bound": bound}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
        ms = timed(fn)
        out[name + "_ms"] = ms
        out[name + "_GBps"] = nbytes / ms / 1e6
    print(json.dumps(out), flush=True)

    # HostPipeline chunk sweep on the bench workload
    sys.argv = ["bench.py"]
    import bench
    from continuum_robot_b200.integrate import HostPipeline

    B, N, S = 65536, 32, 50
    e, beam, x0 = bench.build_ensemble(0, B, N, dev)
    x_host = torch.from_numpy(x0).pin_memory()
    for chunk in [int(c) for c in os.environ.get("PROBE_CHUNKS", "0,2368,4736,9472,4096,8192").split(",")]:
        pipe = HostPipeline(beam, B, chunk_members=chunk)
        for _ in range(3):
            pipe.run(x_host, 0.0, e.h, S)
        pipe.synchronize()
        x_host.copy_(torch.from_numpy(x0))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        reps = 40
        for k in range(reps):
            pipe.run(x_host, k * S * e.h, e.h, S)
        pipe.wait()
        b.record()
        torch.cuda.synchronize()
        pipe.synchronize()
        ms = a.elapsed_time(b) / reps
        print(json.dumps({"chunk_members": pipe.chunk_members, "ms_per_run": ms,
                          "element_steps_per_s": B * N * S / ms * 1e3, "numa_bound": bound}), flush=True)


if __name__ == "__main__":
    main()
