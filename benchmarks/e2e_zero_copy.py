#!/usr/bin/env python
"""Feasibility / measurement: the persistent RK4 kernel working DIRECTLY on the pinned host buffer (its bulk tile
copies then cross PCIe themselves, tile by tile) against the chunked copy pipeline.
    python benchmarks/e2e_zero_copy.py [--nsteps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsteps", type=int, default=20)
    a = ap.parse_args()
    import torch

    from bench import build_ensemble
    from continuum_robot_b200.integrate import HostPipeline, rk4_steps

    dev = torch.device("cuda", 0)
    e, beam, x0 = build_ensemble(0, 65536, 32, dev)
    system = beam.make_system(65536)
    x_host = torch.from_numpy(x0).pin_memory()
    X = torch.from_numpy(x0).to(dev)
    rk4_steps(beam, X, 0.0, e.h, a.nsteps, system=system)  # device-resident result
    rk4_steps(beam, x_host, 0.0, e.h, a.nsteps, system=system)  # the kernel reads / writes the pinned host rows
    torch.cuda.synchronize()
    same = bool(torch.equal(X.cpu(), x_host))
    out = {"nsteps": a.nsteps, "bitwise_equal_to_resident": same}
    ms = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rk4_steps(beam, x_host, 0.0, e.h, a.nsteps, system=system)
        e1.record()
        torch.cuda.synchronize()
        ms.append(round(e0.elapsed_time(e1), 3))
    out["zero_copy_ms"] = ms
    pipe = HostPipeline(beam, 65536)
    pm = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.run(x_host, 0.0, e.h, a.nsteps)
        pipe.wait()
        e1.record()
        pipe.synchronize()
        torch.cuda.synchronize()
        pm.append(round(e0.elapsed_time(e1), 3))
    out["pipeline_ms"] = pm
    out["element_steps_per_s_zero_copy"] = 65536 * 32 * a.nsteps / (min(ms) * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
