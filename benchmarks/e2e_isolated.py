#!/usr/bin/env python
"""One isolated HostPipeline.run call of config 3 (the driver's `--steps 20` e2e leg), repeated, with and without
an idle gap before it:  python benchmarks/e2e_isolated.py [--nsteps 20]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsteps", type=int, default=20)
    ap.add_argument("--chunk", type=int, default=0)
    a = ap.parse_args()
    import torch

    from bench import build_ensemble
    from continuum_robot_b200.integrate import HostPipeline

    dev = torch.device("cuda", 0)
    e, beam, x0 = build_ensemble(0, 65536, 32, dev)
    x_host = torch.from_numpy(x0).pin_memory()
    pipe = HostPipeline(beam, 65536, chunk_members=a.chunk)
    out = {"chunk_members": pipe.chunk_members, "nsteps": a.nsteps}

    def one():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipe.run(x_host, 0.0, e.h, a.nsteps)
        pipe.wait()
        e1.record()
        pipe.synchronize()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(3):
        one()
    out["back_to_back_ms"] = [round(one(), 3) for _ in range(6)]
    gaps = {}
    for gap in (0.001, 0.01, 0.05, 0.2):
        r = []
        for _ in range(4):
            time.sleep(gap)
            r.append(round(one(), 3))
        gaps[str(gap)] = r
    out["after_idle_gap_s"] = gaps
    # plain copies of the same buffer, sequential and concurrent (what the copies alone cost inside one call)
    X = pipe.X
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def copies(concurrent):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1):
            X.copy_(x_host, non_blocking=True)
        if not concurrent:
            s2.wait_stream(s1)
        with torch.cuda.stream(s2):
            x_host.copy_(X, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1), 3)

    out["plain_h2d_then_d2h_ms"] = [copies(False) for _ in range(3)]
    out["plain_h2d_and_d2h_concurrent_ms"] = [copies(True) for _ in range(3)]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
