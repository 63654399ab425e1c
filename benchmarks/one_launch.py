#!/usr/bin/env python
"""A few launches of the config-3 RK4 kernel for ncu captures:  python benchmarks/one_launch.py [--nsteps 50] [--launches 4]
(CRB_LIB=... selects another build of libcrb.so)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsteps", type=int, default=50)
    ap.add_argument("--launches", type=int, default=4)
    ap.add_argument("--members", type=int, default=65536)
    ap.add_argument("--elements", type=int, default=32)
    ap.add_argument("--save-every", type=int, default=0)
    ap.add_argument("--method", default="rk4", choices=["rk4", "midpoint"])
    a = ap.parse_args()
    import torch

    from bench import build_ensemble
    from continuum_robot_b200.integrate import midpoint_steps, rk4_steps

    dev = torch.device("cuda", 0)
    e, beam, x0 = build_ensemble(0, a.members, a.elements, dev)
    X = torch.from_numpy(x0).to(dev)
    system = beam.make_system(a.members)
    Y = None
    if a.save_every:
        Y = torch.empty((a.nsteps // a.save_every, a.members, X.shape[1]), dtype=torch.float64, device=dev)
    for _ in range(a.launches):
        if a.method == "rk4":
            rk4_steps(beam, X, 0.0, e.h, a.nsteps, system=system, Y_out=Y, save_every=a.save_every)
        else:
            midpoint_steps(beam, X, 0.0, e.h, a.nsteps, Y_out=Y, save_every=a.save_every)
    torch.cuda.synchronize()
    print("ok", bool(torch.isfinite(X).all().item()))


if __name__ == "__main__":
    main()
