#!/usr/bin/env python
"""Launch-time table of the config-3 RK4 kernel: ms per launch for nsteps in {1, 2, 5, 10, 20, 50, 100}.

    python benchmarks/launch_sweep.py [--label NAME] [--out FILE]        (CRB_LIB=... selects another build)

Each point: `reps` back-to-back launches on the resident ensemble (65,536 x 32, state + stiffness coefficients
168 MB > L2), CUDA events around every launch; reports median / min and the least-squares line
t = intercept + slope * nsteps (the intercept is the per-launch cost that fused steps do not amortise).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--label", default="build")
    ap.add_argument("--out", default="")
    ap.add_argument("--members", type=int, default=65536)
    ap.add_argument("--elements", type=int, default=32)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--nsteps", default="1,2,5,10,20,50,100")
    args = ap.parse_args()
    import torch

    from bench import build_ensemble
    from continuum_robot_b200.integrate import rk4_steps

    dev = torch.device("cuda", 0)
    e, beam, x0 = build_ensemble(0, args.members, args.elements, dev)
    X0 = torch.from_numpy(x0).to(dev)
    X = X0.clone()
    system = beam.make_system(args.members)
    for _ in range(20):  # clocks up
        rk4_steps(beam, X, 0.0, e.h, 50, system=system)
    torch.cuda.synchronize()
    rows = []
    for ns in [int(s) for s in args.nsteps.split(",")]:
        X.copy_(X0)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for _ in range(3):
            rk4_steps(beam, X, 0.0, e.h, ns, system=system)
        for a, b in ev:
            a.record()
            rk4_steps(beam, X, 0.0, e.h, ns, system=system)
            b.record()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in ev])
        rows.append({"nsteps": ns, "median_ms": float(np.median(ms)), "min_ms": float(ms.min()),
                     "element_steps_per_s": args.members * args.elements * ns / (float(np.median(ms)) * 1e-3)})
    ns = np.array([r["nsteps"] for r in rows], dtype=float)
    t = np.array([r["median_ms"] for r in rows])
    slope, intercept = np.polyfit(ns, t, 1)
    out = {"label": args.label, "lib": os.environ.get("CRB_LIB", "in-tree"), "members": args.members,
           "elements": args.elements, "reps": args.reps, "rows": rows,
           "fit_ms": {"intercept": float(intercept), "slope_per_step": float(slope)},
           "finite": bool(torch.isfinite(X).all().item())}
    s = json.dumps(out)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
