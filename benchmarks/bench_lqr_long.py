#!/usr/bin/env python
"""LQR synthesis throughput vs beam length (designs/s): python benchmarks/bench_lqr_long.py [--out FILE]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--sizes", default="6:8192,10:2048,14:1024,16:592,24:296,32:296")
    a = ap.parse_args()
    import torch

    from continuum_robot_b200 import BatchedDynamicEulerBernoulliBeam, BatchedLinearQuadraticRegulator, ForceParams
    from continuum_robot_b200 import ensembles as ens

    dev = "cuda"
    rows = []
    for item in a.sizes.split(","):
        N, B = (int(x) for x in item.split(":"))
        rng = np.random.default_rng(6)
        m = ens.material()
        par = np.empty((B, N, 7))
        par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
        par[:, :, 1] = 75e9 * np.exp(0.3 * rng.standard_normal((B, 1)))
        par[:, :, 3] *= np.exp(0.2 * rng.standard_normal((B, 1)))
        par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
        beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * N}, ForceParams(enable_gravity_effects=True))
        beam.create_system_func(); beam.create_input_func()
        n = beam.n_free
        Q = torch.diag(torch.cat([torch.full((n,), 100.0), torch.full((n,), 10.0)])).to(dev, torch.float64)
        R = torch.eye(n, dtype=torch.float64, device=dev)
        Md, Kd = beam.dense_matrices()
        best = 1e30
        for _ in range(2):
            lqr = BatchedLinearQuadraticRegulator(Kd, Md, Q, R)  # (a regulator caches its gains)
            torch.cuda.synchronize()
            a0, b0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            K = lqr.compute_gain_matrix()
            b0.record()
            torch.cuda.synchronize()
            best = min(best, a0.elapsed_time(b0))
        rows.append({"elements": N, "n": n, "hamiltonian": 4 * n, "designs": B, "ms": best, "designs_per_s": B / (best * 1e-3),
                     "finite": bool(torch.isfinite(K).all().item())})
        print(json.dumps(rows[-1]), flush=True)
    if a.out:
        json.dump({"what": "crb_lqr_gains, 1 refinement pass, one B200", "rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
