#!/usr/bin/env python
"""The reference's own example scenarios as they ship (a handful of beams, 1 s of simulated time, outputs
every 1 ms: examples/example_utilities.py:116-170, examples/lqr_control.py:87-130), through the drop-in API.
Latency regime (a few members), not the throughput regime of bench.py.  Prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from continuum_robot_b200 import (BatchedDynamicEulerBernoulliBeam, ForceParams, FullStateLinear,
                                      LinearQuadraticRegulator, TipImpulse, solve_ensemble)
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    dev = "cuda"

    def beam_of(N, kind, fp):
        par = np.zeros((1, N, 7))
        par[0, :, 0], par[0, :, 1], par[0, :, 2], par[0, :, 3], par[0, :, 4] = m["length"], m["E"], m["I"], m["rho"], m["A"]
        par[0, :, 5], par[0, :, 6] = m["wetted_area"], m["drag_coef"]
        b = BatchedDynamicEulerBernoulliBeam({"params": par, "type": [kind] * N}, fp)
        b.create_system_func(); b.create_input_func()
        return b

    runs = [
        ("Linear + Gravity, N = 6 (examples/beam_comparison_gravity.py)", 6, "linear", ForceParams(enable_gravity_effects=True), 0.1, None),
        ("Nonlinear + Fluid, N = 6 (examples/beam_comparison_fluid.py)", 6, "nonlinear", ForceParams(fluid_density=1000.0, enable_fluid_effects=True), 0.1, None),
        ("Linear + Gravity, N = 10 (BASELINE config 1)", 10, "linear", ForceParams(enable_gravity_effects=True), 0.1, None),
        ("Nonlinear + Fluid, N = 20 (BASELINE config 2)", 20, "nonlinear", ForceParams(fluid_density=1000.0, enable_fluid_effects=True), 0.1, None),
        ("LQR closed loop, N = 6 (examples/lqr_control.py)", 6, "linear", ForceParams(enable_gravity_effects=True), 10.0, "lqr"),
    ]
    for name, N, kind, fp, amp, ctl in runs:
        beam = beam_of(N, kind, fp)
        n = beam.n_free
        ctrl = None
        h = 2.5e-5
        if ctl == "lqr":
            Q = np.eye(2 * n); Q[:n, :n] *= 100; Q[n:, n:] *= 10
            K = LinearQuadraticRegulator(beam.beam_model.get_stiffness_matrix(), beam.beam_model.get_mass_matrix(), Q, np.eye(n)).compute_gain_matrix()
            ctrl = FullStateLinear(torch.from_numpy(K).to(dev))
            h = 5e-6
        se = int(round(1e-3 / h))
        X0 = torch.zeros(1, 2 * n, dtype=torch.float64, device=dev)
        imp = TipImpulse(torch.tensor([amp], dtype=torch.float64, device=dev))
        for rep in range(2):  # second run: warm
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = solve_ensemble(beam, (0.0, 1.0), X0, method="RK4", h=h, save_every=se, u=imp, controller=ctrl)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        tip = float(res.y[0, n - 2, -1])
        print(json.dumps({"example": name, "steps": int(round(1.0 / h)), "frames": int(res.y.shape[-1]), "wall_s": dt,
                          "tip_w_at_1s": tip, "finite": bool(torch.isfinite(res.y).all())}), flush=True)


if __name__ == "__main__":
    main()
