#!/usr/bin/env python
"""Print the measured parity errors (block inf-norm) of the CUDA path against the reference's
golden trajectories: how much margin there is under the 1e-9 gate."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import block_err, load, make_gpu_beam, params_array  # noqa: E402


def main():
    from continuum_robot_b200 import FullStateLinear, TipImpulse, solve_ensemble

    out = {}
    g = load("cfg12.npz")
    for cfg in ("cfg1", "cfg2"):
        p = cfg + "/"
        beam = make_gpu_beam(params_array(g, p), g[p + "elem_type"], g[p + "bc"], 1000.0 if cfg == "cfg2" else 0.0, cfg == "cfg1")
        n = beam.n_free
        res = solve_ensemble(beam, (0.0, 0.1), torch.zeros(1, 2 * n, dtype=torch.float64, device="cuda"), method="RK4", h=2.5e-5,
                             save_every=40, u=TipImpulse(torch.tensor([0.1], dtype=torch.float64, device="cuda")))
        got, ref = res.y[0].T.cpu().numpy()[1:], g[p + "Y"]
        out[cfg + " (4000 RK4 steps)"] = max(block_err(got[k], ref[k], n) for k in range(len(ref)))
    g = load("cfg3_samples.npz")
    B, N = g["E_parsed"].shape
    par = np.zeros((B, N, 7))
    for k, c in enumerate(("length", "moment_inertia", "density", "cross_area")):
        par[:, :, (0, 2, 3, 4)[k]] = g[c][None, :]
    par[:, :, 1] = g["E_parsed"]
    par[:, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    n = beam.n_free
    X0 = torch.from_numpy(np.concatenate([g["q0"], g["v0"]], axis=1)).cuda()
    res = solve_ensemble(beam, (0.0, 1000 * float(g["h"])), X0, method="RK4", h=float(g["h"]), save_every=250)
    got = res.y.permute(0, 2, 1).cpu().numpy()[:, 1:]
    out["cfg3 samples (1000 RK4 steps, paired fast kernel)"] = max(block_err(got[i, k], g["Y"][i, k], n) for i in range(B) for k in range(g["Y"].shape[1]))
    g = load("cfg5_samples.npz")
    beam = make_gpu_beam(params_array(g)[None], g["elem_type"], g["bc"], 0.0, True)
    n = beam.n_free
    B = len(g["amp"])
    for general in (False, True):
        beam.force_general_kernels = general
        res = solve_ensemble(beam, (0.0, 2000 * float(g["h"])), torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda"), method="RK4",
                             h=float(g["h"]), save_every=500, u=TipImpulse(torch.from_numpy(g["amp"]).cuda()),
                             controller=FullStateLinear(torch.from_numpy(g["gain"]).cuda()))
        got = res.y.permute(0, 2, 1).cpu().numpy()[:, 1:]
        out["cfg5 samples (2000 RK4 steps, %s)" % ("banded kernel" if general else "shared-operator kernel")] = max(
            block_err(got[i, k], g["Y"][i, k], n) for i in range(B) for k in range(g["Y"].shape[1]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
