#!/usr/bin/env python
"""One crb_lqr_gains launch (for ncu): 1480 designs x 6 elements = two full waves of 5 blocks per SM."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from continuum_robot_b200 import BatchedDynamicEulerBernoulliBeam, BatchedLinearQuadraticRegulator
from continuum_robot_b200 import ensembles as ens

B, N = int(os.environ.get("LQR_B", 1480)), int(os.environ.get("LQR_N", 6))
rng = np.random.default_rng(6)
m = ens.material()
par = np.zeros((B, N, 7))
par[:, :, 0], par[:, :, 2], par[:, :, 4] = m["length"], m["I"], m["A"]
par[:, :, 1] = 75e9 * np.exp(0.3 * rng.standard_normal((B, 1)))
par[:, :, 3] = m["rho"] * np.exp(0.2 * rng.standard_normal((B, 1)))
par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * N})
n = beam.n_free
Q = torch.diag(torch.cat([torch.full((n,), 100.0), torch.full((n,), 10.0)])).to("cuda", torch.float64)
R = torch.eye(n, dtype=torch.float64, device="cuda")
Md, Kd = beam.dense_matrices()
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    K = BatchedLinearQuadraticRegulator(Kd, Md, Q, R).compute_gain_matrix()
    b.record()
    torch.cuda.synchronize()
    print("designs", B, "ms", a.elapsed_time(b), "finite", bool(torch.isfinite(K).all()))
