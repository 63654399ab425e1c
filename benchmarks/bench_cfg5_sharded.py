#!/usr/bin/env python
"""BASELINE config 5 as specified: the LQR closed-loop rollout ensemble (examples/lqr_control.py) of 1,048,576
members sharded across the GPUs of one box by member, 2,000 RK4 steps, no collective on the step path, one final
gather of the tip displacements.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
        benchmarks/bench_cfg5_sharded.py [--members 1048576] [--steps 2000] [--designs]

`--designs`: every member has its OWN design (random E, density) and gain, synthesised on its rank's GPU
(crb_lqr_gains) -- the design-ensemble variant; default: one shared design and gain (the example's).
Prints one JSON line on rank 0: member-steps/s over all ranks (device-timed, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from continuum_robot_b200 import (BatchedDynamicEulerBernoulliBeam, BatchedLinearQuadraticRegulator, ForceParams,
                                      FullStateLinear, LinearQuadraticRegulator, TipImpulse)
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from continuum_robot_b200.outputs import tip_displacement
    from continuum_robot_b200.sharding import gather_members, shard_range

    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=1048576)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--designs", action="store_true")
    a = ap.parse_args()
    out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # library banners (NCCL) go to stderr
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    e = ens.config5(a.members)  # seed 555: disturbance amplitudes U(1, 20) N on the tip, x0 = 0
    lo, hi = shard_range(a.members, rank, world)
    B = hi - lo
    m = ens.material()
    N = 6
    par = np.zeros((B if a.designs else 1, N, 7))
    par[:, :, 0], par[:, :, 1], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["E"], m["I"], m["rho"], m["A"]
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    if a.designs:
        rng = np.random.default_rng(1000 + rank)
        par[:, :, 1] *= np.exp(0.15 * rng.standard_normal((B, 1)))
        par[:, :, 3] *= np.exp(0.05 * rng.standard_normal((B, 1)))  # (h = 5e-6 is 58 % of the base design's RK4 stability limit)
    beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * N}, ForceParams(enable_gravity_effects=True),
                                            device=dev)
    beam.create_system_func()
    beam.create_input_func()
    n = beam.n_free
    Qh = np.diag(np.r_[100.0 * np.ones(n), 10.0 * np.ones(n)])  # examples/lqr_control.py:61-66
    synth_ms = None
    if a.designs:
        Md, Kd = beam.dense_matrices()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        gain = BatchedLinearQuadraticRegulator(Kd, Md, torch.from_numpy(Qh).to(dev), torch.eye(n, dtype=torch.float64, device=dev)).compute_gain_matrix()
        t1e.record()
        torch.cuda.synchronize()
        synth_ms = t0e.elapsed_time(t1e)
    else:
        gain = torch.from_numpy(LinearQuadraticRegulator(beam.beam_model.get_stiffness_matrix(), beam.beam_model.get_mass_matrix(),
                                                         Qh, np.eye(n)).compute_gain_matrix()).to(dev)
    ctrl = FullStateLinear(gain)
    imp = TipImpulse(torch.from_numpy(e.impulse_amp[lo:hi]).to(dev))
    X = torch.zeros(B, 2 * n, dtype=torch.float64, device=dev)
    per_launch = 200
    rk4_steps(beam, X, 0.0, e.h, per_launch, u=imp, controller=ctrl)  # warm-up (also builds the cached operators)
    gather_members(X[:, n - 2].contiguous(), a.members)  # warm-up: sets up the point-to-point channels of the gather
    X.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(a.steps // per_launch):
        rk4_steps(beam, X, k * per_launch * e.h, e.h, per_launch, u=imp, controller=ctrl)
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    tips = gather_members(tip_displacement(X.unsqueeze(-1))[:, 0].contiguous(), a.members)  # the only communication
    g1.record()
    torch.cuda.synchronize()
    if rank == 0:
        steps = a.steps // per_launch * per_launch
        print(json.dumps({
            "config": "cfg5: LQR closed-loop rollout, %d members x 6 elements sharded over %d GPU(s) by member, %s" % (
                a.members, world, "one design and gain PER MEMBER (synthesised on the GPU)" if a.designs else "one shared design and gain"),
            "n_gpus": world, "steps": steps, "h": e.h, "ms": float(ms.item()),
            "member_steps_per_s": a.members * steps / (float(ms.item()) * 1e-3),
            "final_gather_ms": g0.elapsed_time(g1), "gathered": list(tips.shape), "finite": bool(torch.isfinite(tips).all()),
            "tip_abs_max_m": float(tips.abs().max()), "lqr_synthesis_ms_rank0": synth_ms}), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
