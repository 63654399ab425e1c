#!/usr/bin/env python
"""Secondary measurements (not the driver's bench line): BASELINE configs 2, 4 and 5 as ensembles.

    python benchmarks/bench_configs.py [--only cfg5] [--members N]

Prints one JSON line per config: element-steps/s (RK4) or element-attempts/s (RK45), CUDA-event
timed, state resident in HBM.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def params(e, B, fluid=True):
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    par = np.empty((B, e.n_elements, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E[:B]
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    return par


_REPS = [3]


def timed(fn, reps=None):
    import torch

    reps = reps or _REPS[0]
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def run(names, slots=0, reps=3):
    """Yield one result dict per measured configuration (names: see the `if name == ...` blocks)."""
    import torch

    from continuum_robot_b200 import (BatchedDynamicEulerBernoulliBeam, ForceParams, FullStateLinear,
                                      LinearQuadraticRegulator, TipImpulse, solve_ensemble)
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps

    class _A:
        pass

    a = _A()
    a.slots = slots
    dev = "cuda"
    _REPS[0] = reps
    for name in names:
        if name == "cfg2":  # nonlinear 20-element + drag, replicated: RK4 h = 2.5e-5
            B, steps = 32768, 200
            e = ens.config2()
            par = np.repeat(params(e, 1), 1, axis=0)
            beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["nonlinear"] * 20},
                                                    ForceParams(fluid_density=1000.0, enable_fluid_effects=True),
                                                    max_slots_per_lane=a.slots)
            beam.create_system_func(); beam.create_input_func()
            X = torch.zeros(B, 120, dtype=torch.float64, device=dev)
            imp = TipImpulse(torch.full((B,), 0.1, dtype=torch.float64, device=dev))
            ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps, u=imp))
            yield {"config": "cfg2 x%d (nonlinear 20 el + drag), RK4" % B, "m": beam._plan.m, "g": beam._plan.g,
                              "element_steps_per_s": B * 20 * steps / (ms * 1e-3), "ms": ms}
        if name == "cfg3p":  # config 3 shape with a PINNED root (constrained DOFs inside an active slot: NC variant of the paired kernel)
            B, steps = 65536, 50
            e = ens.config3(B, 32)
            beam = BatchedDynamicEulerBernoulliBeam({"params": params(e, B), "type": ["linear"] * 32,
                                                     "boundary_condition": [2] + [0] * 31})
            beam.create_system_func(); beam.create_input_func()
            n = beam.n_free
            rng = np.random.default_rng(4)
            X = torch.from_numpy(np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)).to(dev)
            ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps))
            yield {"config": "cfg3 shape with a pinned root (NC variant of the paired kernel)", "m": beam._plan.m, "g": beam._plan.g, "n_free": n,
                              "element_steps_per_s": B * 32 * steps / (ms * 1e-3), "ms": ms}
        if name == "lqr":  # batched LQR synthesis + rollout with one gain per member
            from continuum_robot_b200 import BatchedLinearQuadraticRegulator
            B, steps, N = 8192, 200, 6
            rng = np.random.default_rng(6)
            e = ens.config3(8, N)
            par = np.repeat(params(e, 1), B, axis=0)
            par[:, :, 1] = 75e9 * np.exp(0.3 * rng.standard_normal((B, 1)))
            par[:, :, 3] *= np.exp(0.2 * rng.standard_normal((B, 1)))
            beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * N}, ForceParams(enable_gravity_effects=True))
            beam.create_system_func(); beam.create_input_func()
            n = beam.n_free
            Q = torch.diag(torch.cat([torch.full((n,), 100.0), torch.full((n,), 10.0)])).to(dev, torch.float64)
            R = torch.eye(n, dtype=torch.float64, device=dev)
            Md, Kd = beam.dense_matrices()
            out = {}
            def synth():
                out["K"] = BatchedLinearQuadraticRegulator(Kd, Md, Q, R).compute_gain_matrix()
            ms = timed(synth)
            yield {"config": "LQR synthesis, %d designs x %d elements (n = %d, Hamiltonian %d^2), 1 correction pass" % (B, N, n, 4 * n),
                              "designs_per_s": B / (ms * 1e-3), "ms": ms}
            # rollout: 8 disturbance realisations per design so that the ensemble fills the GPU (65536 members)
            rep = 8
            Bb = B * rep
            big = BatchedDynamicEulerBernoulliBeam({"params": np.tile(par, (rep, 1, 1)), "type": ["linear"] * N},
                                                   ForceParams(enable_gravity_effects=True))
            big.create_system_func(); big.create_input_func()
            ctrl = FullStateLinear(out["K"].repeat(rep, 1, 1))
            X = torch.zeros(Bb, 2 * n, dtype=torch.float64, device=dev)
            imp = TipImpulse(torch.from_numpy(rng.uniform(1, 20, Bb)).to(dev))
            for bm in (big, big.with_slots(2)):
                ms = timed(lambda: rk4_steps(bm, X, 0.0, 5e-6, steps, u=imp, controller=ctrl))
                yield {"config": "LQR rollout with one gain per member, %d x %d el" % (Bb, N), "m": bm._plan.m, "g": bm._plan.g,
                                  "member_steps_per_s": Bb * steps / (ms * 1e-3), "ms": ms}
        if name == "cfg1e":  # config 1 as an ensemble: linear 10-element cantilever, gravity, tip impulse, per-member E
            B, steps = 131072, 100
            rng = np.random.default_rng(2)
            e = ens.config3(8, 10)
            par = np.repeat(params(e, 1), B, axis=0)
            par[:, :, 1] = 75e9 * np.exp(0.2 * rng.standard_normal((B, 1)))
            beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * 10}, ForceParams(enable_gravity_effects=True))
            beam.create_system_func(); beam.create_input_func()
            X = torch.zeros(B, 60, dtype=torch.float64, device=dev)
            imp = TipImpulse(torch.full((B,), 0.1, dtype=torch.float64, device=dev))
            ms = timed(lambda: rk4_steps(beam, X, 0.0, 2.5e-5, steps, u=imp))
            yield {"config": "cfg1 x%d (linear 10 el + gravity + impulse), RK4" % B, "m": beam._plan.m, "g": beam._plan.g,
                              "element_steps_per_s": B * 10 * steps / (ms * 1e-3), "ms": ms}
        if name == "cfg2m":  # config 2 shape with PER-MEMBER density and stiffness (no shared factor set)
            B, steps = 32768, 100
            e = ens.config2()
            rng = np.random.default_rng(9)
            par = np.repeat(params(e, 1), B, axis=0)
            par[:, :, 3] *= np.exp(0.1 * rng.standard_normal((B, 1)))
            par[:, :, 1] *= np.exp(0.2 * rng.standard_normal((B, 1)))
            beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["nonlinear"] * 20},
                                                    ForceParams(fluid_density=1000.0, enable_fluid_effects=True))
            beam.create_system_func(); beam.create_input_func()
            X = torch.zeros(B, 120, dtype=torch.float64, device=dev)
            imp = TipImpulse(torch.full((B,), 0.1, dtype=torch.float64, device=dev))
            ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps, u=imp))
            yield {"config": "cfg2 x%d with per-member density and stiffness, RK4" % B,
                              "element_steps_per_s": B * 20 * steps / (ms * 1e-3), "ms": ms}
        if name == "cfg3g":  # config 3 through the GENERAL kernel (comparison)
            B, steps = 65536, 50
            e = ens.config3(B, 32)
            beam = BatchedDynamicEulerBernoulliBeam({"params": params(e, B), "type": ["linear"] * 32})
            beam.create_system_func(); beam.create_input_func()
            beam.force_general_kernels = True
            X = torch.from_numpy(np.concatenate([e.q0, e.v0], axis=1)).to(dev)
            ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps))
            yield {"config": "cfg3 general kernel", "element_steps_per_s": B * 32 * steps / (ms * 1e-3), "ms": ms}
        if name == "cfg3mid":  # config 3 with the implicit midpoint rule (h = 10 x the RK4 step)
            from continuum_robot_b200 import midpoint_steps
            B, steps = 65536, 50
            e = ens.config3(B, 32)
            beam = BatchedDynamicEulerBernoulliBeam({"params": params(e, B), "type": ["linear"] * 32})
            beam.create_system_func(); beam.create_input_func()
            X = torch.from_numpy(np.concatenate([e.q0, e.v0], axis=1)).to(dev)
            hm = 10 * e.h
            ms = timed(lambda: midpoint_steps(beam, X, 0.0, hm, steps))
            yield {"config": "cfg3 implicit midpoint, h = 2e-4 (per-member factors of M + h^2/4 K)",
                              "element_steps_per_s": B * 32 * steps / (ms * 1e-3), "ms": ms,
                              "simulated_seconds_per_second_per_member": steps * hm / (ms * 1e-3)}
        if name == "cfg3i":  # config 3 with a per-member tip impulse (the reference examples' input): IMP variant
            B, steps = 65536, 50
            e = ens.config3(B, 32)
            beam = BatchedDynamicEulerBernoulliBeam({"params": params(e, B), "type": ["linear"] * 32})
            beam.create_system_func(); beam.create_input_func()
            X = torch.from_numpy(np.concatenate([e.q0, e.v0], axis=1)).to(dev)
            imp = TipImpulse(torch.from_numpy(np.random.default_rng(3).uniform(0.05, 0.5, B)).to(dev), duration=1.0)
            ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps, u=imp))
            yield {"config": "cfg3 + tip impulse (paired fast kernel, forcing variant)",
                              "element_steps_per_s": B * 32 * steps / (ms * 1e-3), "ms": ms}
        if name == "cfg3m":  # config 3 with PER-MEMBER mass (density varies per member): no shared factors
            B, steps = 65536, 50
            e = ens.config3(B, 32)
            par = params(e, B)
            par[:, :, 3] *= np.exp(0.1 * np.random.default_rng(5).standard_normal((B, 1)))
            beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * 32})
            beam.create_system_func(); beam.create_input_func()
            X = torch.from_numpy(np.concatenate([e.q0, e.v0], axis=1)).to(dev)
            ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps))
            yield {"config": "cfg3 with per-member mass (paired fast kernel, per-member factor sets in shared memory)",
                              "element_steps_per_s": B * 32 * steps / (ms * 1e-3), "ms": ms}
        if name in ("cfg4", "cfg4x4"):  # nonlinear 64-element, drag + gravity, adaptive RK45 to 3 ms
            B = 4096 if name == "cfg4" else 16384  # x4: BASELINE's 4096 members are 3.5 waves of warps on a B200; four times as many show the rate without the partial last wave
            e = ens.config4(B)
            if os.environ.get("CRB_CFG4_UNIFORM"):  # experiments: identical members (same number of attempts everywhere)
                e.E[:] = e.E[0]
                e.impulse_amp[:] = e.impulse_amp[0]
            beam = BatchedDynamicEulerBernoulliBeam({"params": params(e, B), "type": ["nonlinear"] * 64},
                                                    ForceParams(fluid_density=1000.0, enable_fluid_effects=True, enable_gravity_effects=True),
                                                    max_slots_per_lane=a.slots)
            beam.create_system_func(); beam.create_input_func()
            X0 = torch.zeros(B, 384, dtype=torch.float64, device=dev)
            imp = TipImpulse(torch.from_numpy(e.impulse_amp).to(dev))
            out = {}

            def run():
                out["r"] = solve_ensemble(beam, (0.0, 0.003), X0, method="RK45", rtol=1e-6, atol=1e-9, u=imp,
                                          t_eval=np.linspace(0, 0.003, 7))

            ms = timed(run, reps=2)
            r = out["r"]
            att = (r.naccept + r.nreject).double()
            yield {"config": "cfg4 %d x 64 nonlinear RK45 to 3 ms" % B, "m": beam._plan.m, "g": beam._plan.g, "ms": ms,
                              "attempts_mean": float(att.mean()), "attempts_max": float(att.max()),
                              "element_attempts_per_s": float(att.sum()) * 64 / (ms * 1e-3), "success": r.success}
        if name == "cfg5":  # LQR rollout, shared N = 6 design, 131072 members per GPU
            B, steps = 131072, 200
            e = ens.config5(B)
            beam = BatchedDynamicEulerBernoulliBeam({"params": params(e, 1), "type": ["linear"] * 6},
                                                    ForceParams(enable_gravity_effects=True), max_slots_per_lane=a.slots)
            beam.create_system_func(); beam.create_input_func()
            n = beam.n_free
            Q = np.eye(2 * n); Q[:n, :n] *= 100; Q[n:, n:] *= 10
            K = LinearQuadraticRegulator(beam.beam_model.get_stiffness_matrix(), beam.beam_model.get_mass_matrix(), Q, np.eye(n)).compute_gain_matrix()
            ctrl = FullStateLinear(torch.from_numpy(K).to(dev))
            X = torch.zeros(B, 2 * n, dtype=torch.float64, device=dev)
            imp = TipImpulse(torch.from_numpy(e.impulse_amp).to(dev))
            for general in (False, True):  # shared-operator tensor-core kernel, then the banded per-member kernel
                beam.force_general_kernels = general
                X.zero_()
                ms = timed(lambda: rk4_steps(beam, X, 0.0, e.h, steps, u=imp, controller=ctrl))
                yield {"config": "cfg5 LQR rollout 131072 x 6 el, RK4, " + ("banded kernel + DMMA feedback" if general else "shared-operator DMMA kernel"),
                                  "element_steps_per_s": B * 6 * steps / (ms * 1e-3), "member_steps_per_s": B * steps / (ms * 1e-3), "ms": ms}
            beam.force_general_kernels = False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="cfg2,cfg4,cfg5,cfg3g")
    ap.add_argument("--slots", type=int, default=0)
    a = ap.parse_args()
    for r in run(a.only.split(","), a.slots):
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
