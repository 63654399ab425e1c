"""GPU parity tests: CUDA path (through the C ABI) vs the reference's golden vectors / the oracle."""

import numpy as np
import pytest

from helpers import block_err, case_names, load, make_gpu_beam, oracle_spec, params_array

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RHS = load("rhs_cases.npz")


@pytest.mark.parametrize("name", case_names(RHS))
@pytest.mark.parametrize("slots", [0, 2, 1])
def test_rhs_matches_reference(name, slots):
    """get_dynamic_system()(t, x, u) on the reference's fixtures and BC edge cases (bit-level
    quirks Q1-Q4 included); tolerance 1e-11 relative to the block inf-norm (FP64 noise ~1e-14)."""
    p = name + "/"
    par = params_array(RHS, p)
    beam = make_gpu_beam(par, RHS[p + "elem_type"], RHS[p + "bc"], RHS[p + "fluid_density"], RHS[p + "gravity"],
                         RHS[p + "gravity_vector"], max_slots_per_lane=slots)
    X = torch.from_numpy(RHS[p + "X"]).cuda()
    U = torch.from_numpy(RHS[p + "U"]).cuda()
    Y = beam.get_dynamic_system()(0.3, X, U).cpu().numpy()
    ref = RHS[p + "Y"]
    n = beam.n_free
    assert Y.shape == ref.shape
    for i in range(len(ref)):
        assert block_err(Y[i], ref[i], n) < 1e-11, (name, i)
    # single-member (1-D) call mirrors the reference signature
    y0 = beam.get_dynamic_system()(0.3, X[2], U[2]).cpu().numpy()
    assert block_err(y0, ref[2], n) < 1e-11


@pytest.mark.parametrize("cfg", ["cfg1", "cfg2"])
def test_rk4_configs_1_2(cfg):
    """BASELINE configs 1 / 2: 4000 classical RK4 steps, every 40th state vs the reference (<= 1e-9)."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble

    g = load("cfg12.npz")
    p = cfg + "/"
    par = params_array(g, p)
    beam = make_gpu_beam(par, g[p + "elem_type"], g[p + "bc"], 1000.0 if cfg == "cfg2" else 0.0, cfg == "cfg1")
    n = beam.n_free
    X0 = torch.zeros(1, 2 * n, dtype=torch.float64, device="cuda")
    res = solve_ensemble(beam, (0.0, 0.1), X0, method="RK4", h=2.5e-5, save_every=40,
                         u=TipImpulse(torch.tensor([0.1], dtype=torch.float64, device="cuda")))
    got = res.y[0].T.cpu().numpy()[1:]  # drop t0
    ref = g[p + "Y"]
    assert got.shape == ref.shape
    worst = max(block_err(got[k], ref[k], n) for k in range(len(ref)))
    assert worst < 1e-9, worst


def test_rk4_config3_samples():
    """BASELINE config 3: 64 sampled members, 1000 RK4 steps, checkpoints every 250 (<= 1e-9)."""
    from continuum_robot_b200 import solve_ensemble

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
    got = res.y.permute(0, 2, 1).cpu().numpy()[:, 1:]  # [B, 4, 2n]
    ref = g["Y"]
    worst = max(block_err(got[i, k], ref[i, k], n) for i in range(B) for k in range(4))
    assert worst < 1e-9, worst


@pytest.mark.parametrize("cfg,rtol,atol", [("cfg2_rk45", 1e-6, 1e-9), ("cfg1_rk45", 1e-3, 1e-6)])
def test_rk45_single(cfg, rtol, atol):
    """Adaptive RK45 vs scipy.solve_ivp on the reference RHS: |dy| <= 10 (atol + rtol |y|) at t_eval,
    and the same number of RHS evaluations within 1 %."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble

    g = load("cfg12.npz")
    p = cfg + "/"
    par = params_array(g, p)
    nl = cfg.startswith("cfg2")
    beam = make_gpu_beam(par, g[p + "elem_type"], g[p + "bc"], 1000.0 if nl else 0.0, not nl)
    n = beam.n_free
    X0 = torch.zeros(1, 2 * n, dtype=torch.float64, device="cuda")
    te = g[p + "rk45_t"]
    res = solve_ensemble(beam, (0.0, 0.02), X0, method="RK45", t_eval=te, rtol=rtol, atol=atol,
                         u=TipImpulse(torch.tensor([0.1], dtype=torch.float64, device="cuda")))
    assert res.success
    got = res.y[0].cpu().numpy()
    ref = g[p + "rk45_y"]
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= 10 * (atol + rtol * np.abs(ref)))
    nfev_ref = int(g[p + "rk45_nfev"])
    assert abs(int(res.nfev[0]) - nfev_ref) <= max(12, 0.01 * nfev_ref), (int(res.nfev[0]), nfev_ref)


@pytest.mark.parametrize("N,B", [(32, 37), (6, 5), (64, 9), (16, 33)])
def test_rk4_fast_kernel_matches_general_and_oracle(N, B):
    """The linear / uniform-mass fast kernel (Nystrom form, recomputed spikes) against the general
    kernel (<= 1e-11, both are the same RK4 up to rounding) and against the NumPy oracle (<= 1e-9);
    ragged member counts exercise the inactive-lane paths."""
    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    e = ens.config3(B, N, seed=7)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    n = beam.n_free
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    amp = torch.linspace(0.05, 0.5, B, dtype=torch.float64, device="cuda")
    imp = TipImpulse(amp, duration=30 * e.h)
    steps = 60
    Xf = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xf, 0.0, e.h, steps, u=imp)
    beam.force_general_kernels = True
    Xg = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xg, 0.0, e.h, steps, u=imp)
    f, g = Xf.cpu().numpy(), Xg.cpu().numpy()
    assert max(block_err(f[i], g[i], n) for i in range(B)) < 1e-11
    for i in (0, B - 1):
        spec = bo.BeamSpec.uniform(N)
        spec.elastic_modulus = e.E[i].copy()
        b = bo.BeamOracle(spec)

        def u(t, a=float(amp[i])):
            v = np.zeros(n)
            if t < 30 * e.h:
                v[-2] = a
            return v

        ref = bo.rk4_solve(lambda t, x: b.rhs(t, x, u), x0[i], 0.0, e.h, steps)
        assert block_err(f[i], ref, n) < 1e-9


def test_rk45_config4_samples():
    """BASELINE config 4 (nonlinear, 64 elements, drag + gravity, per-member E and impulse), 16
    sampled members, adaptive per-member dt: outputs within 10 (atol + rtol |y|) of SciPy RK45 on
    the reference RHS, nfev within 2 %.  Horizon 3 ms: the reference's own 64-element nonlinear
    beam diverges after ~5 ms."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble

    g = load("cfg4_samples.npz")
    B, N = g["E_parsed"].shape
    par = np.zeros((B, N, 7))
    for col, key in ((0, "length"), (2, "moment_inertia"), (3, "density"), (4, "cross_area"), (5, "wetted_area"), (6, "drag_coef")):
        par[:, :, col] = g[key][None, :]
    par[:, :, 1] = g["E_parsed"]
    beam = make_gpu_beam(par, np.ones(N, dtype=int), np.array([1] + [0] * N), 1000.0, True)
    n = beam.n_free
    X0 = torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda")
    rtol, atol = float(g["rtol"]), float(g["atol"])
    res = solve_ensemble(beam, (0.0, 0.003), X0, method="RK45", t_eval=g["t_eval"], rtol=rtol, atol=atol,
                         u=TipImpulse(torch.from_numpy(g["amp"]).cuda()))
    assert res.success
    got, ref = res.y.cpu().numpy(), g["y"]
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= 10 * (atol + rtol * np.abs(ref)))
    nfev = res.nfev.cpu().numpy()
    assert np.all(np.abs(nfev - g["nfev"]) <= np.maximum(12, 0.02 * g["nfev"])), (nfev, g["nfev"])
    # per-member step counts differ across the ensemble: the controller really is per member
    assert len(set(nfev.tolist())) > 4


def test_rk45_config4_at_5ms():
    """Config 4 also pinned at 5 ms (tests/golden/cfg4_5ms.npz, 4 members): the reference is still finite there but
    already growing fast (max |y| ~ 2.5e2), which is where a controller mismatch would show first."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble

    g = load("cfg4_5ms.npz")
    B, N = g["E_parsed"].shape
    par = np.zeros((B, N, 7))
    for col, key in ((0, "length"), (2, "moment_inertia"), (3, "density"), (4, "cross_area"), (5, "wetted_area"), (6, "drag_coef")):
        par[:, :, col] = g[key][None, :]
    par[:, :, 1] = g["E_parsed"]
    beam = make_gpu_beam(par, np.ones(N, dtype=int), np.array([1] + [0] * N), 1000.0, True)
    n = beam.n_free
    rtol, atol = float(g["rtol"]), float(g["atol"])
    res = solve_ensemble(beam, (0.0, 0.005), torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda"), method="RK45",
                         t_eval=g["t_eval"], rtol=rtol, atol=atol, u=TipImpulse(torch.from_numpy(g["amp"]).cuda()))
    assert res.success
    got, ref = res.y.cpu().numpy(), g["y"]
    assert got.shape == ref.shape and np.abs(ref).max() > 50.0
    assert np.all(np.abs(got - ref) <= 10 * (atol + rtol * np.abs(ref))), (np.abs(got - ref) / (atol + rtol * np.abs(ref))).max()
    nfev = res.nfev.cpu().numpy()
    assert np.all(np.abs(nfev - g["nfev"]) <= np.maximum(12, 0.02 * g["nfev"])), (nfev, g["nfev"])


def test_lqr_rollout_config5_samples():
    """BASELINE config 5: closed-loop LQR rollout (shared gain from the golden file, u = K(0 - x)
    plus a per-member tip disturbance, gravity on), 2000 RK4 steps, <= 1e-9 at 4 checkpoints."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse, solve_ensemble

    g = load("cfg5_samples.npz")
    par = params_array(g)[None]
    beam = make_gpu_beam(par, g["elem_type"], g["bc"], 0.0, True)
    n = beam.n_free
    B = len(g["amp"])
    ctrl = FullStateLinear(torch.from_numpy(g["gain"]).cuda())
    X0 = torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda")
    res = solve_ensemble(beam, (0.0, 2000 * float(g["h"])), X0, method="RK4", h=float(g["h"]), save_every=500,
                         u=TipImpulse(torch.from_numpy(g["amp"]).cuda()), controller=ctrl)
    got = res.y.permute(0, 2, 1).cpu().numpy()[:, 1:]
    ref = g["Y"]
    assert got.shape == ref.shape
    worst = max(block_err(got[i, k], ref[i, k], n) for i in range(B) for k in range(ref.shape[1]))
    assert worst < 1e-9, worst
    # host-side synthesis gives the same gain as the one stored with the golden trajectories
    from continuum_robot_b200 import LinearQuadraticRegulator

    Kb, Mb = beam.beam_model.get_stiffness_matrix(), beam.beam_model.get_mass_matrix()
    assert np.abs(Kb - g["K_beam"]).max() <= 1e-14 * np.abs(Kb).max()
    Q = np.eye(2 * n)
    Q[:n, :n] *= 100
    Q[n:, n:] *= 10
    K = LinearQuadraticRegulator(Kb, Mb, Q, np.eye(n)).compute_gain_matrix()
    assert np.abs(K - g["gain"]).max() <= 1e-8 * np.abs(g["gain"]).max()


@pytest.mark.parametrize("N,B", [(32, 37), (64, 9), (6, 5)])
def test_rk4_paired_linear_kernel(N, B):
    """Autonomous force-free linear case (config 3): the paired operator form (two rounds of two
    independent applications of -M^-1 K = the degree-4 Taylor polynomial of the classical tableau)
    against the stage-by-stage fast kernel and the general kernel (<= 1e-11) and the oracle."""
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    e = ens.config3(B, N, seed=3)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    n = beam.n_free
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    steps = 80
    out = {}
    for mode in ("paired", "staged", "general"):
        beam.force_staged_kernels = mode == "staged"
        beam.force_general_kernels = mode == "general"
        X = torch.from_numpy(x0).cuda()
        Y = torch.empty(2, B, 2 * n, dtype=torch.float64, device="cuda")
        rk4_steps(beam, X, 0.0, e.h, steps, Y_out=Y, save_every=40)
        assert torch.equal(Y[1], X)
        out[mode] = (X.cpu().numpy(), Y[0].cpu().numpy())
    for mode in ("staged", "general"):
        for k in range(2):
            assert max(block_err(out["paired"][k][i], out[mode][k][i], n) for i in range(B)) < 1e-11, mode
    i = B - 1
    spec = bo.BeamSpec.uniform(N)
    spec.elastic_modulus = e.E[i].copy()
    b = bo.BeamOracle(spec)
    ref = bo.rk4_solve(lambda t, x: b.rhs(t, x, np.zeros(n)), x0[i], 0.0, e.h, steps)
    assert block_err(out["paired"][0][i], ref, n) < 1e-9


@pytest.mark.parametrize("N", [6, 10, 13])
def test_feedback_tensor_core_path_matches_scalar_path(N):
    """u_c = K (r - x): FP64 mma path (4 lanes per member) vs the shared-memory scalar path (other
    lane layouts) vs the oracle, with a non-zero reference, gravity, drag and an impulse."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(N)
    spec = bo.BeamSpec.uniform(N)
    fs = bo.ForceSpec(1000.0, True, (0.3, -9.81, 0.0), True)
    b = bo.BeamOracle(spec, fs)
    n = b.n
    K = rng.standard_normal((n, 2 * n)) * 0.5  # small enough for h = 5e-6 to stay inside RK4's stability region
    ref = rng.standard_normal(2 * n) * 1e-3
    par = np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                    spec.wetted_area, spec.drag_coef], axis=1)[None]
    B = 21
    x0 = np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)
    amp = np.linspace(1.0, 5.0, B)
    res = {}
    for slots in (2, 4, 1):
        beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N), 1000.0, True, (0.3, -9.81, 0.0),
                             max_slots_per_lane=slots)
        ctrl = FullStateLinear(torch.from_numpy(K).cuda(), reference=torch.from_numpy(ref).cuda())
        X = torch.from_numpy(x0).cuda()
        imp = TipImpulse(torch.from_numpy(amp).cuda(), duration=1e-4)
        # explicit system: no automatic re-layout, so every lane decomposition is really exercised
        drag, grav, _ = beam._active_forces()
        system = beam.make_system(B, drag=drag, gravity=grav, impulse=imp, gain=ctrl.gain_matrix, ref=ctrl.reference)
        rk4_steps(beam, X, 0.0, 5e-6, 40, system=system)
        res[(int(beam._plan.m), int(beam._plan.g))] = X.cpu().numpy()
    assert any(g == 4 for (_, g) in res) and any(g != 4 for (_, g) in res), list(res)
    vals = list(res.values())
    for v in vals[1:]:
        assert max(block_err(v[i], vals[0][i], n) for i in range(B)) < 1e-11
    i = 7

    def f(t, x):
        u = np.zeros(n)
        if t < 1e-4:
            u[-2] = amp[i]
        return b.rhs(t, x, u + bo.full_state_feedback(K, x, ref))

    assert block_err(vals[0][i], bo.rk4_solve(f, x0[i], 0.0, 5e-6, 40), n) < 1e-9


@pytest.mark.parametrize("N,B", [(32, 19), (6, 7)])
def test_rk4_paired_kernel_with_forcing(N, B):
    """Paired operator form with piecewise-constant forcing: constant generalized force, tip impulse
    that switches off in the middle of the run (so some steps straddle the switch), and both;
    against the general kernel (<= 1e-11), the stage-by-stage fast kernel and the oracle (<= 1e-9)."""
    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    e = ens.config3(B, N, seed=21)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    n = beam.n_free
    rng = np.random.default_rng(N)
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    U = rng.standard_normal((B, n))
    amp = np.linspace(0.5, 3.0, B)
    steps, h = 64, e.h
    dur = 20.3 * h  # the impulse ends inside step 20: stage times t, t+h/2 are inside, t+h outside
    for uc, imp in ((True, False), (False, True), (True, True)):
        def fresh():
            X = torch.from_numpy(x0).cuda()
            drag, grav, _ = beam._active_forces()
            system = beam.make_system(B, u_const=torch.from_numpy(U).cuda() if uc else None,
                                      impulse=TipImpulse(torch.from_numpy(amp).cuda(), duration=dur) if imp else None)
            rk4_steps(beam, X, 0.0, h, steps, system=system)
            return X.cpu().numpy()

        beam.force_general_kernels = beam.force_staged_kernels = False
        paired = fresh()
        beam.force_general_kernels = True
        general = fresh()
        beam.force_general_kernels = False
        assert max(block_err(paired[i], general[i], n) for i in range(B)) < 1e-11, (uc, imp)
        if not uc:
            beam.force_staged_kernels = True
            staged = fresh()
            beam.force_staged_kernels = False
            assert max(block_err(paired[i], staged[i], n) for i in range(B)) < 1e-11
        i = B // 2
        spec = bo.BeamSpec.uniform(N)
        spec.elastic_modulus = e.E[i].copy()
        b = bo.BeamOracle(spec)

        def u(t):
            v = U[i].copy() if uc else np.zeros(n)
            if imp and t < dur:
                v[-2] += amp[i]
            return v

        ref = bo.rk4_solve(lambda t, x: b.rhs(t, x, u), x0[i], 0.0, h, steps)
        assert block_err(paired[i], ref, n) < 1e-9, (uc, imp)


@pytest.mark.parametrize("seed", range(12))
def test_rhs_random_topologies_match_oracle(seed):
    """Randomised beams (element count, mixed element types, FIXED / PINNED anywhere, non-uniform
    properties, drag, tilted gravity, per-member parameters) -- RHS vs the oracle, which is itself
    pinned to the reference on every BC pattern of the golden set.  Tolerance 1e-10 (block inf-norm)."""
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(1000 + seed)
    N = int(rng.integers(1, 14))
    B = int(rng.integers(1, 6))
    et = rng.integers(0, 2, N)
    bc = rng.choice([0, 0, 0, 1, 2], size=N)
    if rng.random() < 0.6:
        bc[0] = 1
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    sc = np.exp(0.3 * rng.standard_normal((B, N, 7)))
    base = np.array([m["length"], m["E"], m["I"], m["rho"], m["A"], m["wetted_area"], m["drag_coef"]])
    par = base[None, None, :] * sc
    if rng.random() < 0.5:  # shared mass / geometry, per-member stiffness only
        par[:, :, [0, 3, 4, 5, 6]] = par[0:1, :, [0, 3, 4, 5, 6]]
    fd = float(rng.choice([0.0, 1000.0]))
    grav = bool(rng.random() < 0.7)
    gvec = (float(rng.normal()), -9.81, 0.0)
    slots = int(rng.choice([0, 1, 2, 3]))
    beam = make_gpu_beam(par, et, np.append(bc, 0), fd, grav, gvec, max_slots_per_lane=slots)
    n = beam.n_free
    X = np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)
    U = rng.standard_normal((B, n))
    Y = beam.get_dynamic_system()(0.1, torch.from_numpy(X).cuda(), torch.from_numpy(U).cuda()).cpu().numpy()
    for i in range(B):
        spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], et, bc, par[i, :, 5], par[i, :, 6])
        b = bo.BeamOracle(spec, bo.ForceSpec(fd, fd > 0, gvec, grav))
        assert b.n == n
        assert block_err(Y[i], b.rhs(0.1, X[i], U[i]), n) < 1e-10, (seed, i, N, list(bc), list(et))


def test_rk45_pilot_launch_and_member_order_change_nothing():
    """solve_ensemble(method="RK45") may split the run into a pilot launch (attempt budget) and a main launch that
    resumes every member and hands the members out longest-first (crb_system_t.member_order).  A resumed member takes
    exactly the step sequence of an uninterrupted run: final state, time, step size, counters and dense output are
    bitwise equal for pilots of 1, 5 and 12 attempts -- also when the pilot ends on a rejected attempt."""
    from continuum_robot_b200 import ForceParams, TipImpulse, solve_ensemble
    from continuum_robot_b200 import ensembles as ens

    B = 192
    e = ens.config4(B, seed=9)
    m = ens.material()
    par = np.empty((B, 64, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    from continuum_robot_b200 import BatchedDynamicEulerBernoulliBeam

    beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["nonlinear"] * 64},
                                            ForceParams(fluid_density=1000.0, enable_fluid_effects=True, enable_gravity_effects=True))
    beam.create_system_func(); beam.create_input_func()
    imp = TipImpulse(torch.from_numpy(e.impulse_amp).cuda())
    X0 = torch.zeros(B, 384, dtype=torch.float64, device="cuda")
    te = np.linspace(0.0, 1e-3, 6)
    runs = {}
    for pilot in (0, 1, 5, 12):
        r = solve_ensemble(beam, (0.0, 1e-3), X0, method="RK45", rtol=1e-6, atol=1e-9, u=imp, t_eval=te, pilot_attempts=pilot)
        assert r.success
        runs[pilot] = (r.x_final.clone(), r.y.clone(), r.nfev.clone(), r.naccept.clone(), r.nreject.clone(), r.t_final.clone(), r.h_last.clone())
    assert int(runs[0][4].sum()) > 0  # some attempts are rejected in this ensemble
    for pilot in (1, 5, 12):
        for a, b in zip(runs[0], runs[pilot]):
            assert torch.equal(a, b), pilot


def test_rk45_step_sequence_equals_scipy_controller():
    """GPU RK45 vs the oracle's restated SciPy controller on the same beam: identical accepted /
    rejected step counts and nfev, outputs equal to ~1e-9 of the tolerance band."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble
    from oracle import beam_oracle as bo

    N = 8
    spec = bo.BeamSpec.uniform(N, elem_type=bo.NONLINEAR)
    fs = bo.ForceSpec(1000.0, True, (0.0, -9.81, 0.0), True)
    b = bo.BeamOracle(spec, fs)
    n = b.n
    par = np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                    spec.wetted_area, spec.drag_coef], axis=1)[None]
    beam = make_gpu_beam(par, np.ones(N, dtype=int), np.array([1] + [0] * N), 1000.0, True)
    amps = np.array([0.1, 0.3])
    te = np.linspace(0.0, 0.004, 9)
    res = solve_ensemble(beam, (0.0, 0.004), torch.zeros(2, 2 * n, dtype=torch.float64, device="cuda"), method="RK45",
                         t_eval=te, rtol=1e-6, atol=1e-9, u=TipImpulse(torch.from_numpy(amps).cuda()))
    for i, amp in enumerate(amps):
        def u(t, a=amp):
            v = np.zeros(n)
            if t < 0.01:
                v[-2] = a
            return v

        r = bo.rk45_solve(lambda t, x: b.rhs(t, x, u), (0.0, 0.004), np.zeros(2 * n), t_eval=te, rtol=1e-6, atol=1e-9)
        assert int(res.nfev[i]) == r.nfev and int(res.naccept[i]) == r.naccept and int(res.nreject[i]) == r.nreject
        got = res.y[i].cpu().numpy()
        assert np.all(np.abs(got - r.y) <= 1e-3 * (1e-9 + 1e-6 * np.abs(r.y)))


def test_lqr_rollout_config5_banded_path():
    """Same golden rollout through the per-member banded kernel (feedback on the tensor cores only):
    the path taken when designs differ per member."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps

    g = load("cfg5_samples.npz")
    beam = make_gpu_beam(params_array(g)[None], g["elem_type"], g["bc"], 0.0, True)
    beam.force_general_kernels = True
    n, B, h = beam.n_free, len(g["amp"]), float(g["h"])
    ctrl = FullStateLinear(torch.from_numpy(g["gain"]).cuda())
    X = torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda")
    rk4_steps(beam, X, 0.0, h, 500, u=TipImpulse(torch.from_numpy(g["amp"]).cuda()), controller=ctrl)
    got = X.cpu().numpy()
    assert max(block_err(got[i], g["Y"][i, 0], n) for i in range(B)) < 1e-9


@pytest.mark.parametrize("N,bc0,gravity,with_ref,with_imp,decoupled", [
    (6, 1, True, False, True, False), (6, 1, False, True, False, False), (3, 1, True, True, True, False),
    (8, 1, True, False, False, False), (4, 2, True, True, True, False), (1, 1, True, False, True, False),
    (5, 1, False, False, False, False),
    (6, 1, True, True, True, True), (3, 1, True, False, True, True), (4, 1, True, True, False, True),
    (5, 1, True, False, True, True), (7, 1, True, True, True, True), (8, 1, False, True, True, True),
    (4, 2, True, True, True, True), (2, 1, True, False, False, True),
])
def test_shared_operator_path_matches_banded_path_and_oracle(N, bc0, gravity, with_ref, with_imp, decoupled):
    """Shared-operator RK4 (dense FP64 tensor-core contraction, crb_shared_operator) vs the banded
    per-member kernel and vs the CPU oracle: gains with a non-zero reference, FIXED and PINNED roots,
    gravity on/off, impulse on/off, 1..8 elements (<= 1e-9 block inf-norm over 300 steps).
    `decoupled`: a gain without axial / bending coupling and gravity along -y, as in examples/lqr_control.py --
    the operator's DOFs are reordered and the kernel's tile-skipping code path runs (where one is compiled for the
    shape; tests/test_host_api.py checks which)."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(100 + N)
    m = ens.material()
    par = np.zeros((1, N, 7))
    par[0, :, 0] = m["length"] * (1 + 0.1 * rng.random(N))
    par[0, :, 1], par[0, :, 2], par[0, :, 3], par[0, :, 4] = m["E"], m["I"], m["rho"] * (1 + 0.2 * rng.random(N)), m["A"]
    par[0, :, 5], par[0, :, 6] = m["wetted_area"], m["drag_coef"]
    bc = np.array([bc0] + [0] * N)
    et = np.zeros(N, dtype=int)
    gvec = (0.0, -9.81, 0.0) if decoupled else (1.5, -9.81, 0.0)
    beam = make_gpu_beam(par, et, bc, 0.0, gravity, gvec)
    n, B, h, steps = beam.n_free, 37, 2e-6, 300
    # (stiffness-like and damping-like feedback small enough that the closed loop stays tame over the run)
    gain = np.concatenate([50.0 * rng.standard_normal((n, n)), 0.05 * rng.standard_normal((n, n))], axis=1)
    if decoupled:
        axial = np.array(sum(([True, False, False] if b == 0 else ([False] if b == 2 else []) for b in bc), []))
        cross = axial[:, None] != axial[None, :]
        gain[np.concatenate([cross, cross], axis=1)] = 0.0
    ref = 1e-3 * rng.standard_normal(2 * n) if with_ref else None
    ctrl = FullStateLinear(torch.from_numpy(gain).cuda(), reference=torch.from_numpy(ref).cuda() if with_ref else None)
    amp = rng.uniform(1.0, 5.0, B)
    imp = TipImpulse(torch.from_numpy(amp).cuda(), duration=2.5e-4) if with_imp else None
    x0 = np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)
    Xs = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xs, 0.0, h, steps, u=imp, controller=ctrl)
    beam.force_general_kernels = True
    Xb = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xb, 0.0, h, steps, u=imp, controller=ctrl)
    beam.force_general_kernels = False
    gs, gb = Xs.cpu().numpy(), Xb.cpu().numpy()
    assert not np.array_equal(gs, gb)  # two different kernels really ran
    assert max(block_err(gs[i], gb[i], n) for i in range(B)) < 1e-9
    spec = bo.BeamSpec(par[0, :, 0], par[0, :, 1], par[0, :, 2], par[0, :, 3], par[0, :, 4], et, bc, par[0, :, 5], par[0, :, 6])
    orc = bo.BeamOracle(spec, bo.ForceSpec(gravity_vector=gvec, enable_gravity_effects=gravity))
    r = np.zeros(2 * n) if ref is None else ref
    for i in (0, B - 1):
        def f(t, x, i=i):
            u = gain @ (r - x)
            if with_imp and t < 2.5e-4:
                u = u.copy()
                u[n - 2] += amp[i]
            return orc.rhs(t, x, u)
        want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
        assert block_err(gs[i], want, n) < 1e-9, (i, block_err(gs[i], want, n))


@pytest.mark.parametrize("N,with_imp", [(32, False), (32, True), (16, False), (12, True)])
def test_per_member_mass_fast_path(N, with_imp):
    """Ensembles whose members differ in density / area / element length (also along the beam: the
    N = 16 case is tapered) take the paired fast kernel with per-member mass factors: compared with the
    general kernel on the whole ensemble and with the CPU oracle on sampled members (<= 1e-9)."""
    from continuum_robot_b200 import HostPipeline, TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    B, steps = 203, 120
    rng = np.random.default_rng(7 + N)
    e = ens.config3(B, N, seed=11)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0] = (m["length"] * (1 + 0.1 * rng.random(B)))[:, None]
    par[:, :, 1] = e.E
    par[:, :, 2] = m["I"]
    par[:, :, 3] = (m["rho"] * np.exp(0.2 * rng.standard_normal(B)))[:, None]
    par[:, :, 4] = (m["A"] * (1 + 0.1 * rng.random(B)))[:, None]
    if N == 16:  # tapered beams: density and element length vary ALONG the beam as well
        par[:, :, 3] *= np.linspace(1.0, 0.6, N)[None, :]
        par[:, :, 0] *= np.linspace(1.1, 0.9, N)[None, :]
    par[:, :, 5:] = 1.0
    et, bc = np.zeros(N, dtype=int), np.array([1] + [0] * N)
    beam = make_gpu_beam(par, et, bc)
    assert not beam._mass_shared
    n, h = beam.n_free, 1e-5
    amp = rng.uniform(0.05, 0.5, B)
    imp = TipImpulse(torch.from_numpy(amp).cuda(), duration=5e-4) if with_imp else None
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    Xf = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xf, 0.0, h, steps, u=imp)
    beam.force_general_kernels = True
    Xg = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xg, 0.0, h, steps, u=imp)
    beam.force_general_kernels = False
    gf, gg = Xf.cpu().numpy(), Xg.cpu().numpy()
    assert not np.array_equal(gf, gg)
    assert max(block_err(gf[i], gg[i], n) for i in range(B)) < 1e-9
    for i in (0, 77, B - 1):
        spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], et, bc, par[i, :, 5], par[i, :, 6])
        orc = bo.BeamOracle(spec)
        def f(t, x, i=i):
            u = np.zeros(n)
            if with_imp and t < 5e-4:
                u[n - 2] = amp[i]
            return orc.rhs(t, x, u)
        want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
        assert block_err(gf[i], want, n) < 1e-9, (i, block_err(gf[i], want, n))
    # chunked host pipeline slices the per-member factor sets and coupling blocks
    xh = torch.from_numpy(x0.copy()).pin_memory()
    pipe = HostPipeline(beam, B, chunk_members=64, u=imp)
    pipe.run(xh, 0.0, h, steps)
    pipe.synchronize()
    assert np.array_equal(xh.numpy(), gf)


@pytest.mark.parametrize("seed", range(8))
def test_stepping_random_topologies_match_oracle(seed):
    """Fused RK4 and adaptive RK45 on randomised beams (mixed element types, FIXED / PINNED anywhere,
    drag, tilted gravity, per-member parameters, impulse): the generic kernels with constrained DOFs inside
    active slots, vs the oracle stepping the same systems (RK4 <= 1e-9; RK45 inside the tolerance band)."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(5000 + seed)
    N = int(rng.integers(2, 12))
    B = int(rng.integers(1, 5))
    et = rng.integers(0, 2, N)
    bc = rng.choice([0, 0, 0, 1, 2], size=N)
    if rng.random() < 0.6:
        bc[0] = 1
    if np.all(bc == 0):
        bc[int(rng.integers(0, N))] = 2  # pin something: a completely free beam drifts as a rigid body
    m = ens.material()
    sc = np.exp(0.2 * rng.standard_normal((B, N, 7)))
    base = np.array([m["length"], m["E"], m["I"], m["rho"], m["A"], m["wetted_area"], m["drag_coef"]])
    par = base[None, None, :] * sc
    fd = float(rng.choice([0.0, 1000.0]))
    grav = bool(rng.random() < 0.7)
    gvec = (float(rng.normal()), -9.81, 0.0)
    beam = make_gpu_beam(par, et, np.append(bc, 0), fd, grav, gvec)
    n = beam.n_free
    amp = rng.uniform(0.05, 0.3, B)
    dof = int(rng.integers(0, n))
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    h, steps = 2e-6, 120
    imp = TipImpulse(torch.from_numpy(amp).cuda(), dof=dof, duration=70.5 * h)
    X = torch.from_numpy(x0).cuda()
    rk4_steps(beam, X, 0.0, h, steps, u=imp)
    got = X.cpu().numpy()
    te = np.linspace(0.0, 2e-4, 5)
    res = solve_ensemble(beam, (0.0, 2e-4), torch.from_numpy(x0).cuda(), method="RK45", t_eval=te, rtol=1e-6, atol=1e-9,
                         u=TipImpulse(torch.from_numpy(amp).cuda(), dof=dof, duration=1.0))
    for i in range(B):
        spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], et, bc, par[i, :, 5], par[i, :, 6])
        b = bo.BeamOracle(spec, bo.ForceSpec(fd, fd > 0, gvec, grav))
        def f(t, x, i=i, dur=70.5 * h):
            u = np.zeros(n)
            if t < dur:
                u[dof] = amp[i]
            return b.rhs(t, x, u)
        want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
        assert block_err(got[i], want, n) < 1e-9, (seed, i, N, list(bc), list(et), block_err(got[i], want, n))
        r = bo.rk45_solve(lambda t, x, i=i: f(t, x, i, 1.0), (0.0, 2e-4), x0[i], t_eval=te, rtol=1e-6, atol=1e-9)
        gy = res.y[i].cpu().numpy()
        assert np.all(np.abs(gy - r.y) <= 10 * (1e-9 + 1e-6 * np.abs(r.y))), (seed, i)
        assert abs(int(res.nfev[i]) - r.nfev) <= max(12, 0.02 * r.nfev)


@pytest.mark.parametrize("N,nonlinear", [(20, True), (10, False), (6, True)])
def test_per_member_mass_in_specialised_general_kernels(N, nonlinear):
    """Per-member density / area (no shared factor set) with drag, gravity and an impulse: the
    shape-specialised general kernels stage one compact factor copy per member; vs the oracle (<= 1e-9)."""
    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    B, h, steps = 37, 2.5e-6, 150
    rng = np.random.default_rng(40 + N)
    m = ens.material()
    base = np.array([m["length"], m["E"], m["I"], m["rho"], m["A"], m["wetted_area"], m["drag_coef"]])
    par = np.tile(base, (B, N, 1))
    par[:, :, 3] *= np.exp(0.2 * rng.standard_normal((B, 1)))
    par[:, :, 4] *= (1 + 0.1 * rng.random((B, 1)))
    par[:, :, 1] *= np.exp(0.2 * rng.standard_normal((B, N)))
    et = np.full(N, 1 if nonlinear else 0)
    bc = np.array([1] + [0] * N)
    beam = make_gpu_beam(par, et, bc, 1000.0, True, (0.5, -9.81, 0.0))
    assert not beam._mass_shared
    n = beam.n_free
    amp = rng.uniform(0.05, 0.3, B)
    imp = TipImpulse(torch.from_numpy(amp).cuda(), duration=1.0)
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    X = torch.from_numpy(x0).cuda()
    rk4_steps(beam, X, 0.0, h, steps, u=imp)
    got = X.cpu().numpy()
    for i in (0, 17, B - 1):
        spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], et, bc[:N], par[i, :, 5], par[i, :, 6])
        b = bo.BeamOracle(spec, bo.ForceSpec(1000.0, True, (0.5, -9.81, 0.0), True))
        def f(t, x, i=i):
            u = np.zeros(n)
            u[n - 2] = amp[i]
            return b.rhs(t, x, u)
        want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
        assert block_err(got[i], want, n) < 1e-9, (i, block_err(got[i], want, n))
    # adaptive integrator on the same per-member ensemble
    from continuum_robot_b200 import solve_ensemble

    te = np.linspace(0.0, 3e-4, 4)
    res = solve_ensemble(beam, (0.0, 3e-4), torch.from_numpy(x0).cuda(), method="RK45", t_eval=te, rtol=1e-6, atol=1e-9, u=imp)
    for i in (0, B - 1):
        spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], et, bc[:N], par[i, :, 5], par[i, :, 6])
        b = bo.BeamOracle(spec, bo.ForceSpec(1000.0, True, (0.5, -9.81, 0.0), True))
        def f2(t, x, i=i):
            u = np.zeros(n)
            u[n - 2] = amp[i]
            return b.rhs(t, x, u)
        r = bo.rk45_solve(f2, (0.0, 3e-4), x0[i], t_eval=te, rtol=1e-6, atol=1e-9)
        gy = res.y[i].cpu().numpy()
        assert np.all(np.abs(gy - r.y) <= 10 * (1e-9 + 1e-6 * np.abs(r.y))), i
        assert abs(int(res.nfev[i]) - r.nfev) <= max(12, 0.02 * r.nfev)
