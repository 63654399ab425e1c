"""GPU tests of time-varying inputs and plug-in forces under the fused and the unfused integrators.

Golden data: tests/golden/rk45_inputs.npz, made by tests/golden/make_golden.py::gen_inputs from the UNMODIFIED
reference (its own integration test /root/reference/tests/test_dynamic_beam.py:201-244, u(t) = sin(t) * ones(n) under
default RK45, and the mixed beam + StateAwareForce of tests/test_advanced_composition.py:13-65)."""

import numpy as np
import pytest

from helpers import block_err, load, make_gpu_beam, oracle_spec, params_array

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
G = load("rk45_inputs.npz")


def _beam(name, **kw):
    p = name + "/"
    return make_gpu_beam(params_array(G, p)[None], G[p + "elem_type"], G[p + "bc"], **kw)


def _check_rk45(res, name, nfev_tol=0.01, stability_limited=False):
    p = name + "/"
    rtol, atol = float(G[p + "rtol"]), float(G[p + "atol"])
    assert res.success
    got, ref = res.y[0].cpu().numpy(), G[p + "y"]
    assert got.shape == ref.shape
    band = atol + rtol * np.abs(ref)
    if stability_limited:
        # rtol = 1e-3 on these beams is stability-limited stepping: the axial modes sit on the stability boundary and
        # the accept / reject sequence amplifies rounding.  G[.../noise] is how far the unmodified reference's OWN
        # outputs move when u is scaled by (1 +- 1e-14), (1 +- 1e-13) (up to 27 tolerance bands, make_golden.py):
        # the reference solution is only determined up to that band, so it is added to the stated tolerance.
        assert np.all(np.abs(got - ref) <= 10 * band + 3 * G[p + "noise"]), (np.abs(got - ref) / band).max()
    else:
        assert np.all(np.abs(got - ref) <= 10 * band), (np.abs(got - ref) / band).max()
    nfev_ref = int(G[p + "nfev"])
    assert abs(int(res.nfev[0]) - nfev_ref) <= max(12, nfev_tol * nfev_ref), (int(res.nfev[0]), nfev_ref)


@pytest.mark.parametrize("name", ["sin_lin4", "sin_nl4"])
@pytest.mark.parametrize("how", ["fused", "callable"])
def test_rk45_sinusoidal_input_matches_reference(name, how):
    """The reference's own integration test: u(t) = sin(t) * ones(n), default RK45 tolerances, t in [0, 0.1]; fused
    (SinusoidInput inside crb_rk45) and as a free-form torch callable (unfused driver: crb_rhs per stage +
    crb_rk45_stage / crb_rk45_control)."""
    from continuum_robot_b200 import SinusoidInput, solve_ensemble

    beam = _beam(name)
    n = beam.n_free
    p = name + "/"
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    u = SinusoidInput(ones, omega=1.0) if how == "fused" else (lambda t: torch.sin(t) * ones)
    X0 = torch.zeros(1, 2 * n, dtype=torch.float64, device="cuda")
    res = solve_ensemble(beam, tuple(G[p + "t_span"]), X0, method="RK45", t_eval=G[p + "t_eval"], rtol=float(G[p + "rtol"]),
                         atol=float(G[p + "atol"]), u=u)
    _check_rk45(res, name, stability_limited=True)
    assert res.t[0] == 0.0 and res.t[-1] == 0.1  # what the reference's test asserts (sol.t[0], sol.t[-1])


def _state_aware_force(stiffness, damping):
    """tests/test_advanced_composition.py:36-65 as a torch plug-in (spring-damper on the tip w DOF)."""
    from continuum_robot_b200 import AbstractForce

    class StateAwareForce(AbstractForce):
        def compute_forces(self, x, t):
            ns = x.shape[-1] // 2
            f = torch.zeros(x.shape[:-1] + (ns,), dtype=x.dtype, device=x.device)
            f[..., ns - 2] = -stiffness * x[..., ns - 2] - damping * x[..., 2 * ns - 2]
            return f

        def is_enabled(self):
            return True

    return StateAwareForce()


@pytest.mark.parametrize("fused_input", [True, False])
def test_rk45_plugin_force_and_time_varying_input_match_reference(fused_input):
    """Mixed linear / nonlinear beam with drag + gravity, a state-aware plug-in force and a sinusoidal input on every
    DOF: SciPy RK45 on the reference vs the unfused adaptive driver (per-member controller on the device)."""
    from continuum_robot_b200 import SinusoidInput, solve_ensemble

    name = "plugin_mixed5"
    p = name + "/"
    beam = _beam(name, fluid_density=1000.0, gravity=True)
    beam.force_registry.register(_state_aware_force(500.0, 5.0))
    beam.create_system_func()
    n = beam.n_free
    amp = torch.from_numpy(G[p + "amp"]).cuda()
    w, ph = 2 * np.pi * 120.0, 0.3
    u = SinusoidInput(amp, omega=w, phase=ph) if fused_input else (lambda t: amp * torch.sin(w * t + ph))
    B = 3  # identical members: the per-member controller must reproduce the same result in every row
    X0 = torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda")
    res = solve_ensemble(beam, tuple(G[p + "t_span"]), X0, method="RK45", t_eval=G[p + "t_eval"],
                         rtol=float(G[p + "rtol"]), atol=float(G[p + "atol"]), u=u)
    _check_rk45(res, name)
    assert torch.equal(res.y[0], res.y[1]) and torch.equal(res.y[0], res.y[2])
    assert int(res.naccept[0]) == len(G[p + "steps_t"]) - 1 or abs(int(res.naccept[0]) - (len(G[p + "steps_t"]) - 1)) <= 2


def test_unfused_rk45_equals_scipy_controller_per_member():
    """Unfused driver vs the oracle's restated SciPy controller, members with different inputs (so different step
    sequences): identical nfev / accepted / rejected counts, outputs equal to a small fraction of the tolerance."""
    from continuum_robot_b200 import solve_ensemble
    from oracle import beam_oracle as bo

    N = 6
    spec = bo.BeamSpec.uniform(N, elem_type=bo.NONLINEAR)
    fs = bo.ForceSpec(1000.0, True, (0.0, -9.81, 0.0), False)
    b = bo.BeamOracle(spec, fs)
    n = b.n
    par = np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                    spec.wetted_area, spec.drag_coef], axis=1)[None]
    beam = make_gpu_beam(par, np.ones(N, dtype=int), np.array([1] + [0] * N), 1000.0, False)
    amps = np.array([0.5, 2.0, 5.0])
    A = torch.from_numpy(amps).cuda().unsqueeze(-1)
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    te = np.linspace(0.0, 0.003, 7)
    res = solve_ensemble(beam, (0.0, 0.003), torch.zeros(3, 2 * n, dtype=torch.float64, device="cuda"), method="RK45",
                         t_eval=te, rtol=1e-6, atol=1e-9, u=lambda t: A * torch.cos(900.0 * t) * ones)
    assert res.success
    for i, a in enumerate(amps):
        r = bo.rk45_solve(lambda t, x: b.rhs(t, x, a * np.cos(900.0 * t) * np.ones(n)), (0.0, 0.003), np.zeros(2 * n),
                          t_eval=te, rtol=1e-6, atol=1e-9)
        assert (int(res.nfev[i]), int(res.naccept[i]), int(res.nreject[i])) == (r.nfev, r.naccept, r.nreject)
        got = res.y[i].cpu().numpy()
        assert np.all(np.abs(got - r.y) <= 1e-3 * (1e-9 + 1e-6 * np.abs(r.y)))
    assert len({int(v) for v in res.nfev}) > 1


def test_fused_time_varying_inputs_rhs_and_rk4_match_oracle():
    """SinusoidInput + PiecewiseLinearInput (+ a constant part and a tip impulse) fused into crb_rhs / crb_rk4:
    RHS <= 1e-11 and 400 RK4 steps <= 1e-9 against the oracle driven with the same u(t)."""
    from continuum_robot_b200 import PiecewiseLinearInput, SinusoidInput, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(11)
    N, B = 10, 5
    spec = bo.BeamSpec.uniform(N)
    fs = bo.ForceSpec(0.0, False, (0.0, -9.81, 0.0), True)
    b = bo.BeamOracle(spec, fs)
    n = b.n
    par = np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                    spec.wetted_area, spec.drag_coef], axis=1)[None]
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N), 0.0, True)
    amp = rng.standard_normal((B, n))
    knots = np.array([0.0, 0.001, 0.0025, 0.006])
    tab = rng.standard_normal((len(knots), B, n))
    uc = 0.3 * rng.standard_normal((B, n))
    imp = np.linspace(0.1, 0.5, B)
    w, ph = 700.0, 0.4
    parts = [torch.from_numpy(uc).cuda(), TipImpulse(torch.from_numpy(imp).cuda(), duration=0.004),
             SinusoidInput(torch.from_numpy(amp).cuda(), omega=w, phase=ph),
             PiecewiseLinearInput(knots, torch.from_numpy(tab).cuda())]

    def u_ref(i):
        def u(t):
            v = uc[i] + amp[i] * np.sin(w * t + ph) + np.array([np.interp(t, knots, tab[:, i, r]) for r in range(n)])
            if t < 0.004:
                v = v.copy()
                v[-2] += imp[i]
            return v
        return u

    X = np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)
    Xd = torch.from_numpy(X).cuda()
    f = beam.get_dynamic_system()
    for t in (0.0, 0.0017, 0.0031, 0.01):
        got = f(t, Xd, parts).cpu().numpy()
        for i in range(B):
            ref = b.rhs(t, X[i], u_ref(i))
            assert block_err(got[i], ref, n) <= 1e-11
    h, nsteps = 2.0e-5, 400
    rk4_steps(beam, Xd, 0.0, h, nsteps, u=parts)
    got = Xd.cpu().numpy()
    for i in range(B):
        ref = bo.rk4_solve(lambda t, x, i=i: b.rhs(t, x, u_ref(i)), X[i], 0.0, h, nsteps)
        assert block_err(got[i], ref, n) <= 1e-9


def test_prebuilt_system_rejects_extra_inputs():
    from continuum_robot_b200.integrate import rk4_steps

    beam = _beam("sin_lin4")
    X = torch.zeros(2, 2 * beam.n_free, dtype=torch.float64, device="cuda")
    system = beam.make_system(2)
    with pytest.raises(ValueError, match="prebuilt"):
        rk4_steps(beam, X, 0.0, 1e-5, 1, u=torch.zeros(2, beam.n_free, dtype=torch.float64, device="cuda"), system=system)
    # a system built by the 2-slots-per-lane sibling carries another lane layout: refused, not mis-read
    with pytest.raises(ValueError, match="another beam"):
        rk4_steps(beam, X, 0.0, 1e-5, 1, system=beam.with_slots(2).make_system(2))
    rk4_steps(beam, X, 0.0, 1e-5, 1, system=system)
