"""GPU tests of the drop-in API behaviour (mirrors the reference's tests/test_dynamic_beam.py,
test_functional_composition.py, test_advanced_composition.py; arrays are batched torch tensors)."""

import numpy as np
import pytest

from helpers import block_err, load, make_gpu_beam, oracle_spec, params_array

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
RHS = load("rhs_cases.npz")


def _beam(name="lin4_grav_drag", **kw):
    p = name + "/"
    return make_gpu_beam(params_array(RHS, p), RHS[p + "elem_type"], RHS[p + "bc"], RHS[p + "fluid_density"],
                         RHS[p + "gravity"], RHS[p + "gravity_vector"], **kw)


def test_construction_from_dataframe_and_csv(tmp_path):
    """CSV / DataFrame ingestion with the reference's column names, validation errors and maps."""
    import pandas as pd

    from continuum_robot_b200 import BatchedDynamicEulerBernoulliBeam, ForceParams

    p = "lin4/"
    df = pd.DataFrame({c: RHS[p + c] for c in ("length", "elastic_modulus", "moment_inertia", "density", "cross_area",
                                               "wetted_area", "drag_coef")})
    df["type"] = "linear"
    df["boundary_condition"] = ["FIXED", "NONE", "NONE", "NONE"]
    csv = tmp_path / "beam.csv"
    df.to_csv(csv, index=False, float_format="%.17g")
    for src in (df, str(csv), [df, df]):
        beam = BatchedDynamicEulerBernoulliBeam(src)
        assert beam.n_free == 12 and sorted(beam.constrained_dofs) == [0, 1, 2]
        assert beam.get_state_index(1, "u") == 0 and beam.get_state_index(4, "dphi_dt") == 23
        assert beam.get_state_to_node_param(13) == ("dw_dt", 1)
        assert len(beam.get_state_mapping()) == 24
        with pytest.raises(KeyError):
            beam.get_state_index(0, "u")
        with pytest.raises(RuntimeError, match="not yet created"):
            beam.get_system_func()
        with pytest.raises(RuntimeError, match="must be created first"):
            beam.get_dynamic_system()
    with pytest.raises(ValueError, match="CSV must contain columns"):
        BatchedDynamicEulerBernoulliBeam(df.drop(columns=["density"]))
    with pytest.raises(ValueError, match="Invalid element types"):
        BatchedDynamicEulerBernoulliBeam(df.assign(type="cubic"))
    with pytest.raises(ValueError, match="Invalid boundary conditions"):
        BatchedDynamicEulerBernoulliBeam(df.assign(boundary_condition="WELDED"))
    with pytest.raises(ValueError, match="positive"):
        BatchedDynamicEulerBernoulliBeam(df.assign(length=-1.0))
    # like the reference, CSV rows can only constrain nodes 0..N-1, so node N always stays free (Q6)
    assert BatchedDynamicEulerBernoulliBeam(df.assign(boundary_condition="FIXED")).n_free == 3
    with pytest.raises(FileNotFoundError):
        BatchedDynamicEulerBernoulliBeam(str(tmp_path / "missing.csv"))
    with pytest.raises(ValueError, match="Drag coefficients cannot be negative"):
        BatchedDynamicEulerBernoulliBeam(df.assign(drag_coef=-0.1), ForceParams(fluid_density=1000.0, enable_fluid_effects=True))


def test_system_and_input_functions():
    beam = _beam("lin4")
    n = beam.n_free
    X = torch.from_numpy(RHS["lin4/X"]).cuda()
    U = torch.from_numpy(RHS["lin4/U"]).cuda()
    sysf, inf = beam.get_system_func(), beam.input_func
    total = beam.get_dynamic_system()(0.0, X, U)
    assert torch.allclose(sysf(X) + inf(X, U, 0.0), total, rtol=1e-12, atol=1e-9)
    assert torch.equal(inf(X, U)[:, :n], torch.zeros_like(U))
    assert sysf(X[0]).shape == (2 * n,)  # 1-D call like the reference
    with pytest.raises(ValueError, match="must match position DOFs"):
        inf(X, U[:, :-1])
    with pytest.raises(ValueError):
        inf(X.cpu().numpy(), U)
    with pytest.raises(ValueError):
        beam.get_dynamic_system()(0.0, X[:, :-1], U)
    # callable input u(t)
    f = beam.get_dynamic_system()
    assert torch.equal(f(0.25, X, lambda t: U * t), f(0.25, X, U * 0.25))


def test_registry_toggle_and_user_forces():
    """Built-ins are fused; toggling / unregistering them and adding a torch plug-in take effect
    on the next call (reference tests/test_advanced_composition.py:368-398)."""
    from continuum_robot_b200 import AbstractForce

    name = "nl4_grav_drag"
    beam = _beam(name)
    n = beam.n_free
    X = torch.from_numpy(RHS[name + "/X"]).cuda()
    U = torch.from_numpy(RHS[name + "/U"]).cuda()
    f = beam.get_dynamic_system()
    ref_all = RHS[name + "/Y"]
    assert max(block_err(f(0.3, X, U).cpu().numpy()[i], ref_all[i], n) for i in range(6)) < 1e-11
    drag, grav = beam.force_registry.get_registered_forces()
    assert (drag.fused_kind, grav.fused_kind) == ("drag", "gravity") and len(beam.force_registry) == 2
    grav.set_enabled(False)
    drag.enabled = False
    none = f(0.3, X, U).cpu().numpy()
    ref_none = RHS["nl4/Y"]  # same beam and states without forces? (different random draws) -> use oracle
    from oracle import beam_oracle as bo

    b0 = bo.BeamOracle(oracle_spec(RHS, name + "/"))
    for i in range(6):
        assert block_err(none[i], b0.rhs(0.3, RHS[name + "/X"][i], RHS[name + "/U"][i]), n) < 1e-11
    grav.set_enabled(True)
    bg = bo.BeamOracle(oracle_spec(RHS, name + "/"), bo.ForceSpec(enable_gravity_effects=True))
    got = f(0.3, X, U).cpu().numpy()
    for i in range(6):
        assert block_err(got[i], bg.rhs(0.3, RHS[name + "/X"][i], RHS[name + "/U"][i]), n) < 1e-11

    class Spring(AbstractForce):  # state-aware user force evaluated on the device (unfused path)
        def __init__(self):
            self.on = True

        def compute_forces(self, x, t):
            return -3.0e4 * x[:, :n] - 2.0 * x[:, n:]

        def is_enabled(self):
            return self.on

    sp = Spring()
    beam.force_registry.register(sp)
    got = f(0.3, X, U).cpu().numpy()
    for i in range(6):
        x = RHS[name + "/X"][i]
        ref = bg.system(x, lambda xx, t: bg.builtin_forces(xx) - 3.0e4 * xx[:n] - 2.0 * xx[n:]) + bg.input(RHS[name + "/U"][i])
        assert block_err(got[i], ref, n) < 1e-11
    assert beam.force_registry.unregister(sp) and not beam.force_registry.unregister(sp)
    # a forces_func passed to create_system_func REPLACES the registry (dynamic_beam_model.py:252-254)
    beam.create_system_func(forces_func=lambda x, t: torch.zeros(x.shape[0], n, dtype=x.dtype, device=x.device))
    got = beam.get_dynamic_system()(0.3, X, U).cpu().numpy()
    for i in range(6):
        assert block_err(got[i], b0.rhs(0.3, RHS[name + "/X"][i], RHS[name + "/U"][i]), n) < 1e-11
    with pytest.raises(TypeError):  # NumPy-returning callables are rejected: no CPU fallback
        beam.create_system_func(forces_func=lambda x, t: np.zeros((x.shape[0], n)))
        beam.get_dynamic_system()(0.3, X, U)


def test_builtin_force_objects_compute_forces():
    from oracle import beam_oracle as bo

    name = "mixed5"
    beam = _beam(name)
    b = bo.BeamOracle(oracle_spec(RHS, name + "/"), bo.ForceSpec(800.0, True, RHS[name + "/gravity_vector"], True))
    X = torch.from_numpy(RHS[name + "/X"]).cuda()
    drag, grav = beam.force_registry.get_registered_forces()
    fd = drag.compute_forces(X, 0.0).cpu().numpy()
    fg = grav.compute_forces(X, 0.0).cpu().numpy()
    for i in range(6):
        x = RHS[name + "/X"][i]
        assert np.abs(fd[i] - b.drag(x)).max() <= 1e-14 * max(np.abs(b.drag(x)).max(), 1e-30)
        assert np.abs(fg[i] - b.gravity(x)).max() <= 1e-14 * np.abs(b.gravity(x)).max()


def test_rk4_unfused_user_force_and_trajectory_output():
    """solve_ensemble with a torch plug-in force (one crb_rhs launch per stage) equals the oracle;
    t_eval on the step grid selects frames; per-member nfev/status mirror OdeResult."""
    from continuum_robot_b200 import AbstractForce, solve_ensemble
    from oracle import beam_oracle as bo

    name = "lin4"
    beam = _beam(name)
    n = beam.n_free

    class Damper(AbstractForce):
        def compute_forces(self, x, t):
            return -0.5 * x[:, n:]

        def is_enabled(self):
            return True

    beam.force_registry.register(Damper())
    X0 = torch.from_numpy(RHS[name + "/X"][:3]).cuda()
    h = 2e-5
    res = solve_ensemble(beam, (0.0, 40 * h), X0, method="RK4", h=h, t_eval=[0.0, 20 * h, 40 * h])
    assert res.y.shape == (3, 2 * n, 3) and res.success and int(res.nfev[0]) == 160
    b = bo.BeamOracle(oracle_spec(RHS, name + "/"))
    f = lambda t, x: b.system(x, lambda xx, tt: -0.5 * xx[n:])  # noqa: E731
    for i in range(3):
        ref = bo.rk4_solve(f, RHS[name + "/X"][i], 0.0, h, 40)
        assert block_err(res.y[i, :, 2].cpu().numpy(), ref, n) < 1e-10
    assert torch.equal(res.y[:, :, 0], X0)
    with pytest.raises(ValueError, match="step grid"):
        solve_ensemble(beam, (0.0, 40 * h), X0, method="RK4", h=h, t_eval=[0.5 * h])
    with pytest.raises(ValueError, match="LSODA"):
        solve_ensemble(beam, (0.0, 1e-3), X0, method="LSODA")


def test_rk45_without_t_eval_and_status_fields():
    from continuum_robot_b200 import solve_ensemble

    beam = _beam("nl4_grav_drag")
    n = beam.n_free
    X0 = torch.from_numpy(RHS["nl4_grav_drag/X"]).cuda()
    res = solve_ensemble(beam, (0.0, 2e-3), X0, method="RK45", rtol=1e-6, atol=1e-9)
    assert res.success and res.y.shape == (6, 2 * n, 1) and res.t[-1] == 2e-3
    assert torch.all(res.status == 0) and torch.all(res.nfev > 0)
    assert torch.all(res.nfev == 2 + 6 * (res.naccept + res.nreject))
    assert torch.allclose(res.t_final, torch.full_like(res.t_final, 2e-3))
    with pytest.raises(ValueError, match="not within"):
        solve_ensemble(beam, (0.0, 1e-3), X0, method="RK45", t_eval=[2e-3])


def test_host_pipeline_matches_device_resident_run():
    """HostPipeline (pinned host buffers, chunked copies overlapped with the kernels) is bitwise
    identical to the device-resident fused call, including ragged chunking and per-member sets."""
    from continuum_robot_b200 import HostPipeline, TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps

    B, N = 203, 32
    e = ens.config3(B, N, seed=5)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    for u in (None, TipImpulse(torch.linspace(0.1, 1.0, B, dtype=torch.float64, device="cuda"), duration=1e-3)):
        X = torch.from_numpy(x0).cuda()
        rk4_steps(beam, X, 0.0, e.h, 30, u=u)
        rk4_steps(beam, X, 30 * e.h, e.h, 30, u=u)
        xh = torch.from_numpy(x0.copy()).pin_memory()
        for chunk in (72, 0, 203, 1000):  # ragged chunks, library default (one wave), exact, larger than B
            xh = torch.from_numpy(x0.copy()).pin_memory()
            pipe = HostPipeline(beam, B, chunk_members=chunk, u=u)
            pipe.run(xh, 0.0, e.h, 30)
            pipe.run(xh, 30 * e.h, e.h, 30)  # overlaps the first call chunk by chunk
            pipe.synchronize()
            assert np.array_equal(xh.numpy(), X.cpu().numpy()), chunk
        # stream-ordered completion: wait() makes the current stream see the result
        xh = torch.from_numpy(x0.copy()).pin_memory()
        pipe.run(xh, 0.0, e.h, 30)
        pipe.run(xh, 30 * e.h, e.h, 30)
        pipe.wait()
        torch.cuda.current_stream().synchronize()
        assert np.array_equal(xh.numpy(), X.cpu().numpy())
    with pytest.raises(ValueError, match="pinned"):
        pipe.run(torch.zeros(B, 2 * beam.n_free, dtype=torch.float64), 0.0, e.h, 1)


def test_trajectory_consumers_and_analytic_frequency():
    """Device-side shape / tip extraction mirror the reference helpers (incl. quirk Q7), and the
    linear operator reproduces the analytic first cantilever frequency (example_utilities.py:208-240)."""
    from continuum_robot_b200 import TipImpulse, beam_shapes, cantilever_frequencies, solve_ensemble, tip_displacement
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    N = 16
    par = np.zeros((1, N, 7))
    par[0, :, 0], par[0, :, 1], par[0, :, 2], par[0, :, 3], par[0, :, 4] = m["length"] / 4, m["E"], m["I"], m["rho"], m["A"]
    par[0, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    n = beam.n_free
    X0 = torch.zeros(2, 2 * n, dtype=torch.float64, device="cuda")
    h, steps = 2e-6, 150000
    res = solve_ensemble(beam, (0.0, steps * h), X0, method="RK4", h=h, save_every=250,
                         u=TipImpulse(torch.tensor([0.1, 0.2], dtype=torch.float64, device="cuda"), duration=2e-3))
    tip = tip_displacement(res.y)
    assert tip.shape == (2, len(res.t)) and torch.allclose(tip[1], 2 * tip[0], rtol=1e-9, atol=1e-15)  # linearity
    x, yy = beam_shapes(res.y, N, m["length"] / 4)
    assert x.shape == yy.shape == (2, len(res.t), N + 1) and torch.all(yy[:, :, 0] == 0)
    assert torch.equal(yy[:, :, 1:], res.y[:, n + 1 :: 3, :].permute(0, 2, 1))  # velocities, like the reference
    _, yd = beam_shapes(res.y, N, m["length"] / 4, displacements=True)
    assert torch.equal(yd[:, :, -1], tip)
    # dominant frequency of the free response after the impulse vs the analytic first mode
    sig = tip[0].cpu().numpy()[res.t > 2e-3]
    dt = 250 * h
    spec = np.abs(np.fft.rfft((sig - sig.mean()) * np.hanning(len(sig))))
    f_peak = np.fft.rfftfreq(len(sig), dt)[spec.argmax()]
    f1 = cantilever_frequencies(N * m["length"] / 4, m["E"], m["I"], m["rho"], m["A"])[0]
    assert abs(f_peak - f1) <= 1.5 / (len(sig) * dt), (f_peak, f1)
