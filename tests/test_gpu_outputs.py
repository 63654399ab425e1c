"""Lean in-kernel recording: tip trace / node shapes / arbitrary state rows written by the
kernels themselves (crb_system_t.out_sel_inv) must equal the same rows of the full-state recording, for every
kernel family, and the full-state recording of the persistent kernel (bulk stores) must match the oracle.
What the reference's callers read: examples/lqr_control.py:166-183 (tip), examples/example_utilities.py:173-205."""

import numpy as np
import pytest

from helpers import block_err, make_gpu_beam

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _uniform_beam(N, B, nonlinear=False, fluid=0.0, gravity=False, per_member_E=True, seed=3, bc=None):
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(seed)
    spec = bo.BeamSpec.uniform(N, elem_type=bo.NONLINEAR if nonlinear else bo.LINEAR)
    par = np.stack([spec.length, spec.elastic_modulus, spec.moment_inertia, spec.density, spec.cross_area,
                    spec.wetted_area, spec.drag_coef], axis=1)[None]
    if per_member_E:
        par = np.repeat(par, B, axis=0)
        par[:, :, 1] *= np.exp(0.2 * rng.standard_normal((B, N)))
    bcs = np.array([1] + [0] * N) if bc is None else np.asarray(bc)
    beam = make_gpu_beam(par, np.full(N, int(nonlinear)), bcs, fluid, gravity)
    n = beam.n_free
    X0 = np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)
    return beam, torch.from_numpy(X0).cuda(), spec, par


CASES = {
    # name: (N, B, kwargs, method, h, nsteps, extra solve kwargs)
    "persistent_paired": (32, 37, {}, "RK4", 2e-5, 40, {}),
    "persistent_paired_pinned_root": (15, 21, {"bc": [2] + [0] * 15, "per_member_E": True}, "RK4", 2e-5, 20, {}),
    "fast_gravity": (10, 19, {"gravity": True}, "RK4", 2e-5, 30, {}),
    "general_nonlinear_drag": (20, 11, {"nonlinear": True, "fluid": 1000.0}, "RK4", 2e-5, 20, {}),
    "midpoint": (32, 23, {}, "MIDPOINT", 1e-4, 30, {}),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("se", [1, 5])
def test_lean_frames_equal_rows_of_full_frames(case, se):
    from continuum_robot_b200 import output_selection, solve_ensemble

    N, B, kw, method, h, nsteps, extra = CASES[case]
    beam, X0, _, _ = _uniform_beam(N, B, **kw)
    n = beam.n_free
    full = solve_ensemble(beam, (0.0, nsteps * h), X0, method=method, h=h, save_every=se, **extra)
    assert full.y.shape == (B, 2 * n, nsteps // se + 1) and full.rows is None
    for what in ("tip", "shape", "shape_velocity", [0, 2 * n - 1, n, 5]):
        lean = solve_ensemble(beam, (0.0, nsteps * h), X0, method=method, h=h, save_every=se, outputs=what, **extra)
        rows = output_selection(beam, what)
        assert lean.rows == rows and lean.y.shape == (B, len(rows), nsteps // se + 1)
        assert torch.equal(lean.y, full.y[:, rows, :]), (case, what)
        assert torch.equal(lean.x_final, full.x_final)
    assert output_selection(beam, "tip") == [n - 2]  # examples/lqr_control.py:168
    assert output_selection(beam, "shape_velocity") == list(range(n + 1, 2 * n, 3))  # example_utilities.py:196 (Q7)


def test_lean_frames_closed_loop_kernels():
    """Shared-operator (one gain) and dense-operator (one gain per member) LQR rollouts."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse, solve_ensemble

    N, B = 6, 29
    beam, X0, _, _ = _uniform_beam(N, B, gravity=True, per_member_E=False)
    n = beam.n_free
    rng = np.random.default_rng(5)
    g0 = 0.05 * rng.standard_normal((n, 2 * n))
    g0[:, n:] += 5.0 * np.eye(n)  # velocity feedback: a damped, stable closed loop
    gain = torch.from_numpy(g0).cuda()
    imp = TipImpulse(torch.linspace(1.0, 5.0, B, dtype=torch.float64, device="cuda"))
    for g in (gain, gain.unsqueeze(0).repeat(B, 1, 1) * torch.linspace(0.5, 1.5, B, dtype=torch.float64, device="cuda").view(B, 1, 1)):
        ctrl = FullStateLinear(g)
        full = solve_ensemble(beam, (0.0, 100 * 5e-6), X0, method="RK4", h=5e-6, save_every=10, u=imp, controller=ctrl)
        lean = solve_ensemble(beam, (0.0, 100 * 5e-6), X0, method="RK4", h=5e-6, save_every=10, u=imp, controller=ctrl,
                              outputs="tip")
        assert bool(torch.isfinite(full.y).all())
        assert torch.equal(lean.y, full.y[:, [n - 2], :])


def test_lean_outputs_rk45_fused_and_unfused():
    from continuum_robot_b200 import TipImpulse, solve_ensemble

    N, B = 8, 7
    beam, X0, _, _ = _uniform_beam(N, B, nonlinear=True, fluid=1000.0, gravity=True)
    n = beam.n_free
    te = np.linspace(0.0, 0.002, 6)
    imp = TipImpulse(torch.linspace(0.1, 0.4, B, dtype=torch.float64, device="cuda"))
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    for u in (imp, lambda t: 0.2 * torch.sin(800.0 * t) * ones):
        full = solve_ensemble(beam, (0.0, 0.002), X0, method="RK45", t_eval=te, rtol=1e-6, atol=1e-9, u=u)
        lean = solve_ensemble(beam, (0.0, 0.002), X0, method="RK45", t_eval=te, rtol=1e-6, atol=1e-9, u=u, outputs="shape")
        assert lean.y.shape == (B, len(lean.rows), len(te))
        assert torch.equal(lean.y, full.y[:, lean.rows, :])
        assert torch.equal(lean.nfev, full.nfev)


@pytest.mark.parametrize("N,B,bc", [(32, 101, None), (32, 4, None), (13, 50, None), (15, 33, [2] + [0] * 15)])
def test_persistent_kernel_full_recording_matches_oracle(N, B, bc):
    """Full-state frames of the persistent paired kernel leave through shared memory as bulk stores: every frame
    against the NumPy oracle (<= 1e-9), ragged member counts (partial last tile, fewer tiles than resident warps)."""
    from continuum_robot_b200 import solve_ensemble
    from oracle import beam_oracle as bo

    beam, X0, spec, par = _uniform_beam(N, B, bc=bc)
    n = beam.n_free
    h, nsteps, se = 2e-5, 30, 10
    res = solve_ensemble(beam, (0.0, nsteps * h), X0, method="RK4", h=h, save_every=se)
    got = res.y.cpu().numpy()
    x0 = X0.cpu().numpy()
    for i in sorted({0, 1, B // 2, B - 2, B - 1}):
        sp = bo.BeamSpec.uniform(N)
        sp.elastic_modulus = par[i, :, 1].copy()
        if bc is not None:
            sp.bc = np.asarray(bc[:N])
        b = bo.BeamOracle(sp)
        _, Y = bo.rk4_solve(lambda t, x: b.rhs(t, x, np.zeros(n)), x0[i], 0.0, h, nsteps, save_every=se)
        for f in range(nsteps // se):
            assert block_err(got[i, :, f + 1], Y[f], n) <= 1e-9, (i, f)
    assert np.array_equal(got[:, :, 0], x0)


def test_solve_ensemble_step_grid_validation():
    from continuum_robot_b200 import solve_ensemble

    beam, X0, _, _ = _uniform_beam(4, 2)
    with pytest.raises(ValueError, match="whole number"):
        solve_ensemble(beam, (0.0, 1e-5), X0, method="RK4", h=1e-4)  # nsteps would round to 0
    with pytest.raises(ValueError, match="whole number"):
        solve_ensemble(beam, (0.0, 2.5e-5), X0, method="RK4", h=1e-5)  # the interval end is not on the step grid
    res = solve_ensemble(beam, (0.0, 3e-5), X0.cpu(), method="RK4", h=1e-5, save_every=1)  # CPU X0: frame 0 is the device copy
    assert res.y.is_cuda and torch.equal(res.y[:, :, 0], X0)
    with pytest.raises(ValueError, match="unknown output selection"):
        solve_ensemble(beam, (0.0, 3e-5), X0, method="RK4", h=1e-5, outputs="everything")
