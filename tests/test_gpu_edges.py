"""GPU edge cases: extreme sizes, ragged / single-member ensembles, argument errors, failure status."""

import numpy as np
import pytest

from helpers import block_err, make_gpu_beam

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _uniform_par(N, B=1, seed=0):
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    rng = np.random.default_rng(seed)
    par = np.zeros((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = m["E"] * np.exp(0.1 * rng.standard_normal((B, N)))
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    return par


@pytest.mark.parametrize("N,B,nl", [(1, 1, False), (1, 3, True), (2, 1, True), (128, 2, False), (128, 1, True), (97, 5, True)])
def test_extreme_sizes_match_oracle(N, B, nl):
    """Single element, single member, the 128-node limit of one lane group, phantom-padded sizes."""
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    par = _uniform_par(N, B, seed=N)
    et = np.full(N, 1 if nl else 0)
    beam = make_gpu_beam(par, et, np.array([1] + [0] * N), 1000.0 if nl else 0.0, nl)
    n = beam.n_free
    rng = np.random.default_rng(1)
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    X = torch.from_numpy(x0).cuda()
    h, steps = 5e-6, 6
    rk4_steps(beam, X, 0.0, h, steps)
    got = X.cpu().numpy()
    i = B - 1
    spec = bo.BeamSpec.uniform(N, elem_type=bo.NONLINEAR if nl else bo.LINEAR)
    spec.elastic_modulus = par[i, :, 1].copy()
    b = bo.BeamOracle(spec, bo.ForceSpec(1000.0, nl, (0.0, -9.81, 0.0), nl))
    ref = bo.rk4_solve(lambda t, x: b.rhs(t, x, np.zeros(n)), x0[i], 0.0, h, steps)
    assert block_err(got[i], ref, n) < 1e-9


def test_size_and_argument_errors():
    from continuum_robot_b200 import BatchedDynamicEulerBernoulliBeam, solve_ensemble
    from continuum_robot_b200.integrate import rk4_steps

    with pytest.raises(ValueError, match="limit"):
        make_gpu_beam(_uniform_par(129), np.zeros(129, dtype=int), np.array([1] + [0] * 129))
    beam = make_gpu_beam(_uniform_par(4, 3), np.zeros(4, dtype=int), np.array([1, 0, 0, 0, 0]))
    n = beam.n_free
    X = torch.zeros(3, 2 * n, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError, match="positive and finite"):
        rk4_steps(beam, X, 0.0, -1.0, 3)
    with pytest.raises(ValueError, match="parameter sets"):
        rk4_steps(beam, torch.zeros(5, 2 * n, dtype=torch.float64, device="cuda"), 0.0, 1e-5, 1)
    before = X.clone()
    rk4_steps(beam, X, 0.0, 1e-5, 0)  # zero steps is a no-op
    assert torch.equal(X, before)
    with pytest.raises(ValueError, match="increasing"):
        solve_ensemble(beam, (1.0, 0.0), X, method="RK4", h=1e-5)
    with pytest.raises(ValueError, match="rtol"):
        solve_ensemble(beam, (0.0, 1e-4), X, method="RK45", rtol=0.0)
    with pytest.raises(ValueError, match="shape"):
        solve_ensemble(beam, (0.0, 1e-4), X[:, :-1], method="RK45")
    with pytest.raises(RuntimeError, match="CUDA devices only"):
        BatchedDynamicEulerBernoulliBeam({"params": _uniform_par(4), "type": ["linear"] * 4}, device="cpu")


def test_rk45_reports_failure_per_member_instead_of_raising():
    """A member whose state is non-finite fails alone with status -1 (SciPy: 'Required step size is
    less than spacing between numbers'); the other members finish normally."""
    from continuum_robot_b200 import solve_ensemble

    beam = make_gpu_beam(_uniform_par(4, 4), np.ones(4, dtype=int), np.array([1, 0, 0, 0, 0]), 1000.0, True)
    n = beam.n_free
    X0 = torch.zeros(4, 2 * n, dtype=torch.float64, device="cuda")
    X0[2, 1] = float("nan")
    res = solve_ensemble(beam, (0.0, 1e-3), X0, method="RK45", rtol=1e-6, atol=1e-9)
    st = res.status.cpu().numpy()
    assert not res.success and st[2] == -1 and np.all(st[[0, 1, 3]] == 0)
    assert "step size" in res.message
    assert torch.isfinite(res.y[[0, 1, 3]]).all()


def test_rk45_attempt_budget_and_resume():
    """max_attempts bounds one launch (status 1); relaunching from (x, t, h) reaches t_bound."""
    from continuum_robot_b200 import solve_ensemble

    beam = make_gpu_beam(_uniform_par(6, 2), np.zeros(6, dtype=int), np.array([1] + [0] * 6), 0.0, True)
    n = beam.n_free
    X0 = torch.zeros(2, 2 * n, dtype=torch.float64, device="cuda")
    full = solve_ensemble(beam, (0.0, 2e-3), X0, method="RK45", rtol=1e-6, atol=1e-9)
    part = solve_ensemble(beam, (0.0, 2e-3), X0, method="RK45", rtol=1e-6, atol=1e-9, max_attempts=5)
    assert not part.success and torch.all(part.status == 1) and torch.all(part.t_final < 2e-3)
    assert torch.all(part.naccept + part.nreject == 5)
    assert full.success and torch.all(full.naccept + full.nreject > 5)


@pytest.mark.parametrize("save_every", [0, 7])
def test_impulse_window_split_is_exact(save_every):
    """crb_rk4 runs the steps after the impulse window as the input-free system (separate launch of the
    cheaper kernel variant).  The result must equal stepping launch by launch across the window end,
    where every launch is decided on its own (bitwise for the states saved on the way and at the end)."""
    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    B, N, h, steps = 67, 32, 2e-5, 84
    e = ens.config3(B, N, seed=21)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5:] = 1.0
    beam = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    n = beam.n_free
    amp = np.linspace(0.1, 1.0, B)
    imp = TipImpulse(torch.from_numpy(amp).cuda(), duration=30.5 * h)  # window ends inside step 30
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    X = torch.from_numpy(x0).cuda()
    nfr = steps // save_every if save_every else 0
    Y = torch.zeros(nfr, B, 2 * n, dtype=torch.float64, device="cuda") if nfr else None
    rk4_steps(beam, X, 0.0, h, steps, u=imp, Y_out=Y, save_every=save_every)
    # reference: one launch per step with the impulse object (each launch inside or outside the window)
    Xs = torch.from_numpy(x0).cuda()
    frames = []
    for k in range(steps):
        rk4_steps(beam, Xs, k * h, h, 1, u=imp)
        if save_every and (k + 1) % save_every == 0:
            frames.append(Xs.clone())
    assert max(block_err(X.cpu().numpy()[i], Xs.cpu().numpy()[i], n) for i in range(B)) < 1e-13
    if nfr:
        got, want = Y.cpu().numpy(), torch.stack(frames).cpu().numpy()
        assert max(block_err(got[f, i], want[f, i], n) for f in range(nfr) for i in range(0, B, 11)) < 1e-13
    # and the CPU oracle
    i = 5
    spec = bo.BeamSpec.uniform(N)
    spec.elastic_modulus = e.E[i].copy()
    orc = bo.BeamOracle(spec)
    def f(t, x):
        u = np.zeros(n)
        if t < 30.5 * h:
            u[n - 2] = amp[i]
        return orc.rhs(t, x, u)
    want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
    assert block_err(X.cpu().numpy()[i], want, n) < 1e-9


@pytest.mark.parametrize("N,bcs", [
    (32, {0: 2}),            # config-3 shape on a PINNED root: untrimmed root slot with two constrained DOFs
    (10, {0: 1}),            # cantilever whose 10 active nodes leave 2 phantom slots (m = 3, g = 4)
    (9, {0: 2, 5: 2}),       # pinned root + interior pin
    (7, {3: 1}),             # free ends, clamped in the middle
    (33, {0: 1, 20: 2}),     # 33 active nodes... one lane group of 16 x 3 slots with phantoms + interior pin
    (5, {0: 1, 5: 2}),       # propped cantilever (node N pinned through the bc array)
])
def test_paired_linear_kernel_any_boundary_conditions(N, bcs):
    """All-linear force-free beams with constrained DOFs inside active slots and / or phantom slots run on the NC
    variants of the paired fast kernel (reduced-index state I/O, masked right-hand sides): identical physics to
    the general kernel (<= 1e-11) and to the oracle (<= 1e-9), without input, with a constant force, with an
    impulse and with both; saved frames included."""
    import torch

    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from helpers import block_err, make_gpu_beam
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(900 + N)
    B, h, steps = 21, 2e-6, 90
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0] = m["length"] * (1 + 0.2 * rng.random(N))[None]
    par[:, :, 1] = m["E"] * np.exp(0.2 * rng.standard_normal((B, N)))
    par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["I"], m["rho"] * (1 + 0.3 * rng.random(N))[None], m["A"]
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    bc = np.zeros(N + 1, dtype=int)
    for k, v in bcs.items():
        bc[k] = v
    from continuum_robot_b200.dynamic_beam import BatchedDynamicEulerBernoulliBeam

    beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": [0] * N, "boundary_condition": [int(b) for b in bc[:N]]})
    if bc[N] != 0:  # node N is not reachable through the CSV; nothing to test beyond the rows
        bc[N] = 0
    beam.create_system_func()
    beam.create_input_func()
    n = beam.n_free
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    U = 0.05 * rng.standard_normal((B, n))
    amp = rng.uniform(0.05, 0.3, B)
    dof = int(rng.integers(0, n))
    for uc, imp in ((False, False), (True, False), (False, True), (True, True)):
        def run(general):
            beam.force_general_kernels = general
            X = torch.from_numpy(x0).cuda()
            Y = torch.zeros(steps // 30, B, 2 * n, dtype=torch.float64, device="cuda")
            system = beam.make_system(B, u_const=torch.from_numpy(U).cuda() if uc else None,
                                      impulse=TipImpulse(torch.from_numpy(amp).cuda(), dof=dof, duration=40.5 * h) if imp else None)
            rk4_steps(beam, X, 0.0, h, steps, system=system, Y_out=Y, save_every=30)
            beam.force_general_kernels = False
            return X.cpu().numpy(), Y.cpu().numpy()
        fast, yf = run(False)
        gen, yg = run(True)
        assert max(block_err(fast[i], gen[i], n) for i in range(B)) < 1e-11, (uc, imp)
        assert np.abs(yf - yg).max() <= 1e-11 * np.abs(yg).max()
        assert np.array_equal(yf[-1], fast)
        for i in (0, B - 1):
            spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], np.zeros(N, dtype=int),
                               bc[:N], par[i, :, 5], par[i, :, 6])
            b = bo.BeamOracle(spec)

            def f(t, x, i=i):
                u = U[i].copy() if uc else np.zeros(n)
                if imp and t < 40.5 * h:
                    u[dof] += amp[i]
                return b.rhs(t, x, u)

            want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
            assert block_err(fast[i], want, n) < 1e-9, (uc, imp, i)


@pytest.mark.parametrize("N,B,with_imp", [(10, 33, True), (32, 19, False), (6, 40, True), (13, 5, True), (64, 3, False)])
def test_fast_kernel_with_gravity_matches_general_kernel_and_oracle(N, B, with_imp):
    """Linear cantilevers under (tilted) gravity with per-member stiffness -- config 1 as an ensemble -- run on the
    stage-by-stage fast kernel with the slot-space gravity term; vs the general kernel (<= 1e-11) and the oracle
    (<= 1e-9), with and without the tip impulse, element counts with and without phantom slots."""
    import torch

    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    par = _uniform_par(N, B, seed=70 + N)
    et, bc = np.zeros(N, dtype=int), np.array([1] + [0] * N)
    gvec = (1.3, -9.81, 0.0)
    beam = make_gpu_beam(par, et, bc, 0.0, True, gvec)
    n = beam.n_free
    rng = np.random.default_rng(N)
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    amp = rng.uniform(0.05, 0.5, B)
    h, steps = 2e-6, 150
    imp = TipImpulse(torch.from_numpy(amp).cuda(), duration=60.5 * h) if with_imp else None
    out = []
    for general in (False, True):
        beam.force_general_kernels = general
        X = torch.from_numpy(x0).cuda()
        rk4_steps(beam, X, 0.0, h, steps, u=imp)
        out.append(X.cpu().numpy())
    beam.force_general_kernels = False
    assert max(block_err(out[0][i], out[1][i], n) for i in range(B)) < 1e-11
    for i in (0, B - 1):
        p = par[i]
        orc = bo.BeamOracle(bo.BeamSpec(p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4], et, bc[:N], p[:, 5], p[:, 6]),
                            bo.ForceSpec(0.0, False, gvec, True))

        def f(t, x, i=i):
            u = np.zeros(n)
            if with_imp and t < 60.5 * h:
                u[n - 2] = amp[i]
            return orc.rhs(t, x, u)

        want = bo.rk4_solve(f, x0[i], 0.0, h, steps)
        assert block_err(out[0][i], want, n) < 1e-9
