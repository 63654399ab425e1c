"""GPU tests of the implicit-midpoint integrator (crb_midpoint, crb_assemble_shifted) against the oracle's
restatement of the rule on the reference's M and K, plus size-independent properties at full size."""

import numpy as np
import pytest

from helpers import block_err, make_gpu_beam

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ensemble(B, N, seed, per_member_mass=False, shared=False):
    from continuum_robot_b200 import ensembles as ens

    e = ens.config3(B, N, seed=seed)
    m = ens.material()
    rng = np.random.default_rng(seed + 1)
    Bp = 1 if shared else B
    par = np.zeros((Bp, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E[:Bp]
    if per_member_mass:
        par[:, :, 3] *= np.exp(0.2 * rng.standard_normal(Bp))[:, None] * np.linspace(1.0, 0.7, N)[None, :]
        par[:, :, 0] *= (1 + 0.1 * rng.random(Bp))[:, None]
    par[:, :, 5:] = 1.0
    return e, par


@pytest.mark.parametrize("N,variant", [(32, "plain"), (32, "impulse"), (32, "force"), (32, "pm"), (16, "shared"),
                                       (12, "impulse"), (6, "force"), (64, "plain"), (3, "plain")])
def test_midpoint_matches_oracle(N, variant):
    """crb_midpoint vs oracle.midpoint_solve with h = 10 x the RK4 stability limit: shared design, per-member
    stiffness, per-member tapered mass, constant force, tip impulse whose window ends inside the run."""
    from continuum_robot_b200 import TipImpulse, midpoint_steps
    from oracle import beam_oracle as bo

    B, h, steps = 41, 2e-4, 60
    e, par = _ensemble(B, N, 5 + N, per_member_mass=(variant == "pm"), shared=(variant == "shared"))
    et, bc = np.zeros(N, dtype=int), np.array([1] + [0] * N)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    rng = np.random.default_rng(N)
    amp = rng.uniform(0.05, 0.5, B)
    uconst = 1e-2 * rng.standard_normal((B, n))
    u = None
    if variant == "impulse":
        u = TipImpulse(torch.from_numpy(amp).cuda(), duration=20.3 * h)
    if variant == "force":
        u = torch.from_numpy(uconst).cuda()
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    X = torch.from_numpy(x0).cuda()
    Y = torch.zeros(steps // 20, B, 2 * n, dtype=torch.float64, device="cuda")
    midpoint_steps(beam, X, 0.0, h, steps, u=u, Y_out=Y, save_every=20)
    got, frames = X.cpu().numpy(), Y.cpu().numpy()
    assert np.isfinite(got).all()
    for i in (0, B // 2, B - 1):
        p = par[min(i, par.shape[0] - 1)]
        spec = bo.BeamSpec(p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4], et, bc, p[:, 5], p[:, 6])
        orc = bo.BeamOracle(spec)
        def uf(t, i=i):
            f = np.zeros(n)
            if variant == "impulse" and t < 20.3 * h:
                f[n - 2] = amp[i]
            if variant == "force":
                f = uconst[i].copy()
            return f
        want, wf = bo.midpoint_solve(orc, uf, x0[i], 0.0, h, steps, save_every=20)
        assert block_err(got[i], want, n) < 1e-9, (i, block_err(got[i], want, n))
        assert max(block_err(frames[k, i], wf[k], n) for k in range(len(wf))) < 1e-9


def test_midpoint_full_size_energy_and_convergence():
    """Config-3 size (65,536 x 32): the rule conserves every member's quadratic energy (checked through
    M and K applied on the host for sampled members), and halving h reduces the distance to the RK4
    solution of the same ensemble four-fold (second order)."""
    from continuum_robot_b200 import midpoint_steps, solve_ensemble
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    B, N = 65536, 32
    e, par = _ensemble(B, N, 1234)
    et, bc = np.zeros(N, dtype=int), np.array([1] + [0] * N)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    x0 = np.concatenate([e.q0, e.v0], axis=1)
    X = torch.from_numpy(x0).cuda()
    midpoint_steps(beam, X, 0.0, 5e-4, 400)  # 0.2 s of simulated time in 400 steps (RK4 would need 10,000)
    got = X.cpu().numpy()
    assert np.isfinite(got).all()
    for i in (0, 4097, B - 1):
        p = par[i]
        orc = bo.BeamOracle(bo.BeamSpec(p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4], et, bc, p[:, 5], p[:, 6]))
        K = bo.dense_stiffness(orc)
        en = lambda x: 0.5 * x[n:] @ orc.M @ x[n:] + 0.5 * x[:n] @ K @ x[:n]
        assert abs(en(got[i]) - en(x0[i])) <= 1e-8 * en(x0[i]), i
    # convergence against RK4 over a short horizon
    T = 2e-4
    Xr = torch.from_numpy(x0).cuda()
    rk4_steps(beam, Xr, 0.0, 1e-6, 200)
    errs = []
    for h in (4e-6, 2e-6):
        Xm = torch.from_numpy(x0).cuda()
        midpoint_steps(beam, Xm, 0.0, h, int(round(T / h)))
        errs.append((Xm - Xr).abs().max().item())
    assert 3.0 < errs[0] / errs[1] < 5.0, errs
    # solve_ensemble front end
    small = make_gpu_beam(par[:64], et, bc)
    res = solve_ensemble(small, (0.0, 1e-2), torch.from_numpy(x0[:64]).cuda(), method="MIDPOINT", h=1e-4, save_every=25)
    assert res.y.shape == (64, 2 * n, 5) and res.success and int(res.nfev[0]) == 100
    Xs = torch.from_numpy(x0[:64]).cuda()
    midpoint_steps(small, Xs, 0.0, 1e-4, 100)
    assert torch.equal(res.y[:, :, -1], Xs)


def test_midpoint_argument_errors():
    from continuum_robot_b200 import midpoint_steps

    N = 8
    e, par = _ensemble(4, N, 3, shared=True)
    nl = make_gpu_beam(par, np.ones(N, dtype=int), np.array([1] + [0] * N))
    X = torch.zeros(4, 2 * nl.n_free, dtype=torch.float64, device="cuda")
    with pytest.raises(ValueError, match="all-linear"):
        midpoint_steps(nl, X, 0.0, 1e-4, 1)
    grav = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N), 0.0, True)
    with pytest.raises(TypeError, match="force-free"):
        midpoint_steps(grav, X, 0.0, 1e-4, 1)
    lin = make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))
    with pytest.raises(ValueError, match="positive"):
        midpoint_steps(lin, X, 0.0, -1.0, 1)


@pytest.mark.parametrize("N,bcs,variant", [
    (10, {0: 1}, "impulse"),          # cantilever with phantom slots (10 active nodes on 4 lanes x 3 slots)
    (32, {0: 2}, "plain"),            # pinned root: constrained DOFs inside the root slot
    (9, {0: 2, 5: 2}, "force"),       # pinned root + interior pin
    (7, {3: 1}, "impulse"),           # free ends, clamped in the middle
    (33, {0: 1, 20: 2}, "force"),
])
def test_midpoint_any_boundary_conditions(N, bcs, variant):
    """Implicit midpoint on plans with constrained DOFs inside active slots / phantom slots (NC kernel variants,
    factors of M + h^2/4 K with identity rows) vs the oracle (<= 1e-9), per-member stiffness."""
    from continuum_robot_b200 import TipImpulse, midpoint_steps
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.dynamic_beam import BatchedDynamicEulerBernoulliBeam
    from oracle import beam_oracle as bo

    rng = np.random.default_rng(40 + N)
    B, h, steps = 19, 1e-4, 40
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0] = m["length"] * (1 + 0.2 * rng.random(N))[None]
    par[:, :, 1] = m["E"] * np.exp(0.2 * rng.standard_normal((B, N)))
    par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["I"], m["rho"], m["A"]
    par[:, :, 5:] = 1.0
    bc = np.zeros(N, dtype=int)
    for k, v in bcs.items():
        bc[k] = v
    et = np.zeros(N, dtype=int)
    beam = BatchedDynamicEulerBernoulliBeam({"params": par, "type": [0] * N, "boundary_condition": [int(b) for b in bc]})
    beam.create_system_func()
    beam.create_input_func()
    n = beam.n_free
    amp = rng.uniform(0.05, 0.5, B)
    dof = int(rng.integers(0, n))
    uconst = 1e-2 * rng.standard_normal((B, n))
    u = None
    if variant == "impulse":
        u = TipImpulse(torch.from_numpy(amp).cuda(), dof=dof, duration=20.3 * h)
    if variant == "force":
        u = torch.from_numpy(uconst).cuda()
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    X = torch.from_numpy(x0).cuda()
    Y = torch.zeros(steps // 20, B, 2 * n, dtype=torch.float64, device="cuda")
    midpoint_steps(beam, X, 0.0, h, steps, u=u, Y_out=Y, save_every=20)
    got, frames = X.cpu().numpy(), Y.cpu().numpy()
    for i in (0, B // 2, B - 1):
        p = par[i]
        orc = bo.BeamOracle(bo.BeamSpec(p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4], et, bc, p[:, 5], p[:, 6]))

        def uf(t, i=i):
            f = np.zeros(n)
            if variant == "impulse" and t < 20.3 * h:
                f[dof] = amp[i]
            if variant == "force":
                f = uconst[i].copy()
            return f

        want, wf = bo.midpoint_solve(orc, uf, x0[i], 0.0, h, steps, save_every=20)
        assert block_err(got[i], want, n) < 1e-9, (i, block_err(got[i], want, n))
        assert max(block_err(frames[k, i], wf[k], n) for k in range(len(wf))) < 1e-9
