"""Batched LQR synthesis on the device: crb_dense_matrices_batched + crb_lqr_gains through
the C ABI, against the oracle's restatement of control/linear_quadratic_regulator.py:84-191 (SciPy CARE, and
its Newton-Kleinman-refined fixed point), and the rollout with one gain per member against the oracle."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import block_err, load, make_gpu_beam, params_array  # noqa: E402

# Floating-point tolerances (relative, max-norm over the gain matrix).  The CARE solution is unique; what differs
# between solvers is rounding.  Measured in the build container (NumPy prototype of the device algorithm):
# SciPy's solve_continuous_are is 0.4-2.2e-8 away from the Newton-Kleinman fixed point on these beams, the sign
# function + one correction pass 0.5-3e-10.
TOL_VS_SCIPY = 2e-7
TOL_VS_REFINED = 5e-9


def design_ensemble(B, N, seed, bc0=1, vary_mass=True):
    from continuum_robot_b200 import ensembles as ens

    rng = np.random.default_rng(seed)
    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0] = m["length"] * (1 + (0.1 * rng.random((B, N)) if vary_mass else 0.0))
    par[:, :, 1] = m["E"] * np.exp(0.3 * rng.standard_normal((B, 1)))
    par[:, :, 2] = m["I"]
    par[:, :, 3] = m["rho"] * (np.exp(0.2 * rng.standard_normal((B, 1))) if vary_mass else 1.0)
    par[:, :, 4] = m["A"]
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    bc = np.array([bc0] + [0] * N)
    return par, np.zeros(N, dtype=int), bc


def weights(n, dev="cuda"):
    Q = np.diag(np.r_[100.0 * np.ones(n), 10.0 * np.ones(n)])  # examples/lqr_control.py:61-66
    R = np.eye(n)
    return Q, R, torch.from_numpy(Q).to(dev), torch.from_numpy(R).to(dev)


@pytest.mark.parametrize("N,bc0", [(6, 1), (5, 2), (1, 1), (13, 1)])
def test_dense_matrices_batched_match_host(N, bc0):
    """Device A / B build = the host crb_dense_matrices of every member (same element formulas; <= 2 ulp)."""
    par, et, bc = design_ensemble(11, N, seed=N, bc0=bc0)
    beam = make_gpu_beam(par, et, bc)
    M, K = beam.dense_matrices()
    assert tuple(M.shape) == (11, beam.n_free, beam.n_free)
    M, K = M.cpu().numpy(), K.cpu().numpy()
    for i in range(11):
        Mh, Kh = beam.beam_model.get_mass_matrix(i), beam.beam_model.get_stiffness_matrix(i)
        assert np.abs(M[i] - Mh).max() <= 4e-16 * np.abs(Mh).max()
        assert np.abs(K[i] - Kh).max() <= 4e-16 * np.abs(Kh).max()


@pytest.mark.parametrize("N,B,bc0", [(6, 24, 1), (3, 9, 1), (1, 5, 1), (10, 6, 1), (4, 7, 2)])
def test_lqr_gains_match_oracle(N, B, bc0):
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator
    from oracle import beam_oracle as bo

    par, et, bc = design_ensemble(B, N, seed=10 + N, bc0=bc0)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    Q, R, Qd, Rd = weights(n)
    Md, Kd = beam.dense_matrices()
    lqr = BatchedLinearQuadraticRegulator(Kd, Md, Qd, Rd)
    gain = lqr.compute_gain_matrix()
    assert tuple(gain.shape) == (B, n, 2 * n)
    assert int(lqr.status.abs().sum()) == 0
    G, S = gain.cpu().numpy(), lqr.get_S().cpu().numpy()
    Mh, Kh = Md.cpu().numpy(), Kd.cpu().numpy()
    worst_s = worst_r = 0.0
    for i in range(B):
        ks = bo.lqr_gain(Kh[i], Mh[i], Q, R)
        kr = bo.lqr_gain_refined(Kh[i], Mh[i], Q, R)
        worst_s = max(worst_s, np.abs(G[i] - ks).max() / np.abs(ks).max())
        worst_r = max(worst_r, np.abs(G[i] - kr).max() / np.abs(kr).max())
        assert np.abs(S[i] - S[i].T).max() <= 1e-13 * np.abs(S[i]).max()
        A, Bm = bo.lqr_matrices(Kh[i], Mh[i])
        assert np.linalg.eigvals(A - Bm @ G[i]).real.max() < 0  # linear_quadratic_regulator.py:185-189
    assert worst_s < TOL_VS_SCIPY, worst_s
    assert worst_r < TOL_VS_REFINED, worst_r


def test_lqr_gain_of_config5_design_matches_golden_gain():
    """The gain stored with the config-5 golden trajectories (SciPy CARE on the reference's K and M)."""
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator

    g = load("cfg5_samples.npz")
    n = g["K_beam"].shape[0]
    _, _, Qd, Rd = weights(n)
    lqr = BatchedLinearQuadraticRegulator(torch.from_numpy(g["K_beam"]).cuda(), torch.from_numpy(g["M_beam"]).cuda(), Qd, Rd)
    K = lqr.compute_gain_matrix()[0].cpu().numpy()
    assert np.abs(K - g["gain"]).max() <= TOL_VS_SCIPY * np.abs(g["gain"]).max()
    assert float(lqr.residual[0]) < 1e-6


@pytest.mark.parametrize("N,bc0", [(6, 1), (5, 2), (12, 1)])
def test_lqr_reducible_designs_are_solved_group_by_group(N, bc0):
    """Axial and bending DOFs of a straight beam are not coupled by M, K or the diagonal weights of
    examples/lqr_control.py:61-66: the host class solves the two groups as separate, smaller problems
    (`decouple=True`, default).  Same gains and S as the one 4n x 4n solve to solver accuracy, exact zeros in the
    cross blocks (as SciPy's CARE has them), same closeness to the oracle; a Q that couples the groups falls back."""
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator
    from oracle import beam_oracle as bo

    par, et, bc = design_ensemble(5, N, seed=40 + N, bc0=bc0)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    Q, R, Qd, Rd = weights(n)
    Md, Kd = beam.dense_matrices()
    split = BatchedLinearQuadraticRegulator(Kd, Md, Qd, Rd)
    groups = split._components()
    assert len(groups) == 2 and sorted(int(g.numel()) for g in groups) == sorted([N, n - N])
    whole = BatchedLinearQuadraticRegulator(Kd, Md, Qd, Rd, decouple=False)
    Gs, Gw = split.compute_gain_matrix().cpu().numpy(), whole.compute_gain_matrix().cpu().numpy()
    Ss, Sw = split.get_S().cpu().numpy(), whole.get_S().cpu().numpy()
    assert np.abs(Gs - Gw).max() <= TOL_VS_REFINED * np.abs(Gw).max()
    assert np.abs(Ss - Sw).max() <= 1e-9 * np.abs(Sw).max()
    ax = np.zeros(n, dtype=bool)
    ax[groups[0].cpu().numpy() if groups[0].numel() == N else groups[1].cpu().numpy()] = True
    cross = ax[:, None] != ax[None, :]
    assert not Gs[:, np.concatenate([cross, cross], axis=1)].any()  # exact zeros, e.g. for the tile-skipping rollout kernel
    assert float(split.residual.max()) < 1e-6 and int(split.status.abs().sum()) == 0
    for i in range(5):
        kr = bo.lqr_gain_refined(Kd[i].cpu().numpy(), Md[i].cpu().numpy(), Q, R)
        assert np.abs(Gs[i] - kr).max() <= TOL_VS_REFINED * np.abs(kr).max()
    Qc = Qd.clone()
    Qc[0, 1] = Qc[1, 0] = 1.0  # couples u and w of the first node: one group, one solve
    assert len(BatchedLinearQuadraticRegulator(Kd, Md, Qc, Rd)._components()) == 1


def test_lqr_shared_mass_and_refinement_passes():
    """One mass matrix shared by the ensemble (m_shared), per-member stiffness; 0 / 1 / 2 correction passes."""
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator
    from oracle import beam_oracle as bo

    par, et, bc = design_ensemble(8, 6, seed=5, vary_mass=False)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    Q, R, Qd, Rd = weights(n)
    Md, Kd = beam.dense_matrices()
    errs = []
    for passes in (0, 1, 2):
        G = BatchedLinearQuadraticRegulator(Kd, Md[0], Qd, Rd, refine_passes=passes).compute_gain_matrix().cpu().numpy()
        kr = [bo.lqr_gain_refined(Kd[i].cpu().numpy(), Md[0].cpu().numpy(), Q, R) for i in range(8)]
        errs.append(max(np.abs(G[i] - kr[i]).max() / np.abs(kr[i]).max() for i in range(8)))
    assert errs[0] < 1e-3 and errs[1] < TOL_VS_REFINED and errs[2] < TOL_VS_REFINED, errs


def test_lqr_failures_raise_like_the_reference():
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator

    par, et, bc = design_ensemble(3, 4, seed=2)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    _, _, Qd, Rd = weights(n)
    Md, Kd = beam.dense_matrices()
    Mbad = Md.clone()
    Mbad[1] = 0.0
    with pytest.raises(ValueError, match="Mass matrix is singular"):  # linear_quadratic_regulator.py:100-104
        BatchedLinearQuadraticRegulator(Kd, Mbad, Qd, Rd).compute_gain_matrix()
    with pytest.raises(ValueError, match="LQR"):  # no stabilising solution: :182-189
        BatchedLinearQuadraticRegulator(Kd, Md, -1e14 * Qd, Rd).compute_gain_matrix()
    with pytest.raises(ValueError, match="Q matrix dimension"):
        BatchedLinearQuadraticRegulator(Kd, Md, Qd[:n, :n], Rd)
    with pytest.raises(ValueError, match="R matrix dimension"):
        BatchedLinearQuadraticRegulator(Kd, Md, Qd, Qd)
    with pytest.raises(TypeError):
        BatchedLinearQuadraticRegulator(Kd.cpu(), Md, Qd, Rd)
    big = torch.eye(97, dtype=torch.float64, device="cuda") + 0.01  # one coupled group of 97 DOFs
    with pytest.raises(ValueError, match="exceed the limit"):
        BatchedLinearQuadraticRegulator(big, big, torch.eye(194, dtype=torch.float64, device="cuda"), big).compute_gain_matrix()


@pytest.mark.parametrize("N,gravity", [(6, True), (4, False)])
def test_rollout_with_per_member_gains_matches_oracle(N, gravity):
    """Design ensemble end to end: device A / B build -> device Riccati -> closed-loop RK4 rollout with ONE GAIN PER
    MEMBER (crb_system_t.gain_stride), against the oracle stepping the same closed loop with the same gains
    (<= 1e-9 block inf-norm)."""
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator, FullStateLinear, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    B, h, steps = 13, 5e-6, 400
    par, et, bc = design_ensemble(B, N, seed=77 + N)
    beam = make_gpu_beam(par, et, bc, 0.0, gravity)
    n = beam.n_free
    _, _, Qd, Rd = weights(n)
    Md, Kd = beam.dense_matrices()
    gain = BatchedLinearQuadraticRegulator(Kd, Md, Qd, Rd).compute_gain_matrix()
    rng = np.random.default_rng(3)
    amp = rng.uniform(1.0, 20.0, B)
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    X = torch.from_numpy(x0).cuda()
    rk4_steps(beam, X, 0.0, h, steps, u=TipImpulse(torch.from_numpy(amp).cuda(), duration=1e-3), controller=FullStateLinear(gain))
    got, Gh = X.cpu().numpy(), gain.cpu().numpy()
    worst = 0.0
    for i in range(B):
        spec = bo.BeamSpec(par[i, :, 0], par[i, :, 1], par[i, :, 2], par[i, :, 3], par[i, :, 4], et, bc[:N], par[i, :, 5], par[i, :, 6])
        ob = bo.BeamOracle(spec, bo.ForceSpec(enable_gravity_effects=gravity))

        def f(t, x, i=i, ob=ob):
            u = bo.full_state_feedback(Gh[i], x, np.zeros(2 * n))
            if t < 1e-3:
                u = u.copy()
                u[n - 2] += amp[i]
            return ob.rhs(t, x, u)

        ref = bo.rk4_solve(f, x0[i], 0.0, h, steps)
        ref = ref[-1] if isinstance(ref, tuple) else ref
        worst = max(worst, block_err(got[i], np.asarray(ref).reshape(-1)[-2 * n:], n))
    assert worst < 1e-9, worst
    # the per-member feedback really is per member: swapping in member 0's gain for everyone changes the result
    X2 = torch.from_numpy(x0).cuda()
    rk4_steps(beam, X2, 0.0, h, steps, u=TipImpulse(torch.from_numpy(amp).cuda(), duration=1e-3), controller=FullStateLinear(gain[0].contiguous()))
    assert (X2[1:] - X[1:]).abs().max() > 1e-9


@pytest.mark.parametrize("slots", [0, 2])
def test_staged_per_member_gains_equal_the_global_memory_path(slots):
    """The one-warp-per-block kernel with the members' gains staged in shared memory gives the same trajectories as
    the kernel that re-reads them from global memory (same arithmetic order: bitwise), ragged member count."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps

    B, N, h, steps = 37, 6, 5e-6, 120
    par, et, bc = design_ensemble(B, N, seed=123)
    beam = make_gpu_beam(par, et, bc, 0.0, True)
    if slots:
        beam = beam.with_slots(slots)
    n = beam.n_free
    rng = np.random.default_rng(1)
    gain = torch.from_numpy(np.concatenate([50.0 * rng.standard_normal((B, n, n)), 0.05 * rng.standard_normal((B, n, n))], axis=2)).cuda()
    ref = torch.from_numpy(1e-3 * rng.standard_normal(2 * n)).cuda()
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    amp = torch.from_numpy(rng.uniform(1.0, 5.0, B)).cuda()
    out = []
    beam.use_member_operators = False  # banded kernels only (the dense per-member operator path is tested below)
    for staged_off in (False, True):
        beam.force_staged_kernels = staged_off
        X = torch.from_numpy(x0).cuda()
        rk4_steps(beam, X, 0.0, h, steps, u=TipImpulse(amp, duration=3e-4), controller=FullStateLinear(gain, reference=ref))
        out.append(X.cpu().numpy())
    beam.force_staged_kernels = False
    assert np.isfinite(out[0]).all()
    assert np.array_equal(out[0], out[1])
    # dense per-member closed-loop operators (crb_member_operators + crb_rk4_dense_kernel): same physics, different
    # arithmetic (dense M^-1 instead of the banded factorisation)
    beam.use_member_operators = True
    X = torch.from_numpy(x0).cuda()
    Y = torch.zeros(steps // 40, B, 2 * n, dtype=torch.float64, device="cuda")
    rk4_steps(beam, X, 0.0, h, steps, u=TipImpulse(amp, duration=3e-4), controller=FullStateLinear(gain, reference=ref),
              Y_out=Y, save_every=40)
    dense = X.cpu().numpy()
    assert max(block_err(dense[i], out[0][i], n) for i in range(B)) < 1e-10
    assert np.array_equal(Y[-1].cpu().numpy(), dense)


def test_host_pipeline_with_per_member_gains():
    """Host-resident state through crb_rk4_host: the chunked pipeline slices the per-member gains together with the
    per-member factor sets (crb_system_slice); bitwise equal to the device-resident rollout."""
    from continuum_robot_b200 import FullStateLinear, HostPipeline, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps

    B, N, h, steps = 83, 6, 5e-6, 60
    par, et, bc = design_ensemble(B, N, seed=9)
    beam = make_gpu_beam(par, et, bc, 0.0, True)
    n = beam.n_free
    rng = np.random.default_rng(2)
    gain = torch.from_numpy(np.concatenate([50.0 * rng.standard_normal((B, n, n)), 0.05 * rng.standard_normal((B, n, n))], axis=2)).cuda()
    amp = torch.from_numpy(rng.uniform(1.0, 5.0, B)).cuda()
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    ctrl, imp = FullStateLinear(gain), TipImpulse(amp, duration=1e-4)
    X = torch.from_numpy(x0).cuda()
    rk4_steps(beam, X, 0.0, h, steps, u=imp, controller=ctrl)
    xh = torch.from_numpy(x0.copy()).pin_memory()
    pipe = HostPipeline(beam, B, chunk_members=24, u=imp, controller=ctrl)
    pipe.run(xh, 0.0, h, steps)
    pipe.synchronize()
    assert np.array_equal(xh.numpy(), X.cpu().numpy())


@pytest.mark.parametrize("N,bc0,gravity,with_ref", [(1, 1, True, False), (3, 1, False, True), (4, 2, True, True), (8, 1, True, False),
                                                    (10, 1, False, False), (10, 2, True, True)])
def test_dense_member_operator_path_matches_oracle(N, bc0, gravity, with_ref):
    """crb_member_operators + crb_rk4_dense_kernel (one dense closed-loop operator per member) for n = 3 ... 31 free
    DOFs, FIXED and PINNED roots, gravity on / off, non-zero reference, per-member mass and stiffness: vs the oracle
    stepping the same closed loops (<= 1e-9), saved frames included."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps
    from oracle import beam_oracle as bo

    B, h, steps = 11, 2e-6, 150
    par, et, bc = design_ensemble(B, N, seed=300 + N, bc0=bc0)
    gvec = (0.7, -9.81, 0.0)
    beam = make_gpu_beam(par, et, bc, 0.0, gravity, gvec)
    n = beam.n_free
    rng = np.random.default_rng(N)
    gain = np.concatenate([50.0 * rng.standard_normal((B, n, n)), 0.05 * rng.standard_normal((B, n, n))], axis=2)
    ref = 1e-3 * rng.standard_normal(2 * n) if with_ref else None
    amp = rng.uniform(1.0, 5.0, B)
    x0 = np.concatenate([1e-4 * rng.standard_normal((B, n)), 1e-2 * rng.standard_normal((B, n))], axis=1)
    ctrl = FullStateLinear(torch.from_numpy(gain).cuda(), reference=torch.from_numpy(ref).cuda() if with_ref else None)
    X = torch.from_numpy(x0).cuda()
    Y = torch.zeros(steps // 50, B, 2 * n, dtype=torch.float64, device="cuda")
    rk4_steps(beam, X, 0.0, h, steps, u=TipImpulse(torch.from_numpy(amp).cuda(), duration=80.5 * h), controller=ctrl,
              Y_out=Y, save_every=50)
    got, frames = X.cpu().numpy(), Y.cpu().numpy()
    for i in (0, B // 2, B - 1):
        p = par[i]
        ob = bo.BeamOracle(bo.BeamSpec(p[:, 0], p[:, 1], p[:, 2], p[:, 3], p[:, 4], et, bc[:N], p[:, 5], p[:, 6]),
                           bo.ForceSpec(0.0, False, gvec, gravity))

        def f(t, x, i=i, ob=ob):
            u = bo.full_state_feedback(gain[i], x, ref if with_ref else np.zeros(2 * n)).copy()
            if t < 80.5 * h:
                u[n - 2] += amp[i]
            return ob.rhs(t, x, u)

        want, wf = bo.rk4_solve(f, x0[i], 0.0, h, steps, save_every=50)
        assert block_err(got[i], want, n) < 1e-9, (i, block_err(got[i], want, n))
        assert max(block_err(frames[k, i], wf[k], n) for k in range(len(wf))) < 1e-9
    # the dense path was the one that ran: switching it off changes the rounding, not the physics
    beam.use_member_operators = False
    X2 = torch.from_numpy(x0).cuda()
    rk4_steps(beam, X2, 0.0, h, steps, u=TipImpulse(torch.from_numpy(amp).cuda(), duration=80.5 * h), controller=ctrl)
    b2 = X2.cpu().numpy()
    assert not np.array_equal(b2, got) and max(block_err(got[i], b2[i], n) for i in range(B)) < 1e-10


@pytest.mark.parametrize("N,B", [(15, 3), (16, 2), (32, 2)])
def test_lqr_gains_of_long_beams(N, B):
    """Designs whose 4n x 4n Hamiltonian does not fit shared memory (n > 42: up to the 32-element cantilever of
    config 3, n = 96, a 384 x 384 Hamiltonian): the work matrix lives in the kernel's global workspace.  Same
    parity bar as the short beams."""
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator
    from oracle import beam_oracle as bo

    par, et, bc = design_ensemble(B, N, seed=900 + N)
    beam = make_gpu_beam(par, et, bc)
    n = beam.n_free
    Q, R, Qd, Rd = weights(n)
    Md, Kd = beam.dense_matrices()
    lqr = BatchedLinearQuadraticRegulator(Kd, Md, Qd, Rd)
    G = lqr.compute_gain_matrix().cpu().numpy()
    assert int(lqr.status.abs().sum()) == 0
    Mh, Kh = Md.cpu().numpy(), Kd.cpu().numpy()
    for i in range(B):
        ks = bo.lqr_gain(Kh[i], Mh[i], Q, R)
        kr = bo.lqr_gain_refined(Kh[i], Mh[i], Q, R)
        assert np.abs(G[i] - ks).max() / np.abs(ks).max() < TOL_VS_SCIPY
        assert np.abs(G[i] - kr).max() / np.abs(kr).max() < TOL_VS_REFINED
