"""Parity at BASELINE.json's FULL sizes (config 3: 65,536 members x 32 elements, 1000 RK4 steps;
config 5: 131,072 members): sampled members against the reference's golden trajectories, plus
size-independent properties (exact linearity, agreement of independent kernel families)."""

import numpy as np
import pytest

from helpers import block_err, load, make_gpu_beam

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _cfg3_beam(e):
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    B, N = e.E.shape
    par = np.empty((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    return make_gpu_beam(par, np.zeros(N, dtype=int), np.array([1] + [0] * N))


def test_config3_full_ensemble_against_reference_samples():
    """The whole 65,536-member ensemble is integrated for 1000 steps exactly as bench.py does
    (50 fused steps per launch); the 64 members the reference integrated on the CPU must agree
    <= 1e-9 at steps 250, 500, 750, 1000.  (E differs from the golden run by pandas' CSV rounding,
    <= 1e-13 relative)"""
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps

    g = load("cfg3_samples.npz")
    e = ens.config3()
    beam = _cfg3_beam(e)
    n = beam.n_free
    X = torch.from_numpy(np.concatenate([e.q0, e.v0], axis=1)).cuda()
    idx = torch.from_numpy(g["idx"]).cuda()
    system = beam.make_system(e.n_members)
    worst = 0.0
    for k in range(20):
        rk4_steps(beam, X, k * 50 * e.h, e.h, 50, system=system)
        if (k + 1) % 5 == 0:
            got = X[idx].cpu().numpy()
            ref = g["Y"][:, (k + 1) // 5 - 1]
            worst = max(worst, max(block_err(got[i], ref[i], n) for i in range(len(ref))))
    assert torch.isfinite(X).all()
    assert worst < 1e-9, worst


def test_config3_full_size_linearity_and_kernel_families():
    """Linear force-free beams: scaling the initial state by 2 scales every trajectory by exactly 2
    (power-of-two scaling commutes with every FP64 operation of the kernel), and the three
    independent kernel families (paired / staged / general) agree to <= 1e-11 on all 65,536 members."""
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps

    e = ens.config3()
    beam = _cfg3_beam(e)
    n = beam.n_free
    x0 = torch.from_numpy(np.concatenate([e.q0, e.v0], axis=1)).cuda()
    out = {}
    for mode in ("paired", "staged", "general"):
        beam.force_staged_kernels = mode == "staged"
        beam.force_general_kernels = mode == "general"
        X = x0.clone()
        rk4_steps(beam, X, 0.0, e.h, 100)
        out[mode] = X
    beam.force_staged_kernels = beam.force_general_kernels = False
    X2 = 2.0 * x0
    rk4_steps(beam, X2, 0.0, e.h, 100)
    assert torch.equal(X2, 2.0 * out["paired"])
    for mode in ("staged", "general"):
        for sl in (slice(0, n), slice(n, 2 * n)):
            num = (out[mode][:, sl] - out["paired"][:, sl]).abs().amax(dim=1)
            den = out["paired"][:, sl].abs().amax(dim=1)
            assert float((num / den).max()) < 1e-11, mode


def test_config5_full_shard_against_reference_samples():
    """One GPU's shard of config 5 (131,072 members, shared LQR gain on the tensor cores): the 32
    members of the golden file are planted at the head of the shard and must match <= 1e-9."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.integrate import rk4_steps
    from helpers import params_array

    g = load("cfg5_samples.npz")
    B = 131072
    e = ens.config5(B)
    amp = e.impulse_amp.copy()
    amp[: len(g["amp"])] = g["amp"]
    beam = make_gpu_beam(params_array(g)[None], g["elem_type"], g["bc"], 0.0, True)
    n = beam.n_free
    ctrl = FullStateLinear(torch.from_numpy(g["gain"]).cuda())
    X = torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda")
    imp = TipImpulse(torch.from_numpy(amp).cuda())
    h = float(g["h"])
    for k in range(4):
        rk4_steps(beam, X, k * 500 * h, h, 500, u=imp, controller=ctrl)
        got = X[: len(g["amp"])].cpu().numpy()
        ref = g["Y"][:, k]
        assert max(block_err(got[i], ref[i], n) for i in range(len(ref))) < 1e-9
    assert torch.isfinite(X).all()


def test_config4_full_size_member_independence():
    """Config 4 at full size (4,096 nonlinear 64-element members, drag + gravity + per-member impulse,
    adaptive RK45 with per-member dt): a member's trajectory, step counts and final step size do not depend
    on which other members share its warp, block or launch -- re-running a scattered subset alone
    reproduces the full run bit for bit (members are independent in the reference: one solve_ivp each)."""
    from continuum_robot_b200 import TipImpulse, solve_ensemble
    from continuum_robot_b200 import ensembles as ens

    e = ens.config4()
    m = ens.material()
    B, N = e.n_members, e.n_elements
    par = np.empty((B, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 3], par[:, :, 4] = m["length"], m["I"], m["rho"], m["A"]
    par[:, :, 1] = e.E
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    et, bc = np.ones(N, dtype=int), np.array([1] + [0] * N)
    te = np.linspace(0.0, 1.5e-3, 4)

    def run(sel):
        beam = make_gpu_beam(par[sel], et, bc, 1000.0, True)
        n = beam.n_free
        X0 = torch.zeros(len(sel), 2 * n, dtype=torch.float64, device="cuda")
        amp = torch.from_numpy(e.impulse_amp[sel]).cuda()
        return solve_ensemble(beam, (0.0, 1.5e-3), X0, method="RK45", t_eval=te, rtol=1e-6, atol=1e-9, u=TipImpulse(amp))

    full = run(np.arange(B))
    assert full.success and torch.isfinite(full.y).all()
    sel = np.array([0, 1, 17, 1023, 2048, 2049, 4000, B - 1])
    part = run(sel)
    idx = torch.from_numpy(sel).cuda()
    assert torch.equal(part.y, full.y[idx])
    assert torch.equal(part.nfev, full.nfev[idx]) and torch.equal(part.naccept, full.naccept[idx])
    assert torch.equal(part.h_last, full.h_last[idx])
    assert len(set(full.nfev.tolist())) > 10  # the ensemble really is ragged


def test_config5_full_shard_of_a_design_ensemble():
    """One GPU's shard of config 5 (131,072 members) as a DESIGN ensemble: 2,048 designs synthesised on the device
    (crb_dense_matrices_batched + crb_lqr_gains), each rolled out 64 times with its own gain (staged-gain kernel).
    Size-independent properties: every synthesis succeeds with a small Riccati residual, replicas of a design are
    bitwise identical wherever they sit in the ensemble, and the closed loop damps the disturbance."""
    from continuum_robot_b200 import BatchedLinearQuadraticRegulator, FullStateLinear, TipImpulse
    from continuum_robot_b200 import ensembles as ens
    from continuum_robot_b200.dynamic_beam import BatchedDynamicEulerBernoulliBeam
    from continuum_robot_b200.force_params import ForceParams
    from continuum_robot_b200.integrate import rk4_steps

    # h: the fastest closed loop of the 2,048 designs (stiff, light tail of the distribution) needs h < 4.7e-6
    D, rep, N, h = 2048, 64, 6, 2e-6
    B = D * rep
    rng = np.random.default_rng(555)
    m = ens.material()
    par = np.zeros((D, N, 7))
    par[:, :, 0], par[:, :, 2], par[:, :, 4] = m["length"], m["I"], m["A"]
    par[:, :, 1] = m["E"] * np.exp(0.3 * rng.standard_normal((D, 1)))
    par[:, :, 3] = m["rho"] * np.exp(0.2 * rng.standard_normal((D, 1)))
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    designs = BatchedDynamicEulerBernoulliBeam({"params": par, "type": ["linear"] * N}, ForceParams(enable_gravity_effects=True))
    n = designs.n_free
    Q = torch.diag(torch.cat([torch.full((n,), 100.0), torch.full((n,), 10.0)])).to("cuda", torch.float64)
    R = torch.eye(n, dtype=torch.float64, device="cuda")
    Md, Kd = designs.dense_matrices()
    lqr = BatchedLinearQuadraticRegulator(Kd, Md, Q, R)
    gain = lqr.compute_gain_matrix()
    assert int(lqr.status.abs().sum()) == 0 and float(lqr.residual.max()) < 1e-5
    beam = BatchedDynamicEulerBernoulliBeam({"params": np.tile(par, (rep, 1, 1)), "type": ["linear"] * N},
                                            ForceParams(enable_gravity_effects=True))
    beam.create_system_func()
    beam.create_input_func()
    amp = torch.from_numpy(np.tile(rng.uniform(1.0, 20.0, D), rep)).cuda()
    ctrl = FullStateLinear(gain.repeat(rep, 1, 1))
    X = torch.zeros(B, 2 * n, dtype=torch.float64, device="cuda")
    rk4_steps(beam, X, 0.0, h, 5200, u=TipImpulse(amp, duration=0.01), controller=ctrl)
    assert bool(torch.isfinite(X).all())
    Xr = X.view(rep, D, 2 * n)
    assert bool((Xr == Xr[0:1]).all())
    peak = float(X[:, :n].abs().max())
    rk4_steps(beam, X, 5200 * h, h, 4000, controller=ctrl)  # impulse over: the regulator pulls the state back
    assert float(X[:, n:].abs().max()) < 10.0 and float(X[:, :n].abs().max()) <= 1.5 * peak
