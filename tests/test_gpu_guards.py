"""Out-of-bounds WRITE check without compute-sanitizer (closed on the GPU pool): every array a kernel writes is a
view into a larger buffer whose surroundings hold a sentinel bit pattern; ragged member counts make the last tile /
block / warp partial in every kernel family.  After the launch the sentinels must be intact, bit for bit.
Covers crb_rk4 (persistent paired kernel, per-member-mass variant, gravity fast kernel, general nonlinear kernel,
shared-operator kernel, dense per-member operators), crb_midpoint, crb_rk45 (pilot + member order, dense output) with
full and lean recording."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import make_gpu_beam  # noqa: E402

SENT = float(np.frombuffer(np.uint64(0x7FF8DEADBEEF1234).tobytes(), dtype=np.float64)[0])  # a NaN with a payload
PAD = 4096  # doubles on each side


class Guarded:
    """A CUDA tensor of the given shape inside sentinel-filled padding."""

    def __init__(self, shape, dtype=torch.float64, fill=0.0):
        n = int(np.prod(shape))
        self.dtype = dtype
        if dtype == torch.float64:
            self.buf = torch.full((n + 2 * PAD,), SENT, dtype=dtype, device="cuda")
        else:
            self.buf = torch.full((n + 2 * PAD,), -1234567, dtype=dtype, device="cuda")
        self.t = self.buf[PAD:PAD + n].view(*shape)
        self.t.fill_(fill)
        self.n = n

    def intact(self):
        lo, hi = self.buf[:PAD], self.buf[PAD + self.n:]
        if self.dtype == torch.float64:
            want = torch.full((PAD,), SENT, dtype=torch.float64, device="cuda").view(torch.int64)
            return bool(torch.equal(lo.view(torch.int64), want)) and bool(torch.equal(hi.view(torch.int64), want))
        return bool((lo == -1234567).all()) and bool((hi == -1234567).all())


def _params(B, N, rng, vary_mass=False):
    from continuum_robot_b200 import ensembles as ens

    m = ens.material()
    par = np.zeros((B, N, 7))
    par[:, :, 0] = m["length"]
    par[:, :, 1] = m["E"] * np.exp(0.2 * rng.standard_normal((B, N)))
    par[:, :, 2], par[:, :, 4] = m["I"], m["A"]
    par[:, :, 3] = m["rho"] * (np.exp(0.1 * rng.standard_normal((B, 1))) if vary_mass else 1.0)
    par[:, :, 5], par[:, :, 6] = m["wetted_area"], m["drag_coef"]
    return par


def _state(B, n, rng):
    X = Guarded((B, 2 * n))
    X.t.copy_(torch.from_numpy(np.concatenate([1e-3 * rng.standard_normal((B, n)), 1e-1 * rng.standard_normal((B, n))], axis=1)))
    return X


@pytest.mark.parametrize("B", [1, 5, 37, 1001])
@pytest.mark.parametrize("kind", ["lin32", "lin32_pm", "lin10_grav", "nl20_drag", "pinned33"])
def test_fixed_step_kernels_write_only_their_arrays(kind, B):
    from continuum_robot_b200 import TipImpulse
    from continuum_robot_b200.integrate import midpoint_steps, rk4_steps
    from continuum_robot_b200.outputs import output_selection

    rng = np.random.default_rng(B)
    N = {"lin32": 32, "lin32_pm": 32, "lin10_grav": 10, "nl20_drag": 20, "pinned33": 32}[kind]
    par = _params(B, N, rng, vary_mass=kind == "lin32_pm")
    et = np.full(N, 1 if kind == "nl20_drag" else 0)
    bc = np.array([2 if kind == "pinned33" else 1] + [0] * N)
    beam = make_gpu_beam(par, et, bc, 1000.0 if kind == "nl20_drag" else 0.0, kind == "lin10_grav")
    n, h, steps = beam.n_free, 2e-6, 6
    amp = Guarded((B,))
    amp.t.copy_(torch.from_numpy(rng.uniform(0.1, 1.0, B)))
    u = TipImpulse(amp.t, duration=5e-6) if kind in ("lin10_grav", "nl20_drag") else None
    for sel in (None, "tip"):
        X = _state(B, n, rng)
        width = 2 * n if sel is None else len(output_selection(beam, sel))
        Y = Guarded((3, B, width))
        rk4_steps(beam, X.t, 0.0, h, steps, u=u, Y_out=Y.t, save_every=2,
                  out_sel=None if sel is None else output_selection(beam, sel))
        torch.cuda.synchronize()
        assert X.intact() and Y.intact() and amp.intact(), (kind, B, sel)
        assert bool(torch.isfinite(X.t).all()) and bool(torch.isfinite(Y.t).all())
    if kind in ("lin32", "lin32_pm", "pinned33"):
        X = _state(B, n, rng)
        Y = Guarded((2, B, 2 * n))
        midpoint_steps(beam, X.t, 0.0, 2e-5, 4, Y_out=Y.t, save_every=2)
        torch.cuda.synchronize()
        assert X.intact() and Y.intact() and bool(torch.isfinite(X.t).all())


@pytest.mark.parametrize("B", [1, 9, 70])
def test_closed_loop_kernels_write_only_their_arrays(B):
    """Shared-operator kernel (one gain, with and without axial / bending coupling) and dense per-member operators."""
    from continuum_robot_b200 import FullStateLinear, TipImpulse
    from continuum_robot_b200.integrate import rk4_steps

    rng = np.random.default_rng(50 + B)
    N = 6
    par1 = _params(1, N, rng)
    et, bc = np.zeros(N, dtype=int), np.array([1] + [0] * N)
    beam = make_gpu_beam(par1, et, bc, 0.0, True)
    n = beam.n_free
    amp = Guarded((B,))
    amp.t.copy_(torch.from_numpy(rng.uniform(1.0, 5.0, B)))
    for decoupled in (False, True):
        gain = np.concatenate([50.0 * rng.standard_normal((n, n)), 0.05 * rng.standard_normal((n, n))], axis=1)
        if decoupled:
            ax = np.arange(n) % 3 == 0
            cross = ax[:, None] != ax[None, :]
            gain[np.concatenate([cross, cross], axis=1)] = 0.0
        X = _state(B, n, rng)
        Y = Guarded((2, B, 2 * n))
        rk4_steps(beam, X.t, 0.0, 2e-6, 8, u=TipImpulse(amp.t, duration=1e-5),
                  controller=FullStateLinear(torch.from_numpy(gain).cuda()), Y_out=Y.t, save_every=4)
        torch.cuda.synchronize()
        assert X.intact() and Y.intact() and amp.intact(), (B, decoupled)
        assert bool(torch.isfinite(X.t).all())
    # one design and gain per member: dense per-member operators
    parB = _params(B, N, rng)
    beamB = make_gpu_beam(parB, et, bc, 0.0, True)
    gains = torch.from_numpy(np.concatenate([50.0 * rng.standard_normal((B, n, n)), 0.05 * rng.standard_normal((B, n, n))], axis=2)).cuda()
    X = _state(B, n, rng)
    Y = Guarded((2, B, 2 * n))
    rk4_steps(beamB, X.t, 0.0, 2e-6, 8, u=TipImpulse(amp.t, duration=1e-5), controller=FullStateLinear(gains), Y_out=Y.t, save_every=4)
    torch.cuda.synchronize()
    assert X.intact() and Y.intact() and amp.intact() and bool(torch.isfinite(X.t).all())


@pytest.mark.parametrize("B,N,nl", [(3, 64, True), (41, 64, True), (130, 20, True), (77, 4, False)])
def test_adaptive_kernel_writes_only_its_arrays(B, N, nl):
    """crb_rk45 through the C ABI with guarded X, t, h, status, counters, Y_eval, in two launches (attempt budget, then a
    resume in reversed member order)."""
    import ctypes as C

    from continuum_robot_b200 import _lib

    rng = np.random.default_rng(7 * B)
    par = _params(B, N, rng)
    et, bc = np.full(N, 1 if nl else 0), np.array([1] + [0] * N)
    beam = make_gpu_beam(par, et, bc, 1000.0 if nl else 0.0, True)
    if N <= 64:
        beam = beam.with_slots(2)
    n = beam.n_free
    X = _state(B, n, rng)
    X.t.mul_(1e-2)
    t, hh = Guarded((B,)), Guarded((B,))
    status, counters = Guarded((B,), torch.int32, 0), Guarded((B, 3), torch.int64, 0)
    te = torch.linspace(0.0, 2e-4, 5, dtype=torch.float64, device="cuda")
    Y = Guarded((5, B, 2 * n))
    order = Guarded((B,), torch.int32, 0)
    order.t.copy_(torch.arange(B - 1, -1, -1, dtype=torch.int32, device="cuda"))
    sysm, keep = beam.make_system(B, drag=beam._active_forces()[0], gravity=beam._active_forces()[1])
    lib = _lib.load()
    for budget, use_order in ((3, False), (100000, True)):
        sysm.member_order = order.t.data_ptr() if use_order else None
        with torch.cuda.device(beam.device):
            rc = lib.crb_rk45(C.byref(beam._plan), C.byref(sysm), X.t.data_ptr(), t.t.data_ptr(), hh.t.data_ptr(), 2e-4, 1e-6, 1e-9,
                              te.data_ptr(), 5, Y.t.data_ptr(), status.t.data_ptr(), counters.t.data_ptr(), budget, beam._stream())
        _lib.check(rc)
        torch.cuda.synchronize()
        for g in (X, t, hh, status, counters, Y, order):
            assert g.intact(), (B, N, budget)
    assert int(status.t.abs().sum()) == 0 and bool((t.t == 2e-4).all()) and bool(torch.isfinite(Y.t).all())


def test_guard_detects_a_stray_write():
    g = Guarded((7, 3))
    assert g.intact()
    g.buf[PAD + g.n] = 0.0  # one double past the end
    assert not g.intact()
    g2 = Guarded((5,), torch.int32, 0)
    g2.buf[PAD - 1] = 0
    assert not g2.intact()


@pytest.mark.parametrize("B,N", [(1, 6), (13, 6), (5, 14), (3, 20)])
def test_lqr_and_rhs_entry_points_write_only_their_arrays(B, N):
    """crb_dense_matrices_batched, crb_lqr_gains (shared-memory and workspace paths), crb_member_operators, crb_rhs and
    crb_forces through the C ABI on guarded outputs (the workspace is guarded too)."""
    import ctypes as C

    from continuum_robot_b200 import _lib

    rng = np.random.default_rng(B + N)
    par = _params(B, N, rng, vary_mass=True)
    et, bc = np.zeros(N, dtype=int), np.array([1] + [0] * N)
    beam = make_gpu_beam(par, et, bc, 1000.0, True)
    n = beam.n_free
    lib = _lib.load()
    dpar = torch.from_numpy(par).cuda()
    M, K = Guarded((B, n, n)), Guarded((B, n, n))
    with torch.cuda.device(beam.device):
        _lib.check(lib.crb_dense_matrices_batched(C.byref(beam._plan), dpar.data_ptr(), B, bytes(int(t) for t in et), bytes(int(b) for b in bc),
                                                  M.t.data_ptr(), K.t.data_ptr(), beam._stream()))
        torch.cuda.synchronize()
        assert M.intact() and K.intact()
        Q = torch.diag(torch.cat([torch.full((n,), 100.0), torch.full((n,), 10.0)])).to("cuda", torch.float64)
        R = torch.eye(n, dtype=torch.float64, device="cuda")
        need = C.c_size_t(0)
        _lib.check(lib.crb_lqr_workspace_bytes(n, B, C.byref(need)))
        ws = Guarded(((need.value + 7) // 8,))
        gain, S, resid, status = Guarded((B, n, 2 * n)), Guarded((B, 2 * n, 2 * n)), Guarded((B,)), Guarded((B,), torch.int32, 0)
        _lib.check(lib.crb_lqr_gains(n, B, M.t.data_ptr(), 0, K.t.data_ptr(), 0, Q.data_ptr(), R.data_ptr(), 1, gain.t.data_ptr(),
                                     S.t.data_ptr(), resid.t.data_ptr(), status.t.data_ptr(), ws.t.data_ptr(), need.value, beam._stream()))
        torch.cuda.synchronize()
        for g in (M, K, ws, gain, S, resid, status):
            assert g.intact(), (B, N)
        assert int(status.t.abs().sum()) == 0 and bool(torch.isfinite(gain.t).all())
        if n <= 32:
            op, st = Guarded((B, n, 3 * n + 1)), Guarded((B,), torch.int32, 0)
            _lib.check(lib.crb_member_operators(n, B, M.t.data_ptr(), 0, K.t.data_ptr(), 0, gain.t.data_ptr(), None, op.t.data_ptr(),
                                                st.t.data_ptr(), beam._stream()))
            torch.cuda.synchronize()
            assert op.intact() and st.intact() and gain.intact()
        drag, grav, _ = beam._active_forces()
        sysm, keep = beam.make_system(B, drag=drag, gravity=grav)
        X = _state(B, n, rng)
        dX, F = Guarded((B, 2 * n)), Guarded((B, n))
        _lib.check(lib.crb_rhs(C.byref(beam._plan), C.byref(sysm), X.t.data_ptr(), 0.0, dX.t.data_ptr(), beam._stream()))
        _lib.check(lib.crb_forces(C.byref(beam._plan), C.byref(sysm), X.t.data_ptr(), F.t.data_ptr(), beam._stream()))
        torch.cuda.synchronize()
        assert X.intact() and dX.intact() and F.intact() and bool(torch.isfinite(dX.t).all())
