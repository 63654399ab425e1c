"""Ensemble time integration: the integrator entry point of the drop-in boundary.

Replaces ``scipy.integrate.solve_ivp(fun, t_span, y0, method, t_eval, rtol, atol)`` as the
reference's callers use it (examples/example_utilities.py:153-168, examples/lqr_control.py:117-128,
examples/pyodide_example/pyodide_example.py:69-75) with a batched call whose result mirrors
``OdeResult`` (``.t``, ``.y``, ``.nfev``, ``.status``, ``.success``, ``.message``), per member.

  * ``method="RK4"``  classical fixed-step RK4, ``nsteps`` fused per launch (crb_rk4);
  * ``method="RK45"`` Dormand-Prince 5(4) with SciPy's controller, per-member dt (crb_rk45);
  * ``method="MIDPOINT"`` implicit midpoint rule for all-linear beams (crb_midpoint): not a reference code
    path, the stiff-capable alternative to the LSODA runs of the reference's examples.
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .dynamic_beam import (BatchedDynamicEulerBernoulliBeam, PiecewiseLinearInput, SinusoidInput, TipImpulse,
                           split_input)

MESSAGES = {
    0: "The solver successfully reached the end of the integration interval.",
    -1: "Required step size is less than spacing between numbers.",
    1: "Attempt budget of the launch exhausted before reaching t_bound.",
}


@dataclass
class EnsembleResult:
    t: "np.ndarray"          # [T] output times
    y: "object"              # torch [B, 2n, T]
    nfev: "object"           # torch int64 [B]
    status: "object"         # torch int32 [B]
    success: bool
    message: str
    naccept: "object" = None
    nreject: "object" = None
    njev: int = 0
    nlu: int = 0
    rows: "object" = None     # state indices of the rows of y for a lean recording (None: the whole state)
    x_final: "object" = None  # torch [B, 2n] full state at the end of the interval


def with_selection(beam, system, out_sel):
    """``(crb_system_t, keepalive)`` of ``beam.make_system`` with the lean-recording table of ``out_sel`` attached: for
    callers that step repeatedly with ``rk4_steps(..., system=...)`` and want the table uploaded once."""
    s2, keep, _ = _with_selection(beam, system[0], system[1], out_sel)
    return s2, keep


def _check_system_owner(beam, sysm):
    """A prebuilt crb_system_t carries device arrays in ONE beam's lane layout (e.g. ``beam.with_slots(2)`` is a
    different layout than ``beam``): using it with another beam would read mfac / kcoef / red_index wrongly."""
    key = getattr(sysm, "_crb_plan_key", None)
    if key is not None and key != beam._plan_key():
        raise ValueError("this prebuilt system was made by another beam object (different lane layout / factor arrays); "
                         "build it with the beam it is used with")


def _with_selection(beam, sysm, keep, out_sel):
    """Copy of the crb_system_t with the lean-recording table of ``out_sel`` (state indices) attached."""
    import torch

    if out_sel is None:
        return sysm, keep, 2 * beam.n_free
    idx = list(out_sel)
    inv = np.full(2 * beam.n_free, -1, dtype=np.int32)
    inv[np.asarray(idx, dtype=np.int64)] = np.arange(len(idx), dtype=np.int32)
    d_inv = torch.from_numpy(inv).to(beam.device)
    s2 = type(sysm).from_buffer_copy(sysm)
    if hasattr(sysm, "_crb_plan_key"):
        s2._crb_plan_key = sysm._crb_plan_key
    s2.out_sel_inv = d_inv.data_ptr()
    s2.out_n_sel = len(idx)
    return s2, list(keep) + [d_inv], len(idx)


def _feedback(controller):
    """(gain, ref) from a FullStateLinear-like controller, or (None, None)."""
    if controller is None:
        return None, None
    if not controller.is_enabled():
        return None, None
    return controller.gain_matrix, getattr(controller, "reference", None)


def _feedback_layout(beam, controller):
    """State feedback runs on the FP64 tensor cores when a member spans 4 lanes (8 members = the 8
    rows of mma.m8n8k4): short beams are re-assembled with 2 slots per lane to get that layout."""
    if controller is not None and controller.is_enabled() and int(beam._plan.g) != 4 and beam._plan.p_act <= 8:
        if getattr(getattr(controller, "gain_matrix", None), "ndim", 2) == 3:
            return beam  # one gain per member: no shared operand for the tensor cores, keep the default layout
        return beam.with_slots(2)
    return beam


def rk4_steps(beam: BatchedDynamicEulerBernoulliBeam, X, t0: float, h: float, nsteps: int, *, u=None,
              controller=None, Y_out=None, save_every: int = 0, system=None, out_sel=None):
    """Advance X[B,2n] in place by ``nsteps`` classical RK4 steps in ONE kernel launch.

    ``out_sel``: state indices for a lean recording (``outputs.output_selection``): ``Y_out`` is then
    ``[T, B, len(out_sel)]`` and only those entries are written by the kernel.

    ``system``: a prebuilt ``(crb_system_t, keepalive)`` from ``beam.make_system`` to skip the
    per-call struct fill (used by bench.py).
    """
    import torch

    if system is None:
        drag, grav, user = beam._active_forces()
        if user or beam._forces_func is not None:
            raise TypeError("fused RK4 supports built-in forces only; use solve_ensemble(...) for torch force callables")
        uc, impulse, tv, other = split_input(u)
        if other is not None:
            raise TypeError("fused RK4 needs tensor / TipImpulse / SinusoidInput / PiecewiseLinearInput inputs; "
                            "use solve_ensemble for free-form callables")
        gain, ref = _feedback(controller)
        beam = _feedback_layout(beam, controller)
        system = beam.make_system(X.shape[0], drag=drag, gravity=grav, u_const=uc, impulse=impulse, gain=gain, ref=ref,
                                  time_inputs=tv)
    elif u is not None or controller is not None:
        raise ValueError("rk4_steps: `u` / `controller` are part of a prebuilt `system`; pass one or the other")
    sysm, _keep = system
    _check_system_owner(beam, sysm)
    if out_sel is not None:
        sysm, _keep, width = _with_selection(beam, sysm, _keep, out_sel)
        if Y_out is not None and (Y_out.shape[-1] != width or not Y_out.is_contiguous()):
            raise ValueError(f"Y_out must be contiguous [T, B, {width}] for this output selection")
    with torch.cuda.device(beam.device):
        rc = _lib.load().crb_rk4(
            C.byref(beam._plan), C.byref(sysm), X.data_ptr(), float(t0), float(h), int(nsteps),
            Y_out.data_ptr() if Y_out is not None else None, int(save_every), beam._stream(),
        )
    _lib.check(rc)
    return X


def midpoint_steps(beam: BatchedDynamicEulerBernoulliBeam, X, t0: float, h: float, nsteps: int, *, u=None,
                   Y_out=None, save_every: int = 0, out_sel=None, system=None):
    """Advance X[B,2n] in place by ``nsteps`` implicit-midpoint steps (Newmark average acceleration) in ONE
    kernel launch (crb_midpoint).  All-linear beams without drag / gravity / feedback; ``u``: constant
    tensor [B,n] or TipImpulse, evaluated at the step midpoints.  Unconditionally stable: ``h`` is chosen
    for accuracy (the rule is second order), not for the highest element frequency as with RK4 -- the
    reason the reference's examples integrate with LSODA (examples/example_utilities.py:153-159).
    ``system``: a prebuilt ``(crb_system_t, keepalive)`` (``beam.make_system``, optionally with a lean-recording table
    from ``with_selection``) to skip the per-call struct fill."""
    import torch

    if system is None:
        drag, grav, user = beam._active_forces()
        if user or beam._forces_func is not None or drag is not None or grav is not None:
            raise TypeError("the implicit midpoint rule supports force-free all-linear beams (inputs: tensor or TipImpulse)")
        uc, impulse, tv, other = split_input(u)
        if tv or other is not None:
            raise TypeError("implicit midpoint needs a constant tensor or TipImpulse input")
        sysm, _keep = beam.make_system(X.shape[0], u_const=uc, impulse=impulse)
    else:
        if u is not None:
            raise ValueError("midpoint_steps: `u` is part of a prebuilt `system`; pass one or the other")
        sysm, _keep = system
        _check_system_owner(beam, sysm)
    if out_sel is not None:
        sysm, _keep, width = _with_selection(beam, sysm, _keep, out_sel)
        if Y_out is not None and (Y_out.shape[-1] != width or not Y_out.is_contiguous()):
            raise ValueError(f"Y_out must be contiguous [T, B, {width}] for this output selection")
    afac, shared = beam.shifted_factors(0.25 * float(h) * float(h))
    with torch.cuda.device(beam.device):
        rc = _lib.load().crb_midpoint(
            C.byref(beam._plan), C.byref(sysm), afac.data_ptr(), int(shared), X.data_ptr(), float(t0), float(h), int(nsteps),
            Y_out.data_ptr() if Y_out is not None else None, int(save_every), beam._stream(),
        )
    _lib.check(rc)
    return X


class HostPipeline:
    """Fixed-step RK4 on a HOST-resident ensemble: ``x_host[B, 2n]`` (pinned) is advanced in place.

    Thin wrapper of the native pipeline (``crb_rk4_host``, csrc/crb_host.cu): the ensemble is cut into
    member chunks of whole kernel waves; chunk c+1 is copied host->device and chunk c-1 device->host
    while chunk c integrates (three CUDA streams, PCIe is full duplex), and consecutive ``run`` calls
    overlap chunk by chunk, so a sequence of calls costs about max(copy-in, compute, copy-out)
    instead of their sum.  Members are independent, so chunking changes nothing numerically.

    ``chunk_members=0`` picks two full waves of the kernel the library dispatches (measured best).
    """

    def __init__(self, beam: BatchedDynamicEulerBernoulliBeam, n_members: int, chunk_members: int = 0, *, u=None,
                 controller=None):
        import torch

        self.B = int(n_members)
        drag, grav, user = beam._active_forces()
        if user or beam._forces_func is not None:
            raise TypeError("HostPipeline supports built-in forces only")
        impulse = u if isinstance(u, TipImpulse) else None
        uc = None if impulse is not None else u
        gain, ref = _feedback(controller)
        beam = self.beam = _feedback_layout(beam, controller)
        self.n2 = 2 * beam.n_free
        self.X = torch.empty((self.B, self.n2), dtype=torch.float64, device=beam.device)  # device workspace
        self.system = beam.make_system(self.B, drag=drag, gravity=grav, u_const=uc, impulse=impulse, gain=gain, ref=ref)
        lib = _lib.load()
        with torch.cuda.device(beam.device):
            wave = C.c_int32(0)
            _lib.check(lib.crb_rk4_wave_members(C.byref(beam._plan), C.byref(self.system[0]), C.byref(wave)))
            self.chunk_members = int(chunk_members) if chunk_members and chunk_members > 0 else 2 * int(wave.value)
            self._handle = C.c_void_p()
            _lib.check(lib.crb_pipeline_create(C.byref(self._handle)))

    def run(self, x_host, t0: float, h: float, nsteps: int):
        """Enqueue H2D -> nsteps fused RK4 steps -> D2H for every chunk and return immediately.  The work
        is ordered after the current stream; call ``wait()`` (stream-ordered) or ``synchronize()`` (host)
        before reading ``x_host``."""
        import torch

        if x_host.device.type != "cpu" or not x_host.is_pinned() or tuple(x_host.shape) != (self.B, self.n2) \
                or x_host.dtype != torch.float64 or not x_host.is_contiguous():
            raise ValueError("x_host must be a pinned, contiguous float64 CPU tensor of shape [B, 2n]")
        beam = self.beam
        with torch.cuda.device(beam.device):
            rc = _lib.load().crb_rk4_host(self._handle, C.byref(beam._plan), C.byref(self.system[0]), x_host.data_ptr(),
                                          self.X.data_ptr(), self.chunk_members, float(t0), float(h), int(nsteps),
                                          beam._stream())
        _lib.check(rc)
        return x_host

    def wait(self):
        """Make the current CUDA stream wait for every outstanding chunk."""
        import torch

        with torch.cuda.device(self.beam.device):
            _lib.check(_lib.load().crb_pipeline_wait(self._handle, self.beam._stream()))

    def synchronize(self):
        import torch

        with torch.cuda.device(self.beam.device):
            _lib.check(_lib.load().crb_pipeline_synchronize(self._handle))

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            _lib.load().crb_pipeline_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _input_force(B, n, ts, uc, impulse, tv, other):
    """u(ts)[B, n] assembled on the device from the parts of split_input; ``ts``: float or per-member tensor [B]."""
    import torch

    force = None

    def add(a):
        nonlocal force
        if not isinstance(a, torch.Tensor):
            raise ValueError("State and input must be torch tensors")
        if a.shape[-1] != n:
            raise ValueError(f"Input vector length {a.shape[-1]} must match position DOFs {n}. Expected {n}, got {a.shape[-1]}")
        a = a if a.ndim == 2 else a.unsqueeze(0)
        force = a if force is None else force + a

    if uc is not None:
        add(uc)
    for part in tv:
        add(part(ts))
    if other is not None:
        # the reference calls u(t) with a scalar; members of an adaptive ensemble sit at different times, so the
        # callable receives a column [B, 1] there (torch broadcasting keeps `torch.sin(t) * ones(n)` working)
        add(other(ts.unsqueeze(-1) if isinstance(ts, torch.Tensor) else ts))
    if impulse is not None:
        amp = impulse.amplitude
        if not isinstance(amp, torch.Tensor):
            amp = torch.as_tensor(np.asarray(amp, dtype=np.float64))
        dev = force.device if force is not None else (ts.device if isinstance(ts, torch.Tensor) else amp.device)
        amp = amp.to(device=dev, dtype=torch.float64).reshape(-1).expand(B)
        gate = (ts < impulse.duration) if isinstance(ts, torch.Tensor) else float(ts < impulse.duration)
        z = torch.zeros((B, n), dtype=torch.float64, device=dev)
        z[:, impulse.dof] = amp * gate
        add(z)
    return force


def _rk4_unfused(beam, X, t0, h, nsteps, u, controller, Y_out, save_every):
    """RK4 with user torch force / input callables: one crb_rhs launch per stage."""
    gain, ref = _feedback(controller)
    uc, impulse, tv, other = split_input(u)
    B, n = X.shape[0], beam.n_free

    def f(t, x):
        force = _input_force(B, n, t, uc, None, [], other)
        return beam._rhs(t, x, u=force, impulse=impulse, gain=gain, ref=ref, time_inputs=tv)

    x = X
    for k in range(nsteps):
        t = t0 + k * h
        k1 = f(t, x)
        k2 = f(t + 0.5 * h, x + (0.5 * h) * k1)
        k3 = f(t + 0.5 * h, x + (0.5 * h) * k2)
        k4 = f(t + h, x + h * k3)
        x = x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        if Y_out is not None and (k + 1) % save_every == 0:
            Y_out[(k + 1) // save_every - 1].copy_(x)
    X.copy_(x)
    return X


def _rk45_unfused(beam, X, t0, tf, rtol, atol, u, controller, te, Y_eval, max_attempts, first_step, check_every=8):
    """Adaptive RK45 when the right-hand side contains user code (plug-in forces, free-form ``u(t)``).

    Per attempt: 6 x (user torch code + crb_rhs) for the stage derivatives, 6 x crb_rk45_stage, 1 x crb_rk45_control;
    the host looks at the device only every ``check_every`` attempts (to see whether every member is done)."""
    import torch

    lib = _lib.load()
    dev = beam.device
    B, n2 = X.shape
    n = n2 // 2
    gain, ref = _feedback(controller)
    uc, impulse, tv, other = split_input(u)

    def f(ts, x, out=None):
        # built-in forces are time-independent (dynamic_beam_model.py:265 passes t = 0.0), feedback depends on the
        # state only: every time dependence sits in the input, evaluated here at the members' own stage times
        force = _input_force(B, n, ts, uc, impulse, tv, other)
        return beam._rhs(0.0, x, u=force, gain=gain, ref=ref, out=out)

    t = torch.full((B,), t0, dtype=torch.float64, device=dev)
    K = torch.empty((7, B, n2), dtype=torch.float64, device=dev)
    f(t, X, out=K[0])
    counters = torch.zeros((B, 3), dtype=torch.int64, device=dev)
    counters[:, 0] = 1
    span = abs(tf - t0)
    if first_step:
        h_abs = torch.full((B,), float(first_step), dtype=torch.float64, device=dev)
    else:  # select_initial_step (scipy/integrate/_ivp/common.py:68-135), order = 4
        rms = lambda a: torch.sqrt((a * a).mean(dim=1))  # noqa: E731
        scale = atol + X.abs() * rtol
        d0, d1 = rms(X / scale), rms(K[0] / scale)
        h0 = torch.where((d0 < 1e-5) | (d1 < 1e-5), torch.full_like(d0, 1e-6), 0.01 * d0 / d1).clamp(max=span)
        f1 = f(t + h0, X + h0.unsqueeze(-1) * K[0])
        d2 = rms((f1 - K[0]) / scale) / h0
        h1 = torch.where((d1 <= 1e-15) & (d2 <= 1e-15), torch.clamp(h0 * 1e-3, min=1e-6),
                         (0.01 / torch.maximum(d1, d2)) ** 0.2)
        h_abs = torch.minimum(torch.minimum(100.0 * h0, h1), torch.full_like(h0, span))
        counters[:, 0] += 1
    h_abs = h_abs.contiguous()
    h_step = torch.zeros(B, dtype=torch.float64, device=dev)
    t_next = t.clone()
    ts = t.clone()
    Ys = torch.empty_like(X)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    flags = torch.full((B,), 5, dtype=torch.int32, device=dev)
    ie = torch.zeros(B, dtype=torch.int32, device=dev)
    n_eval = 0 if te is None else len(te)
    d_te = None
    if n_eval:
        d_te = torch.from_numpy(te).to(dev)
        first = int(np.searchsorted(te, t0, side="right"))  # SciPy emits t_eval == t0 as y0 itself
        if first:
            Y_eval[:first] = X
        ie.fill_(first)
    stream = beam._stream()

    def control(begin_only):
        _lib.check(lib.crb_rk45_control(
            n2, B, X.data_ptr(), K.data_ptr(), Ys.data_ptr(), t.data_ptr(), t_next.data_ptr(), h_abs.data_ptr(),
            h_step.data_ptr(), float(tf), float(rtol), float(atol), d_te.data_ptr() if d_te is not None else None, n_eval,
            ie.data_ptr(), Y_eval.data_ptr() if Y_eval is not None else None, status.data_ptr(), counters.data_ptr(),
            flags.data_ptr(), int(begin_only), stream))

    with torch.cuda.device(dev):
        control(1)
        attempts = 0
        while attempts < max_attempts:
            for _ in range(min(check_every, max_attempts - attempts)):
                for s in range(1, 7):
                    _lib.check(lib.crb_rk45_stage(n2, B, s, X.data_ptr(), K.data_ptr(), t.data_ptr(), h_step.data_ptr(),
                                                  Ys.data_ptr(), ts.data_ptr(), stream))
                    f(ts, Ys, out=K[s])
                control(0)
                attempts += 1
            if not bool((flags & 1).any().item()):
                break
        else:
            status[(flags & 1).bool()] = 1  # attempt budget exhausted
    return t, h_abs, status, counters


def _pilot_attempts(beam, B: int, requested, max_attempts: int, sm_count: Optional[int] = None) -> int:
    """Attempts of the pilot launch of an adaptive solve (0: one launch in natural order).  ``requested`` None: a pilot
    of 8 attempts when the ensemble is more than one and at most 64 waves of the device's resident warps (one wave
    starts everything at once, and the tail of a very long launch is short relative to its length);
    CRB_RK45_PILOT overrides."""
    import torch

    if requested is None and os.environ.get("CRB_RK45_PILOT") is not None:
        requested = int(os.environ["CRB_RK45_PILOT"])
    if requested is not None:
        return max(0, min(int(requested), int(max_attempts) - 1))
    if sm_count is None:
        sm_count = torch.cuda.get_device_properties(beam.device).multi_processor_count
    slots = 8 * sm_count  # resident warps of the adaptive kernel
    warps = -(-B * int(beam._plan.g) // 32)
    return 8 if slots < warps <= 64 * slots and max_attempts > 8 else 0


def solve_ensemble(beam: BatchedDynamicEulerBernoulliBeam, t_span: Sequence[float], X0, *, method: str = "RK45",
                   h: Optional[float] = None, t_eval=None, rtol: float = 1e-3, atol: float = 1e-6, u=None,
                   controller=None, save_every: Optional[int] = None, max_attempts: int = 10_000_000,
                   first_step: Optional[float] = None, outputs="state", pilot_attempts: Optional[int] = None) -> EnsembleResult:
    """Integrate every member of the ensemble over ``t_span``.

    RK4: ``h`` is required; ``nsteps = round((tf - t0) / h)``; outputs are stored every
    ``save_every`` steps (or at ``t_eval`` if its points fall on the step grid).
    RK45: SciPy semantics -- with ``t_eval`` the dense output is sampled, without it only the
    final state is returned (the per-member step sequences differ, so there is no common grid).

    ``outputs``: ``"state"`` (``.y`` is ``[B, 2n, T]`` like ``OdeResult.y``) or a lean recording -- ``"tip"``,
    ``"shape"``, ``"shape_velocity"`` or a list of state indices (``outputs.output_selection``): the kernels
    then store only those rows (``.y`` is ``[B, len(selection), T]``, ``.rows`` names them) instead of
    materialising the whole trajectory; ``.x_final`` always holds the full final state.

    ``pilot_attempts`` (RK45, fused path): attempts of the pilot launch that orders the members longest-first for
    the main launch (``_pilot_attempts``; None = automatic, 0 = one launch in natural order).  Results do not
    depend on it: every member takes the same step sequence either way.
    """
    import torch

    if beam.system_func is None or beam.input_func is None:
        raise RuntimeError("System and input functions must be created first")
    X, squeeze = beam._as_state(X0)
    X = X.clone()
    X_init = X.clone()  # frame 0 comes from the normalised device copy, whatever device / dtype the caller's X0 had
    B, n2 = X.shape
    t0, tf = float(t_span[0]), float(t_span[1])
    if not tf > t0:
        raise ValueError("t_span must be increasing")
    method = method.upper()
    sel = None
    if not (isinstance(outputs, str) and outputs == "state"):
        from .outputs import output_selection

        sel = output_selection(beam, outputs)
    width = n2 if sel is None else len(sel)
    d_sel = torch.as_tensor(sel, dtype=torch.int64, device=beam.device) if sel is not None else None
    drag, grav, user = beam._active_forces()
    uc, impulse, tv, other = split_input(u)
    needs_unfused = bool(user) or beam._forces_func is not None or other is not None

    if method in ("RK4", "MIDPOINT"):
        if h is None or not h > 0:
            raise ValueError(f"{method} needs a positive step h")
        nsteps = int(round((tf - t0) / h))
        if nsteps < 1 or abs(t0 + nsteps * h - tf) > 1e-9 * max(abs(tf - t0), abs(h)):
            raise ValueError(f"{method}: t_span length {tf - t0} is not a whole number (>= 1) of steps h = {h}")
        if t_eval is not None:
            te = np.asarray(t_eval, dtype=np.float64)
            k = np.rint((te - t0) / h).astype(np.int64)
            if np.any(np.abs(t0 + k * h - te) > 1e-9 * max(1.0, abs(tf))) or np.any(k < 0) or np.any(k > nsteps):
                raise ValueError("for RK4, t_eval must lie on the step grid t0 + k*h")
            ks = k[k > 0]
            se = int(np.gcd.reduce(ks)) if len(ks) else nsteps
            save_every = se
        se = int(save_every) if save_every else nsteps
        if se < 1:
            raise ValueError("save_every must be >= 1")
        nframes = nsteps // se
        fused = method == "MIDPOINT" or not needs_unfused
        Y = torch.empty((nframes, B, width if fused else n2), dtype=torch.float64, device=beam.device) if nframes else None
        if method == "MIDPOINT":
            if needs_unfused or controller is not None:
                raise TypeError("MIDPOINT supports force-free all-linear beams with tensor / TipImpulse inputs only")
            midpoint_steps(beam, X, t0, h, nsteps, u=u, Y_out=Y, save_every=se, out_sel=sel)
        elif needs_unfused:
            _rk4_unfused(beam, X, t0, h, nsteps, u, controller, Y, se)
            if sel is not None and Y is not None:
                Y = Y[:, :, d_sel]
        else:
            rk4_steps(beam, X, t0, h, nsteps, u=u, controller=controller, Y_out=Y, save_every=se, out_sel=sel)
        tt = t0 + se * h * np.arange(1, nframes + 1)
        first = (X_init if sel is None else X_init[:, d_sel]).reshape(1, B, width)
        frames = torch.cat([first, Y], dim=0) if Y is not None else first
        tt = np.concatenate([[t0], tt])
        if t_eval is not None:
            pick = np.rint((np.asarray(t_eval) - t0) / (se * h)).astype(np.int64)
            frames = frames[torch.as_tensor(pick, device=frames.device)]
            tt = tt[pick]
        y = frames.permute(1, 2, 0).contiguous()
        nfev = torch.full((B,), (4 if method == "RK4" else 1) * nsteps, dtype=torch.int64, device=beam.device)
        status = torch.zeros(B, dtype=torch.int32, device=beam.device)
        res = EnsembleResult(tt, y[0] if squeeze else y, nfev, status, True, MESSAGES[0])
        res.x_final = X
        res.rows = sel
        return res

    if method != "RK45":
        raise ValueError(f"method must be 'RK4', 'RK45' or 'MIDPOINT', got {method!r} (LSODA is out of scope)")
    gain, ref = _feedback(controller)
    if not needs_unfused:
        rk45_slots = int(os.environ.get("CRB_RK45_SLOTS", "2"))
        if beam.n_elements <= 64 and rk45_slots > 0:
            beam = beam.with_slots(rk45_slots)  # the adaptive kernel keeps 7 stage vectors: 2 slots per lane avoid spills
        sysm, keep = beam.make_system(B, drag=drag, gravity=grav, u_const=uc, impulse=impulse, gain=gain, ref=ref,
                                      time_inputs=tv)
        sysm, keep, _ = _with_selection(beam, sysm, keep, sel)
    dev = beam.device
    t = torch.full((B,), t0, dtype=torch.float64, device=dev)
    hh = torch.full((B,), float(first_step) if first_step else 0.0, dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    counters = torch.zeros((B, 3), dtype=torch.int64, device=dev)
    if t_eval is not None:
        te = np.ascontiguousarray(np.asarray(t_eval, dtype=np.float64))
        if te.ndim != 1:
            raise ValueError("`t_eval` must be 1-dimensional.")
        if np.any(te < t0) or np.any(te > tf):
            raise ValueError("Values in `t_eval` are not within `t_span`.")
        if np.any(np.diff(te) <= 0):
            raise ValueError("Values in `t_eval` are not properly sorted.")
        d_te = torch.from_numpy(te).to(dev)
        Y = torch.zeros((len(te), B, n2 if needs_unfused else width), dtype=torch.float64, device=dev)
    else:
        te, d_te, Y = np.zeros(0), None, None
    if needs_unfused:
        t, hh, status, counters = _rk45_unfused(beam, X, t0, tf, rtol, atol, u, controller, te if Y is not None else None,
                                                Y, int(max_attempts), first_step)
    else:
        def launch(budget):
            with torch.cuda.device(dev):
                rc = _lib.load().crb_rk45(
                    C.byref(beam._plan), C.byref(sysm), X.data_ptr(), t.data_ptr(), hh.data_ptr(), tf, float(rtol), float(atol),
                    d_te.data_ptr() if d_te is not None else None, len(te), Y.data_ptr() if Y is not None else None,
                    status.data_ptr(), counters.data_ptr(), int(budget), beam._stream(),
                )
            _lib.check(rc)

        # Longest members first.  Members take different numbers of attempts (config 4: 92 on average, 178 at most) and
        # a launch ends with its stragglers; when the ensemble is a few waves of warps, a PILOT launch of a few attempts
        # per member (no work is repeated: the second launch resumes every member exactly where the budget stopped it)
        # tells how many steps each member still wants, (tf - t) / h, and the second launch hands the members out in
        # that order (crb_system_t.member_order).
        pilot = _pilot_attempts(beam, B, pilot_attempts, max_attempts)
        if pilot > 0:
            launch(pilot)
            remaining = torch.where(status == 1, (tf - t) / hh.abs().clamp_min(1e-300), torch.zeros_like(t))
            order = torch.argsort(remaining, descending=True).to(torch.int32)
            sysm.member_order = order.data_ptr()
            keep = list(keep) + [order]
            launch(int(max_attempts) - pilot)
        else:
            launch(max_attempts)
        hh.abs_()  # (a budget stop after a rejected attempt leaves -h: the C ABI's resume flag, of no use to the caller here)
    if Y is not None:
        if needs_unfused and sel is not None:
            Y = Y[:, :, d_sel]
        y = Y.permute(1, 2, 0).contiguous()
        tt = te
    else:  # no common time grid across members: only the final state (SciPy would return every accepted step)
        y = (X if sel is None else X[:, d_sel]).unsqueeze(-1)
        tt = np.array([tf])
    ok = bool((status == 0).all().item())
    msg = MESSAGES[0] if ok else MESSAGES.get(-1 if bool((status == -1).any().item()) else 1, "failed")
    res = EnsembleResult(tt, y[0] if squeeze else y, counters[:, 0].clone(), status, ok, msg,
                         naccept=counters[:, 1].clone(), nreject=counters[:, 2].clone())
    res.t_final = t
    res.h_last = hh
    res.x_final = X
    res.rows = sel
    return res
