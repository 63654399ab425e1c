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
from .dynamic_beam import BatchedDynamicEulerBernoulliBeam, TipImpulse

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
              controller=None, Y_out=None, save_every: int = 0, system=None):
    """Advance X[B,2n] in place by ``nsteps`` classical RK4 steps in ONE kernel launch.

    ``system``: a prebuilt ``(crb_system_t, keepalive)`` from ``beam.make_system`` to skip the
    per-call struct fill (used by bench.py).
    """
    import torch

    if system is None:
        drag, grav, user = beam._active_forces()
        if user or beam._forces_func is not None:
            raise TypeError("fused RK4 supports built-in forces only; use solve_ensemble(...) for torch force callables")
        impulse = u if isinstance(u, TipImpulse) else None
        uc = None if (u is None or impulse is not None) else u
        if callable(uc) and not isinstance(uc, torch.Tensor):
            raise TypeError("fused RK4 needs a constant tensor or TipImpulse input; use solve_ensemble for callables")
        gain, ref = _feedback(controller)
        beam = _feedback_layout(beam, controller)
        system = beam.make_system(X.shape[0], drag=drag, gravity=grav, u_const=uc, impulse=impulse, gain=gain, ref=ref)
    sysm, _keep = system
    with torch.cuda.device(beam.device):
        rc = _lib.load().crb_rk4(
            C.byref(beam._plan), C.byref(sysm), X.data_ptr(), float(t0), float(h), int(nsteps),
            Y_out.data_ptr() if Y_out is not None else None, int(save_every), beam._stream(),
        )
    _lib.check(rc)
    return X


def midpoint_steps(beam: BatchedDynamicEulerBernoulliBeam, X, t0: float, h: float, nsteps: int, *, u=None,
                   Y_out=None, save_every: int = 0):
    """Advance X[B,2n] in place by ``nsteps`` implicit-midpoint steps (Newmark average acceleration) in ONE
    kernel launch (crb_midpoint).  All-linear beams without drag / gravity / feedback; ``u``: constant
    tensor [B,n] or TipImpulse, evaluated at the step midpoints.  Unconditionally stable: ``h`` is chosen
    for accuracy (the rule is second order), not for the highest element frequency as with RK4 -- the
    reason the reference's examples integrate with LSODA (examples/example_utilities.py:153-159)."""
    import torch

    drag, grav, user = beam._active_forces()
    if user or beam._forces_func is not None or drag is not None or grav is not None:
        raise TypeError("the implicit midpoint rule supports force-free all-linear beams (inputs: tensor or TipImpulse)")
    impulse = u if isinstance(u, TipImpulse) else None
    uc = None if (u is None or impulse is not None) else u
    if callable(uc) and not isinstance(uc, torch.Tensor):
        raise TypeError("implicit midpoint needs a constant tensor or TipImpulse input")
    sysm, _keep = beam.make_system(X.shape[0], u_const=uc, impulse=impulse)
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


def _rk4_unfused(beam, X, t0, h, nsteps, u, controller, Y_out, save_every):
    """RK4 with user torch force / input callables: one crb_rhs launch per stage."""
    import torch

    gain, ref = _feedback(controller)

    def f(t, x):
        impulse = u if isinstance(u, TipImpulse) else None
        force = None if impulse is not None else (u(t) if callable(u) else u)
        if force is not None and force.ndim == 1:
            force = force.unsqueeze(0)
        return beam._rhs(t, x, u=force, impulse=impulse, gain=gain, ref=ref)

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


def solve_ensemble(beam: BatchedDynamicEulerBernoulliBeam, t_span: Sequence[float], X0, *, method: str = "RK45",
                   h: Optional[float] = None, t_eval=None, rtol: float = 1e-3, atol: float = 1e-6, u=None,
                   controller=None, save_every: Optional[int] = None, max_attempts: int = 10_000_000,
                   first_step: Optional[float] = None) -> EnsembleResult:
    """Integrate every member of the ensemble over ``t_span``.

    RK4: ``h`` is required; ``nsteps = round((tf - t0) / h)``; outputs are stored every
    ``save_every`` steps (or at ``t_eval`` if its points fall on the step grid).
    RK45: SciPy semantics -- with ``t_eval`` the dense output is sampled, without it only the
    final state is returned (the per-member step sequences differ, so there is no common grid).
    """
    import torch

    if beam.system_func is None or beam.input_func is None:
        raise RuntimeError("System and input functions must be created first")
    X, squeeze = beam._as_state(X0)
    X = X.clone()
    B, n2 = X.shape
    t0, tf = float(t_span[0]), float(t_span[1])
    if not tf > t0:
        raise ValueError("t_span must be increasing")
    method = method.upper()
    drag, grav, user = beam._active_forces()
    needs_unfused = bool(user) or beam._forces_func is not None or (callable(u) and not isinstance(u, (TipImpulse, torch.Tensor)))

    if method in ("RK4", "MIDPOINT"):
        if h is None or not h > 0:
            raise ValueError(f"{method} needs a positive step h")
        nsteps = int(round((tf - t0) / h))
        if t_eval is not None:
            te = np.asarray(t_eval, dtype=np.float64)
            k = np.rint((te - t0) / h).astype(np.int64)
            if np.any(np.abs(t0 + k * h - te) > 1e-9 * max(1.0, abs(tf))) or np.any(k < 0) or np.any(k > nsteps):
                raise ValueError("for RK4, t_eval must lie on the step grid t0 + k*h")
            ks = k[k > 0]
            se = int(np.gcd.reduce(ks)) if len(ks) else nsteps
            save_every = se
        se = int(save_every) if save_every else nsteps
        nframes = nsteps // se
        Y = torch.empty((nframes, B, n2), dtype=torch.float64, device=beam.device) if nframes else None
        if method == "MIDPOINT":
            if needs_unfused or controller is not None:
                raise TypeError("MIDPOINT supports force-free all-linear beams with tensor / TipImpulse inputs only")
            midpoint_steps(beam, X, t0, h, nsteps, u=u, Y_out=Y, save_every=se)
        elif needs_unfused:
            _rk4_unfused(beam, X, t0, h, nsteps, u, controller, Y, se)
        else:
            rk4_steps(beam, X, t0, h, nsteps, u=u, controller=controller, Y_out=Y, save_every=se)
        tt = t0 + se * h * np.arange(1, nframes + 1)
        frames = torch.cat([X0.reshape(1, B, n2).to(Y.dtype), Y], dim=0) if Y is not None else X0.reshape(1, B, n2)
        tt = np.concatenate([[t0], tt])
        if t_eval is not None:
            sel = np.rint((np.asarray(t_eval) - t0) / (se * h)).astype(np.int64)
            frames = frames[torch.as_tensor(sel, device=frames.device)]
            tt = tt[sel]
        y = frames.permute(1, 2, 0).contiguous()
        nfev = torch.full((B,), (4 if method == "RK4" else 1) * nsteps, dtype=torch.int64, device=beam.device)
        status = torch.zeros(B, dtype=torch.int32, device=beam.device)
        return EnsembleResult(tt, y[0] if squeeze else y, nfev, status, True, MESSAGES[0])

    if method != "RK45":
        raise ValueError(f"method must be 'RK4', 'RK45' or 'MIDPOINT', got {method!r} (LSODA is out of scope)")
    if needs_unfused:
        raise TypeError("RK45 supports built-in forces, tensor/TipImpulse inputs and FullStateLinear feedback only")
    impulse = u if isinstance(u, TipImpulse) else None
    uc = None if impulse is not None else u
    gain, ref = _feedback(controller)
    rk45_slots = int(os.environ.get("CRB_RK45_SLOTS", "2"))
    if beam.n_elements <= 64 and rk45_slots > 0:
        beam = beam.with_slots(rk45_slots)  # the adaptive kernel keeps 7 stage vectors: 2 slots per lane avoid spills
    sysm, keep = beam.make_system(B, drag=drag, gravity=grav, u_const=uc, impulse=impulse, gain=gain, ref=ref)
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
        Y = torch.zeros((len(te), B, n2), dtype=torch.float64, device=dev)
    else:
        te, d_te, Y = np.zeros(0), None, None
    with torch.cuda.device(dev):
        rc = _lib.load().crb_rk45(
            C.byref(beam._plan), C.byref(sysm), X.data_ptr(), t.data_ptr(), hh.data_ptr(), tf, float(rtol), float(atol),
            d_te.data_ptr() if d_te is not None else None, len(te), Y.data_ptr() if Y is not None else None,
            status.data_ptr(), counters.data_ptr(), int(max_attempts), beam._stream(),
        )
    _lib.check(rc)
    if Y is not None:
        y = Y.permute(1, 2, 0).contiguous()
        tt = te
    else:
        y = X.unsqueeze(-1)
        tt = np.array([tf])
    ok = bool((status == 0).all().item())
    worst = int(status.abs().max().item()) if not ok else 0
    msg = MESSAGES[0] if ok else MESSAGES.get(-1 if bool((status == -1).any().item()) else 1, "failed")
    res = EnsembleResult(tt, y[0] if squeeze else y, counters[:, 0].clone(), status, ok, msg,
                         naccept=counters[:, 1].clone(), nreject=counters[:, 2].clone())
    res.t_final = t
    res.h_last = hh
    res.x_final = X
    return res
