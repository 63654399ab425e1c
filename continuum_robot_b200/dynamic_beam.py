"""Batched beam system: host-side mirror of the reference's beam-system API.

Reference classes mirrored (paths under /root/reference/src/continuum_robot/):
  * ``EulerBernoulliBeam``          models/euler_bernoulli_beam.py:16-511  -> ``BatchedEulerBernoulliBeam``
  * ``DynamicEulerBernoulliBeam``   models/dynamic_beam_model.py:16-364    -> ``BatchedDynamicEulerBernoulliBeam``
  * ``FluidDragForce``/``GravityForce`` models/fluid_forces.py:24-142, gravity_forces.py:6-173

Same method names, argument meaning and exception types; arrays are torch CUDA float64 tensors
with a leading member axis (``x[B, 2n]``, ``u[B, n]``).  All numerics run in libcrb.so
(hand-written sm_100a CUDA, include/crb.h); this module only validates, owns device buffers and
fills the C structs.  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import pathlib
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np

from . import _lib
from .abstractions import AbstractForce, BoundaryConditionType, ElementType, Properties
from .force_params import ForceParams
from .force_registry import ForceRegistry, InputRegistry

_PARAM_COLS = ["length", "elastic_modulus", "moment_inertia", "density", "cross_area"]
_FLUID_COLS = ["wetted_area", "drag_coef"]
_BC_CODE = {"NONE": 0, "FIXED": 1, "PINNED": 2}
_DOF_NAMES = ("u", "w", "phi")


def _torch():
    import torch

    return torch


# --------------------------------------------------------------------------------------------
# Built-in force components (fused into the RHS kernel)
# --------------------------------------------------------------------------------------------
class FluidDragForce(AbstractForce):
    """Quadratic drag on transverse DOFs (models/fluid_forces.py:24-142); fused in-kernel."""

    fused_kind = "drag"

    def __init__(self, beam: "BatchedDynamicEulerBernoulliBeam", fluid_density: float, enabled: bool = True):
        self._beam = beam
        self.fluid_density = fluid_density
        self.enabled = enabled

    def is_enabled(self) -> bool:
        return self.enabled

    def compute_forces(self, x, t):
        return self._beam._builtin_forces(x, drag=self, gravity=None)


class GravityForce(AbstractForce):
    """Rotation-aware gravity with the reference's reduced-index placement
    (models/gravity_forces.py:66-148); fused in-kernel."""

    fused_kind = "gravity"

    def __init__(self, beam: "BatchedDynamicEulerBernoulliBeam", gravity_vector=None, enabled: bool = True):
        self._beam = beam
        self.gravity_vector = np.array(gravity_vector if gravity_vector is not None else [0.0, -9.81, 0.0], dtype=float)
        if len(self.gravity_vector) != 3:
            raise ValueError("Gravity vector must have exactly 3 components [gx, gy, gz]")
        self.enabled = enabled

    def is_enabled(self) -> bool:
        return self.enabled

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = enabled

    def set_gravity_vector(self, gravity_vector) -> None:
        if len(gravity_vector) != 3:
            raise ValueError("Gravity vector must have exactly 3 components [gx, gy, gz]")
        self.gravity_vector = np.array(gravity_vector, dtype=float)

    def get_gravity_vector(self) -> np.ndarray:
        return self.gravity_vector.copy()

    def compute_forces(self, x, t):
        return self._beam._builtin_forces(x, drag=None, gravity=self)


class TipImpulse:
    """Parametrised input u(t): ``amplitude[b]`` on reduced DOF ``dof`` while ``t < duration``.

    The examples' disturbance (examples/example_utilities.py:144-148, lqr_control.py:32-43),
    evaluated inside the kernel at every stage time.
    """

    def __init__(self, amplitude, dof: int = -2, duration: float = 0.01):
        self.amplitude = amplitude
        self.dof = dof
        self.duration = float(duration)

    def as_tensor(self, t: float, n: int):
        """u(t)[B, n] as a tensor (for code that wants the reference's callable-input form)."""
        torch = _torch()
        amp = self.amplitude
        u = torch.zeros(amp.shape[0], n, dtype=torch.float64, device=amp.device)
        if t < self.duration:
            u[:, self.dof] = amp
        return u


class SinusoidInput:
    """Time-varying input ``u(t)[b, r] = amplitude[b, r] * sin(omega * t + phase)`` fused into the RHS kernels
    (crb_system_t.u_sin_*): the form the reference's own integration tests drive the beam with,
    ``u(t) = np.sin(t) * np.ones(n)`` (tests/test_dynamic_beam.py:214-215, 234-235).

    ``amplitude``: torch tensor ``[n]`` (shared by the ensemble) or ``[B, n]``.  Also callable, ``u(t)``, so it can
    be passed wherever the reference passes a callable input (``get_dynamic_system()(t, X, u)``).
    """

    def __init__(self, amplitude, omega: float = 1.0, phase: float = 0.0):
        torch = _torch()
        if not isinstance(amplitude, torch.Tensor):
            amplitude = torch.as_tensor(np.asarray(amplitude, dtype=np.float64))
        if amplitude.ndim not in (1, 2):
            raise ValueError("SinusoidInput amplitude must have shape [n] or [B, n]")
        self.amplitude = amplitude
        self.omega = float(omega)
        self.phase = float(phase)

    def __call__(self, t):
        torch = _torch()
        tt = torch.as_tensor(t, dtype=torch.float64, device=self.amplitude.device)
        if tt.ndim == 1:  # per-member times [B] -> [B, 1]
            tt = tt.unsqueeze(-1)
        return self.amplitude * torch.sin(self.omega * tt + self.phase)


class PiecewiseLinearInput:
    """Time-varying input interpolated linearly from a table (``numpy.interp`` semantics: held constant outside
    the knots), fused into the RHS kernels (crb_system_t.u_tab_*).  ``times``: ascending ``[K]`` (K >= 2);
    ``values``: torch tensor ``[K, n]`` (shared) or ``[K, B, n]``."""

    def __init__(self, times, values):
        torch = _torch()
        tk = np.ascontiguousarray(np.asarray(times, dtype=np.float64))
        if tk.ndim != 1 or len(tk) < 2 or np.any(np.diff(tk) <= 0):
            raise ValueError("PiecewiseLinearInput times must be a strictly ascending 1-D array with at least 2 knots")
        if not isinstance(values, torch.Tensor):
            values = torch.as_tensor(np.asarray(values, dtype=np.float64))
        if values.ndim not in (2, 3) or values.shape[0] != len(tk):
            raise ValueError("PiecewiseLinearInput values must have shape [K, n] or [K, B, n]")
        self.times = tk
        self.values = values

    def __call__(self, t):
        torch = _torch()
        v = self.values
        tk = torch.as_tensor(self.times, dtype=torch.float64, device=v.device)
        tt = torch.as_tensor(t, dtype=torch.float64, device=v.device).reshape(-1)  # [1] or [B]
        tc = tt.clamp(float(self.times[0]), float(self.times[-1]))
        k0 = (torch.searchsorted(tk, tc, right=True) - 1).clamp(0, len(self.times) - 2)
        w = ((tc - tk[k0]) / (tk[k0 + 1] - tk[k0])).unsqueeze(-1)
        if v.ndim == 2:
            out = v[k0] + w * (v[k0 + 1] - v[k0])  # [len(tt), n]
        else:
            b = torch.arange(v.shape[1], device=v.device)
            out = v[k0, b] + w * (v[k0 + 1, b] - v[k0, b])
        return out[0] if (v.ndim == 2 and tt.numel() == 1 and np.ndim(t) == 0) else out


def split_input(u):
    """Decompose an input ``u`` into the parts the kernels fuse: (constant tensor or None, TipImpulse or None,
    [SinusoidInput / PiecewiseLinearInput ...], other callable or None).  ``u`` may be one part or a list / tuple of
    parts (their sum)."""
    torch = _torch()
    parts = list(u) if isinstance(u, (list, tuple)) else ([] if u is None else [u])
    uc, imp, tv, other = None, None, [], None
    for part in parts:
        if isinstance(part, TipImpulse):
            if imp is not None:
                raise ValueError("at most one TipImpulse per input")
            imp = part
        elif isinstance(part, (SinusoidInput, PiecewiseLinearInput)):
            if any(type(x) is type(part) for x in tv):
                raise ValueError(f"at most one {type(part).__name__} per input")
            tv.append(part)
        elif isinstance(part, torch.Tensor):
            uc = part if uc is None else uc + part
        elif callable(part):
            if other is not None:
                raise ValueError("at most one free-form callable per input")
            other = part
        else:
            raise TypeError(f"input parts must be torch tensors, TipImpulse, SinusoidInput, PiecewiseLinearInput or "
                            f"torch callables u(t), got {type(part).__name__}")
    return uc, imp, tv, other


# --------------------------------------------------------------------------------------------
# Parameter ingestion
# --------------------------------------------------------------------------------------------
def _frame_from(obj):
    import pandas as pd

    if isinstance(obj, (str, pathlib.Path)):
        try:
            return pd.read_csv(obj)
        except FileNotFoundError:
            raise FileNotFoundError(f"Parameter file {obj} not found")
    if isinstance(obj, pd.DataFrame):
        return obj.copy()
    raise TypeError("Parameters must be filepath or pandas DataFrame")


def _validate_frame(df, fluid: bool, need_bc: bool) -> None:
    """Column / value checks of dynamic_beam_model.py:76-118 and euler_bernoulli_beam.py:83-109."""
    required = _PARAM_COLS + ["type"] + (["boundary_condition"] if need_bc else [])
    if fluid:
        required = required + _FLUID_COLS
    if not all(c in df.columns for c in required):
        raise ValueError(f"CSV must contain columns: {', '.join(required)}")
    valid = {t.value for t in ElementType}
    bad = set(df["type"].str.lower()) - valid
    if bad:
        raise ValueError(f"Invalid element types: {bad}")
    if need_bc:
        badbc = set(df["boundary_condition"]) - set(_BC_CODE)
        if badbc:
            raise ValueError(f"Invalid boundary conditions: {badbc}")
    if (df[_PARAM_COLS] <= 0).any().any():
        raise ValueError("All numeric parameters must be positive")
    if fluid:
        if (df["drag_coef"] < 0).any():
            raise ValueError("Drag coefficients cannot be negative")
        if (df["wetted_area"] < 0).any():
            raise ValueError("Wetted areas cannot be negative")


def _frame_to_arrays(df, fluid: bool):
    N = len(df)
    par = np.zeros((N, 7), dtype=np.float64)
    for k, c in enumerate(_PARAM_COLS):
        par[:, k] = df[c].to_numpy(dtype=np.float64)
    for k, c in enumerate(_FLUID_COLS):
        if c in df.columns:
            par[:, 5 + k] = df[c].to_numpy(dtype=np.float64)
    et = np.array([0 if t.lower() == "linear" else 1 for t in df["type"]], dtype=np.uint8)
    bc = np.zeros(N + 1, dtype=np.uint8)
    if "boundary_condition" in df.columns:
        bc[:N] = [_BC_CODE[b] for b in df["boundary_condition"]]  # row i -> node i (Q6)
    return par, et, bc


class BatchedEulerBernoulliBeam:
    """Assembly-level view (``beam.beam_model`` of the reference): matrices and DOF maps."""

    def __init__(self, owner: "BatchedDynamicEulerBernoulliBeam"):
        self._o = owner
        self.parameters = owner.params
        self.segments = [
            Properties(
                length=float(r["length"]), elastic_modulus=float(r["elastic_modulus"]),
                moment_inertia=float(r["moment_inertia"]), density=float(r["density"]),
                cross_area=float(r["cross_area"]), segment_id=i, element_type=str(r["type"]),
                wetted_area=float(r["wetted_area"]) if "wetted_area" in owner.params.columns else None,
                drag_coef=float(r["drag_coef"]) if "drag_coef" in owner.params.columns else None,
            )
            for i, (_, r) in enumerate(owner.params.iterrows())
        ] if owner.params is not None else []
        self.dof_to_node_param = dict(owner._pos_map)
        self.node_param_to_dof = {v: k for k, v in owner._pos_map.items()}

    # -- matrices (host copies for LQR synthesis / inspection) --------------------------------
    def _dense(self, member: int, want_k: bool):
        o = self._o
        n = o.n_free
        par = np.ascontiguousarray(o._params_np[min(member, o._params_np.shape[0] - 1)])
        M = np.zeros((n, n))
        K = np.zeros((n, n)) if want_k else None
        rc = _lib.load().crb_dense_matrices(
            C.byref(o._plan), par.ctypes.data_as(C.c_void_p), o._etype.tobytes(), o._bc.tobytes(),
            M.ctypes.data_as(C.c_void_p), K.ctypes.data_as(C.c_void_p) if want_k else None,
        )
        _lib.check(rc, ValueError)
        return M, K

    def get_mass_matrix(self, member: int = 0) -> np.ndarray:
        return self._dense(member, False)[0]

    def get_stiffness_matrix(self, member: int = 0) -> np.ndarray:
        """Dense BC-reduced K; raises ValueError on nonlinear segments
        (euler_bernoulli_beam.py:422-456)."""
        return self._dense(member, True)[1]

    @property
    def M(self):
        return self.get_mass_matrix(0)

    def get_constrained_dofs(self) -> List[int]:
        return list(self._o.constrained_dofs)

    def get_segment_count(self) -> int:
        return self._o.n_elements

    def get_segment_types(self) -> List[ElementType]:
        return [ElementType.LINEAR if t == 0 else ElementType.NONLINEAR for t in self._o._etype]

    def is_hybrid(self) -> bool:
        return len(set(self._o._etype.tolist())) > 1

    def get_length(self) -> float:
        return float(self._o._params_np[0, :, 0].sum())

    def get_dof_index(self, node_idx: int, param: str) -> int:
        if (param, node_idx) not in self.node_param_to_dof:
            raise KeyError(f"Invalid node/parameter combination: ({node_idx}, {param})")
        return self.node_param_to_dof[(param, node_idx)]

    def get_dof_to_node_param(self, dof_idx: int):
        if dof_idx not in self.dof_to_node_param:
            raise KeyError(f"Invalid DOF index: {dof_idx}")
        return self.dof_to_node_param[dof_idx]


class BatchedDynamicEulerBernoulliBeam:
    """Ensemble of Euler-Bernoulli beams sharing one topology (element types + BCs).

    ``params`` may be
      * a CSV path or DataFrame (one design; the ensemble size is taken from the state batch),
      * a list of DataFrames / CSV paths (one per member; parsed values are used),
      * a dict ``{"params": array[Bp,N,7], "type": [N] str|int, "boundary_condition": [N] str|int}``
        for synthetic ensembles (columns: length, elastic_modulus, moment_inertia, density,
        cross_area, wetted_area, drag_coef).
    """

    def __init__(self, params, force_params: Optional[ForceParams] = None, *, device=None,
                 max_slots_per_lane: int = 0):
        torch = _torch()
        self._ctor = (params, max_slots_per_lane)
        self._siblings = {}
        self.force_params = force_params or ForceParams()
        fluid = bool(self.force_params.enable_fluid_effects)
        if fluid and self.force_params.fluid_density <= 0:
            raise ValueError("Fluid density must be positive")
        self.params = None
        if isinstance(params, dict):
            par = np.ascontiguousarray(np.asarray(params["params"], dtype=np.float64))
            if par.ndim == 2:
                par = par[None]
            if par.ndim != 3 or par.shape[2] != 7:
                raise ValueError("params array must have shape [Bp, N, 7]")
            N = par.shape[1]
            et = np.array([(0 if str(t).lower() in ("0", "linear") else 1) for t in params["type"]], dtype=np.uint8)
            bcs = params.get("boundary_condition", ["FIXED"] + ["NONE"] * (N - 1))
            bc = np.zeros(N + 1, dtype=np.uint8)
            bc[: len(bcs)] = [(_BC_CODE[b] if isinstance(b, str) else int(b)) for b in bcs]
            if len(et) != N:
                raise ValueError("type must have one entry per element")
            if (par[:, :, :5] <= 0).any():
                raise ValueError("All numeric parameters must be positive")
            if fluid and (par[:, :, 5:] < 0).any():
                raise ValueError("Drag coefficients cannot be negative")
            import pandas as pd

            df = pd.DataFrame({c: par[0, :, k] for k, c in enumerate(_PARAM_COLS + _FLUID_COLS)})
            df["type"] = ["linear" if t == 0 else "nonlinear" for t in et]
            inv = {v: k for k, v in _BC_CODE.items()}
            df["boundary_condition"] = [inv[int(b)] for b in bc[:N]]
            self.params = df
        else:
            frames = [_frame_from(p) for p in params] if isinstance(params, (list, tuple)) else [_frame_from(params)]
            for df in frames:
                _validate_frame(df, fluid, need_bc=True)
            arrs = [_frame_to_arrays(df, fluid) for df in frames]
            et, bc = arrs[0][1], arrs[0][2]
            for a in arrs[1:]:
                if a[0].shape != arrs[0][0].shape or not np.array_equal(a[1], et) or not np.array_equal(a[2], bc):
                    raise ValueError("All members of an ensemble must share element types and boundary conditions")
            par = np.ascontiguousarray(np.stack([a[0] for a in arrs]))
            self.params = frames[0]
        self._params_np = par
        self._etype = np.ascontiguousarray(et)
        self._bc = np.ascontiguousarray(bc)
        self.n_elements = int(par.shape[1])
        self.n_param_sets = int(par.shape[0])
        if int((bc != 0).sum()) == self.n_elements + 1:
            raise ValueError("Cannot constrain all nodes with boundary conditions")
        self.boundary_conditions = {
            i: (BoundaryConditionType.FIXED if b == 1 else BoundaryConditionType.PINNED)
            for i, b in enumerate(bc) if b != 0
        }
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("continuum_robot_b200 runs on CUDA devices only (no CPU fallback)")

        lib = _lib.load()
        self._plan = _lib.CrbPlan()
        _lib.check(lib.crb_plan(self.n_elements, self._bc.tobytes(), int(max_slots_per_lane), C.byref(self._plan)), ValueError)
        self.n_free = int(self._plan.n_free)

        # constrained DOFs / state maps (euler_bernoulli_beam.py:240-262, dynamic_beam_model.py:120-149)
        self.constrained_dofs = []
        self._pos_map: Dict[int, tuple] = {}
        r = 0
        for node in range(self.n_elements + 1):
            for d in range(3):
                if bc[node] == 1 or (bc[node] == 2 and d < 2):
                    self.constrained_dofs.append(3 * node + d)
                else:
                    self._pos_map[r] = (_DOF_NAMES[d], node)
                    r += 1
        n = self.n_free
        self.state_to_node_param = dict(self._pos_map)
        self.state_to_node_param.update({i + n: (f"d{p}_dt", node) for i, (p, node) in self._pos_map.items()})
        self.node_param_to_state = {v: k for k, v in self.state_to_node_param.items()}

        self._assemble()
        self.beam_model = BatchedEulerBernoulliBeam(self)
        self.system_func = None
        self.input_func = None
        self._forces_func = None
        self.force_registry = ForceRegistry()
        self.input_registry = InputRegistry()
        self._auto_register_forces()

    @property
    def M_inv(self):
        """Dense inverse of member 0's mass matrix on the HOST, for inspection only (the reference
        exposes ``M_inv``, dynamic_beam_model.py:60).  The integration path never forms it: the
        kernels apply the banded factorisation written by crb_assemble."""
        return np.linalg.inv(self.beam_model.get_mass_matrix(0))

    # -- device assembly ----------------------------------------------------------------------
    def _assemble(self) -> None:
        torch = _torch()
        lib = _lib.load()
        par = self._params_np
        Bp = par.shape[0]

        def shared(cols):
            return Bp == 1 or bool(np.all(par[:, :, cols] == par[0:1, :, cols]))

        # dispatch hints for the fast kernels: all-linear beam, one (rho A L, L) for every element
        self._all_linear = bool(np.all(self._etype == 0))
        L0, rho0, A0 = par[0, 0, 0], par[0, 0, 3], par[0, 0, 4]
        self.force_general_kernels = False
        self.force_staged_kernels = False
        self._mass_shared = shared([0, 3, 4])
        self._stiff_shared = shared([0, 1, 2, 4])
        self._force_shared = shared([0, 3, 4, 5, 6])
        dev = self.device
        P = int(self._plan.p)
        n_mass = 1 if self._mass_shared else Bp
        n_stiff = 1 if self._stiff_shared else Bp
        n_force = 1 if self._force_shared else Bp
        self._d_params = torch.from_numpy(par).to(dev)
        self._d_mfac = torch.empty(n_mass * int(self._plan.mfac_doubles), dtype=torch.float64, device=dev)
        self._d_kcoef = torch.empty(n_stiff * P * 4, dtype=torch.float64, device=dev)
        self._d_etype = torch.empty(P, dtype=torch.uint8, device=dev)
        self._d_drag = torch.zeros(n_force * P, dtype=torch.float64, device=dev)
        self._d_grav = torch.zeros(n_force * P * 2, dtype=torch.float64, device=dev)
        self._d_hmass = torch.zeros(n_force * self.n_elements, dtype=torch.float64, device=dev)
        self._d_red = torch.from_numpy(np.ctypeslib.as_array(self._plan.red_index)[: 3 * P].astype(np.int32)).to(dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            rc = lib.crb_assemble(
                C.byref(self._plan), self._d_params.data_ptr(), Bp, self._etype.tobytes(), self._bc.tobytes(),
                n_mass, n_stiff, n_force, float(self.force_params.fluid_density or 0.0),
                self._d_mfac.data_ptr(), self._d_kcoef.data_ptr(), self._d_etype.data_ptr(),
                self._d_drag.data_ptr(), self._d_grav.data_ptr(), self._d_hmass.data_ptr(), stream,
            )
        _lib.check(rc)

    def _auto_register_forces(self) -> None:
        """dynamic_beam_model.py:220-241: drag first, then gravity."""
        if self.force_params.enable_fluid_effects:
            self.force_registry.register(FluidDragForce(self, self.force_params.fluid_density, True))
        if self.force_params.enable_gravity_effects:
            self.force_registry.register(GravityForce(self, self.force_params.get_gravity_vector(), True))

    def dense_matrices(self, want_stiffness: bool = True):
        """BC-reduced dense ``M[Bp,n,n]`` and ``K[Bp,n,n]`` of every parameter set, built on the device
        (crb_dense_matrices_batched): the per-member A / B build of an LQR design ensemble
        (``beam.beam_model.get_mass_matrix()`` / ``get_stiffness_matrix()`` of the reference, for all members
        at once; euler_bernoulli_beam.py:357-361, 422-511).  ``Bp`` = 1 for one shared design."""
        torch = _torch()
        n, Bp = self.n_free, int(self._params_np.shape[0])
        par = torch.from_numpy(np.ascontiguousarray(self._params_np)).to(self.device)
        M = torch.empty((Bp, n, n), dtype=torch.float64, device=self.device)
        K = torch.empty((Bp, n, n), dtype=torch.float64, device=self.device) if want_stiffness else None
        with torch.cuda.device(self.device):
            rc = _lib.load().crb_dense_matrices_batched(
                C.byref(self._plan), par.data_ptr(), Bp, self._etype.tobytes(), self._bc.tobytes(),
                M.data_ptr(), K.data_ptr() if want_stiffness else None, self._stream())
        _lib.check(rc, ValueError)
        return M, K

    def shifted_factors(self, shift: float):
        """(afac, shared): block-LDL^T factors of M + shift K on the device (crb_assemble_shifted), one set if
        mass and stiffness are shared by all members, else one per member; cached per shift."""
        torch = _torch()
        if not self._all_linear:
            raise ValueError("M + shift K needs an all-linear beam")
        cache = self.__dict__.setdefault("_shift_cache", {})
        key = float(shift)
        if key not in cache:
            if len(cache) >= 2:  # a per-member factor set is as large as 13 KB x members
                cache.clear()
            shared = bool(self._mass_shared and self._stiff_shared)
            Bp = self._params_np.shape[0]
            n_sets = 1 if shared else Bp
            afac = torch.empty((n_sets, int(self._plan.mfac_doubles)), dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                rc = _lib.load().crb_assemble_shifted(
                    C.byref(self._plan), self._d_params.data_ptr(), Bp, self._etype.tobytes(), self._bc.tobytes(),
                    n_sets, key, afac.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check(rc)
            cache[key] = (afac, shared)
        return cache[key]

    def with_slots(self, max_slots_per_lane: int) -> "BatchedDynamicEulerBernoulliBeam":
        """Same ensemble assembled for another lane decomposition (e.g. 2 slots per lane for the
        register-hungry adaptive kernel).  Registries and function state are SHARED with ``self``."""
        if max_slots_per_lane == self._ctor[1] or int(self._plan.m) <= max_slots_per_lane:
            return self
        if max_slots_per_lane not in self._siblings:
            sib = BatchedDynamicEulerBernoulliBeam(
                {"params": self._params_np, "type": [int(t) for t in self._etype],
                 "boundary_condition": [int(b) for b in self._bc[: self.n_elements]]},
                self.force_params, device=self.device, max_slots_per_lane=max_slots_per_lane)
            self._siblings[max_slots_per_lane] = sib
        sib = self._siblings[max_slots_per_lane]
        sib.force_registry, sib.input_registry = self.force_registry, self.input_registry
        sib._forces_func, sib.system_func, sib.input_func = self._forces_func, self.system_func, self.input_func
        sib.force_general_kernels, sib.force_staged_kernels = self.force_general_kernels, self.force_staged_kernels
        sib.use_member_operators = getattr(self, "use_member_operators", True)
        return sib

    # -- state maps -----------------------------------------------------------------------------
    def get_state_to_node_param(self, state_idx):
        if state_idx not in self.state_to_node_param:
            raise KeyError(f"Invalid state index: {state_idx}")
        return self.state_to_node_param[state_idx]

    def get_state_index(self, node_idx, param):
        if (param, node_idx) not in self.node_param_to_state:
            raise KeyError(f"Invalid node/parameter combination: ({node_idx}, {param})")
        return self.node_param_to_state[(param, node_idx)]

    def get_state_mapping(self):
        return dict(self.state_to_node_param)

    def get_node_param_mapping(self):
        return dict(self.node_param_to_state)

    # -- C struct ---------------------------------------------------------------------------------
    def _check_members(self, B: int) -> None:
        if self.n_param_sets not in (1, B):
            raise ValueError(f"state batch has {B} members but the ensemble was built with {self.n_param_sets} parameter sets")

    def _plan_key(self):
        pl = self._plan
        return (int(pl.n_elements), int(pl.n_free), int(pl.m), int(pl.g), int(pl.p), int(pl.contiguous), id(self._d_mfac))

    def make_system(self, B: int, *, drag: Optional[FluidDragForce] = None, gravity: Optional[GravityForce] = None,
                    u_const=None, impulse: Optional[TipImpulse] = None, gain=None, ref=None, f_ext=None,
                    member_range=None, time_inputs=()):
        """Fill a crb_system_t; returns (struct, keepalive list).

        ``member_range=(lo, hi)`` describes the sub-ensemble lo..hi-1 of a B-member ensemble (every
        per-member pointer is offset); used to pipeline host<->device copies chunk by chunk."""
        torch = _torch()
        self._check_members(B)
        n = self.n_free
        keep = []
        s = _lib.CrbSystem()
        s._crb_plan_key = self._plan_key()  # which lane layout the device arrays of this struct were assembled for
        lo, hi = member_range if member_range is not None else (0, B)
        if not 0 <= lo < hi <= B:
            raise ValueError(f"bad member_range {member_range} for {B} members")
        s.n_members = hi - lo
        s.mass_shared = int(self._mass_shared)
        s.stiff_shared = int(self._stiff_shared)
        s.force_shared = int(self._force_shared)
        P = int(self._plan.p)
        s.mfac = self._d_mfac.data_ptr() + (0 if self._mass_shared else 8 * lo * int(self._plan.mfac_doubles))
        s.kcoef = self._d_kcoef.data_ptr() + (0 if self._stiff_shared else 8 * lo * P * 4)
        fo = 0 if self._force_shared else lo
        s.elem_type = self._d_etype.data_ptr()
        s.red_index = self._d_red.data_ptr()
        s.all_linear = int(self._all_linear)
        s.all_nonlinear = int(bool(np.all(self._etype == 1)))
        s.force_general = int(self.force_general_kernels)
        s.force_staged = int(self.force_staged_kernels)
        tc = self.__dict__.get("_tile_counter")
        if tc is None:  # ticket counter of the persistent kernels (one per beam object: its launches share a stream)
            tc = self.__dict__["_tile_counter"] = torch.zeros(1, dtype=torch.int32, device=self.device)
        s.tile_counter = tc.data_ptr()
        if drag is not None:
            if drag.fluid_density != self.force_params.fluid_density:
                raise ValueError("FluidDragForce.fluid_density differs from the assembled ForceParams.fluid_density")
            s.drag = self._d_drag.data_ptr() + 8 * fo * P
        if gravity is not None:
            gv = gravity.gravity_vector
            s.gx, s.gy = float(gv[0]), float(gv[1])
            s.seg_half_mass = self._d_hmass.data_ptr() + 8 * fo * self.n_elements
            if self._plan.contiguous:
                s.grav = self._d_grav.data_ptr() + 8 * fo * P * 2
                s.grav_mode = 1
            else:
                s.grav_mode = 2
        def dev64(tn, shape, name):
            if not isinstance(tn, torch.Tensor):
                raise TypeError(f"{name} must be a torch tensor on {self.device} (no CPU path)")
            tn = tn.to(device=self.device, dtype=torch.float64)
            if tn.shape != shape:
                tn = tn.expand(shape)
            tn = tn.contiguous()
            keep.append(tn)
            return tn
        if u_const is not None:
            s.u_const = dev64(u_const, (B, n), "u").data_ptr() + 8 * lo * n
        if f_ext is not None:
            s.f_ext = dev64(f_ext, (B, n), "forces").data_ptr() + 8 * lo * n
        if impulse is not None:
            amp = impulse.amplitude
            if not isinstance(amp, torch.Tensor):
                amp = torch.as_tensor(np.asarray(amp, dtype=np.float64))
            s.imp_amp = dev64(amp.reshape(-1), (B,), "impulse amplitude").data_ptr() + 8 * lo
            dof = impulse.dof if impulse.dof >= 0 else n + impulse.dof
            if not 0 <= dof < n:
                raise ValueError(f"impulse dof {impulse.dof} outside the {n} position DOFs")
            s.imp_dof = dof
            s.imp_duration = impulse.duration
        shared_flags = []
        for part in time_inputs:  # fused time-varying inputs (SinusoidInput / PiecewiseLinearInput)
            if isinstance(part, SinusoidInput):
                a = part.amplitude
                shared_flags.append(a.ndim == 1)
                s.u_sin_amp = dev64(a, (n,) if a.ndim == 1 else (B, n), "sinusoid amplitude").data_ptr() + (0 if a.ndim == 1 else 8 * lo * n)
                s.u_sin_omega, s.u_sin_phase = part.omega, part.phase
            elif isinstance(part, PiecewiseLinearInput):
                v = part.values
                K = len(part.times)
                shared_flags.append(v.ndim == 2)
                if v.ndim == 3 and (lo, hi) != (0, B):
                    raise ValueError("a per-member input table cannot be sliced by member_range")
                s.u_tab_v = dev64(v, (K, n) if v.ndim == 2 else (K, B, n), "input table").data_ptr()
                tk = torch.from_numpy(part.times).to(self.device)
                keep.append(tk)
                s.u_tab_t = tk.data_ptr()
                s.u_tab_k = K
            else:
                raise TypeError(f"unsupported fused input {type(part).__name__}")
        if shared_flags:
            if len(set(shared_flags)) != 1:
                raise ValueError("fused time-varying inputs must all be shared ([n]) or all per member ([B, n])")
            s.u_time_shared = int(shared_flags[0])
        if gain is not None:
            if not isinstance(gain, torch.Tensor):
                gain = torch.as_tensor(np.asarray(gain, dtype=np.float64))
            per_member = gain.ndim == 3
            if per_member:  # one design-specific gain per member (BatchedLinearQuadraticRegulator / crb_lqr_gains)
                if tuple(gain.shape) != (B, n, 2 * n):
                    raise ValueError(f"Per-member gains must have shape ({B}, {n}, {2 * n}), got {tuple(gain.shape)}")
                gdev = dev64(gain, (B, n, 2 * n), "gain")
                s.gain = gdev.data_ptr() + 8 * lo * n * 2 * n
                s.gain_stride = n * 2 * n
                rdev = None
                if ref is not None:
                    if not isinstance(ref, torch.Tensor):
                        ref = torch.as_tensor(np.asarray(ref, dtype=np.float64))
                    rdev = dev64(ref.reshape(-1), (2 * n,), "reference")
                    s.ref = rdev.data_ptr()
                # all-linear designs: the closed loop of every member as one dense operator (crb_member_operators)
                if (getattr(self, "use_member_operators", True) and self._all_linear and drag is None and u_const is None and f_ext is None and not time_inputs and n <= 32
                        and self.n_elements <= 16 and not self.force_general_kernels and not self.force_staged_kernels):
                    op = self._member_operators(gdev, rdev)
                    keep.append(op)
                    s.member_op = op.data_ptr() + 8 * lo * n * (3 * n + 1)
                return s, keep
            if gain.ndim != 2 or tuple(gain.shape) != (n, 2 * n):
                raise ValueError(f"Gain matrix must have shape ({n}, {2 * n}), got {tuple(gain.shape)}")
            gdev = dev64(gain, (n, 2 * n), "gain")
            s.gain = gdev.data_ptr()
            cache_key = (gdev.data_ptr(), gdev._version)
            cached = getattr(self, "_gain_frag_cache", None)
            if int(self._plan.g) == 4 and cached is not None and cached[0] == cache_key:
                keep.append(cached[1])
                s.gain_frag = cached[1].data_ptr()
            elif int(self._plan.g) == 4:  # shared operator -> FP64 tensor-core fragments (crb_gain_fragments)
                lib = _lib.load()
                cnt = int(lib.crb_gain_fragments(C.byref(self._plan), None, None))
                gh = np.ascontiguousarray(gdev.detach().cpu().numpy())
                fr = np.empty(cnt, dtype=np.float64)
                got = lib.crb_gain_fragments(C.byref(self._plan), gh.ctypes.data_as(C.c_void_p), fr.ctypes.data_as(C.c_void_p))
                if got != cnt:
                    _lib.check(int(got))
                frd = torch.from_numpy(fr).to(self.device)
                keep.append(frd)
                s.gain_frag = frd.data_ptr()
                self._gain_frag_cache = (cache_key, frd, gdev)  # gdev kept alive so its address stays unique
            rdev = None
            if ref is not None:
                if not isinstance(ref, torch.Tensor):
                    ref = torch.as_tensor(np.asarray(ref, dtype=np.float64))
                rdev = dev64(ref.reshape(-1), (2 * n,), "reference")
                s.ref = rdev.data_ptr()
            # one design + one gain shared by every member: the closed-loop RHS is a dense contraction
            # with member-independent operators (crb_shared_operator -> FP64 tensor cores)
            if (self._mass_shared and self._stiff_shared and self._all_linear and drag is None and u_const is None
                    and f_ext is None and not time_inputs and n <= 24 and (gravity is None or self.n_elements <= 8)
                    and not self.force_general_kernels):
                blob = self._shared_operator(gdev, rdev, (s.gx, s.gy) if gravity is not None else None,
                                             int(s.imp_dof) if impulse is not None else -1)
                keep.append(blob)
                s.shared_op = blob.data_ptr()
                s.shared_op_doubles = blob.numel()
        return s, keep

    def _member_operators(self, gdev, rdev):
        """[B, n, 3n+1] closed-loop operators of the members for the per-member gains ``gdev`` (crb_member_operators);
        cached on the identity and version of the gain / reference tensors."""
        torch = _torch()
        key = (gdev.data_ptr(), gdev._version, None if rdev is None else (rdev.data_ptr(), rdev._version))
        cached = getattr(self, "_member_op_cache", None)
        if cached is not None and cached[0] == key:
            return cached[1]
        B, n = int(gdev.shape[0]), self.n_free
        M, K = self.dense_matrices()
        op = torch.empty((B, n, 3 * n + 1), dtype=torch.float64, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.load().crb_member_operators(
                n, B, M.data_ptr(), int(M.shape[0] == 1), K.data_ptr(), int(K.shape[0] == 1), gdev.data_ptr(),
                rdev.data_ptr() if rdev is not None else None, op.data_ptr(), status.data_ptr(), self._stream())
        _lib.check(rc, ValueError)
        if bool(status.any().item()):
            raise ValueError("Mass matrix is singular and cannot be inverted")
        self._member_op_cache = (key, op, gdev, rdev)  # tensors kept alive so that their addresses stay unique
        return op

    def _shared_operator(self, gdev, rdev, grav_xy, imp_dof: int):
        """Device copy of the shared-operator blob for (gain, ref, gravity vector, impulse DOF); cached on
        the operator's content."""
        torch = _torch()
        gh = np.ascontiguousarray(gdev.detach().cpu().numpy())
        rh = np.ascontiguousarray(rdev.detach().cpu().numpy()) if rdev is not None else None
        key = (gh.tobytes(), rh.tobytes() if rh is not None else None, grav_xy, imp_dof)
        cache = self.__dict__.setdefault("_shared_op_cache", {})
        if key not in cache:
            if len(cache) >= 8:
                cache.clear()
            lib = _lib.load()
            par = np.ascontiguousarray(self._params_np[0])
            gx, gy = grav_xy if grav_xy is not None else (0.0, 0.0)
            args = (C.byref(self._plan), par.ctypes.data_as(C.c_void_p), self._etype.tobytes(), self._bc.tobytes(),
                    gh.ctypes.data_as(C.c_void_p), rh.ctypes.data_as(C.c_void_p) if rh is not None else None,
                    float(gx), float(gy), int(grav_xy is not None), int(imp_dof))
            cnt = int(lib.crb_shared_operator(*args, None))
            if cnt < 0:
                _lib.check(cnt)
            out = np.empty(cnt, dtype=np.float64)
            got = int(lib.crb_shared_operator(*args, out.ctypes.data_as(C.c_void_p)))
            if got != cnt:
                _lib.check(got if got < 0 else -1)
            cache[key] = torch.from_numpy(out).to(self.device)
        return cache[key]

    def _stream(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    def _as_state(self, x, name="State"):
        torch = _torch()
        if not isinstance(x, torch.Tensor):
            raise ValueError("State and input must be torch tensors on the CUDA device (no CPU path)")
        squeeze = x.ndim == 1
        if squeeze:
            x = x.unsqueeze(0)
        if x.ndim != 2 or x.shape[1] != 2 * self.n_free:
            raise ValueError(f"{name} must have shape [B, {2 * self.n_free}], got {tuple(x.shape)}")
        x = x.to(device=self.device, dtype=torch.float64).contiguous()
        return x, squeeze

    def _active_forces(self):
        """(drag, gravity, user plug-ins) currently enabled; polled on every call."""
        if self._forces_func is not None:
            return None, None, []
        fused, user = self.force_registry.split()
        drag = next((f for f in fused if f.fused_kind == "drag"), None)
        grav = next((f for f in fused if f.fused_kind == "gravity"), None)
        return drag, grav, user

    def _external_forces(self, x, user):
        """Additive f_ext from a replacing forces_func or from user plug-ins, at t = 0.0 (Q4)."""
        torch = _torch()
        f = None
        if self._forces_func is not None:
            f = self._forces_func(x, 0.0)
        else:
            for comp in user:
                c = comp.compute_forces(x, 0.0)
                f = c if f is None else f + c
        if f is None:
            return None
        if not isinstance(f, torch.Tensor):
            raise TypeError("force callables must return torch tensors on the CUDA device (no CPU fallback)")
        return f

    def _rhs(self, t: float, x, u=None, impulse=None, gain=None, ref=None, time_inputs=(), out=None):
        torch = _torch()
        B = x.shape[0]
        drag, grav, user = self._active_forces()
        f_ext = self._external_forces(x, user)
        sysm, keep = self.make_system(B, drag=drag, gravity=grav, u_const=u, impulse=impulse, gain=gain, ref=ref, f_ext=f_ext,
                                      time_inputs=time_inputs)
        if out is None:
            out = torch.empty_like(x)
        elif out.shape != x.shape or not out.is_contiguous() or out.dtype != torch.float64:
            raise ValueError("out must be a contiguous float64 tensor shaped like x")
        with torch.cuda.device(self.device):
            rc = _lib.load().crb_rhs(C.byref(self._plan), C.byref(sysm), x.data_ptr(), float(t), out.data_ptr(), self._stream())
        _lib.check(rc)
        return out

    def _builtin_forces(self, x, drag, gravity):
        """compute_forces(x, t) of a built-in component: crb_forces with only that component on."""
        torch = _torch()
        xs, squeeze = self._as_state(x)
        B = xs.shape[0]
        sysm, keep = self.make_system(B, drag=drag, gravity=gravity)
        out = torch.zeros((B, self.n_free), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().crb_forces(C.byref(self._plan), C.byref(sysm), xs.data_ptr(), out.data_ptr(), self._stream()))
        return out[0] if squeeze else out

    # -- reference API ------------------------------------------------------------------------------
    def create_system_func(self, forces_func: Callable = None) -> None:
        """dynamic_beam_model.py:243-274.  ``forces_func(x[B,2n], t) -> f[B,n]`` (torch) REPLACES
        the registry when given."""
        self._forces_func = forces_func

        def system(x):
            xs, squeeze = self._as_state(x)
            out = self._rhs(0.0, xs)
            return out[0] if squeeze else out

        self.system_func = system

    def create_input_func(self) -> None:
        """dynamic_beam_model.py:276-330: (x, u, t) -> [0 ; M^-1 u]."""
        torch = _torch()

        def input_function(x, u, t: float = 0.0):
            if not isinstance(x, torch.Tensor) or not isinstance(u, torch.Tensor):
                raise ValueError("State and input must be torch tensors")
            xs, squeeze = self._as_state(x)
            if u.ndim != x.ndim:
                raise ValueError("State and input must have matching batch dimensions")
            uu = u.unsqueeze(0) if squeeze else u
            n = self.n_free
            if uu.shape[-1] != n:
                raise ValueError(
                    f"Input vector length {uu.shape[-1]} must match position DOFs {n}. Expected {n}, got {uu.shape[-1]}"
                )
            B = xs.shape[0]
            # [0 ; M^-1 u] = rhs(x = 0, no forces, u) because k(0) = 0
            sysm, keep = self.make_system(B, u_const=uu)
            zero = torch.zeros_like(xs)
            out = torch.empty_like(xs)
            with torch.cuda.device(self.device):
                _lib.check(_lib.load().crb_rhs(C.byref(self._plan), C.byref(sysm), zero.data_ptr(), float(t), out.data_ptr(), self._stream()))
            return out[0] if squeeze else out

        self.input_func = input_function

    def get_system_func(self) -> Callable:
        if self.system_func is None:
            raise RuntimeError("System function not yet created")
        return self.system_func

    def get_dynamic_system(self) -> Callable:
        """dynamic_beam_model.py:332-364: f(t, x[B,2n], u) with u a tensor [B,n] / [n], a callable
        of t returning one, a TipImpulse / SinusoidInput / PiecewiseLinearInput (evaluated inside the kernel), a
        list of such parts (their sum), or None."""
        if self.system_func is None or self.input_func is None:
            raise RuntimeError("System and input functions must be created first")
        torch = _torch()

        def dynamic_system(t, x, u=None):
            xs, squeeze = self._as_state(x)
            uc, impulse, tv, other = split_input(u)
            force = uc
            if other is not None:
                val = other(t)
                if not isinstance(val, torch.Tensor):
                    raise ValueError("State and input must be torch tensors")
                force = val if force is None else force + val
            if force is not None:
                if force.shape[-1] != self.n_free:
                    raise ValueError(
                        f"Input vector length {force.shape[-1]} must match position DOFs {self.n_free}. "
                        f"Expected {self.n_free}, got {force.shape[-1]}"
                    )
                if force.ndim == 1:
                    force = force.unsqueeze(0)
            out = self._rhs(float(t), xs, u=force, impulse=impulse, time_inputs=tv)
            return out[0] if squeeze else out

        return dynamic_system
