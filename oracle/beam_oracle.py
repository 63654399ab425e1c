"""CPU oracle for the batched beam RHS + RK4 / RK45 path.  TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference's algorithm for the hot path.  It is
imported only by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- never by the product package
(``continuum_robot_b200``), which has no CPU fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *unmodified* reference
(imported from ``/root/reference/src`` in the build container) and stores its outputs in
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks every function below against those
files.  The adaptive integrator is SciPy's own ``solve_ivp(method="RK45")`` in the golden
files; ``rk45_solve`` below restates its controller (scipy/integrate/_ivp/rk.py, common.py,
SciPy 1.18.1) and is pinned against the same golden trajectories.

Each function cites the reference lines (relative to /root/reference/src/continuum_robot/)
whose behaviour it follows.  Structure intentionally mirrors the reference where that matters
for the CPU baseline (a Python loop over elements per RHS, an explicit inverse of M).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

LINEAR = 0
NONLINEAR = 1

BC_NONE = 0
BC_FIXED = 1
BC_PINNED = 2

# ----------------------------------------------------------------------------------------------
# Element level (models/segments.py)
# ----------------------------------------------------------------------------------------------

# Consistent-mass pattern, models/segments.py:64-78 (identical at :105-119).  Entry (i, j) is
# (integer, power of L); the element matrix is pattern * rho*A*L/420.
_MASS_INT = np.array(
    [
        [140, 0, 0, 70, 0, 0],
        [0, 156, -22, 0, 54, 13],
        [0, -22, 4, 0, -13, -3],
        [70, 0, 0, 140, 0, 0],
        [0, 54, -13, 0, 156, 22],
        [0, 13, -3, 0, 22, 4],
    ],
    dtype=np.float64,
)
_MASS_LPOW = np.array(
    [
        [0, 0, 0, 0, 0, 0],
        [0, 0, 1, 0, 0, 1],
        [0, 1, 2, 0, 1, 2],
        [0, 0, 0, 0, 0, 0],
        [0, 0, 1, 0, 0, 1],
        [0, 1, 2, 0, 1, 2],
    ]
)


def element_mass(L: float, rho: float, A: float) -> np.ndarray:
    """6x6 consistent mass, DOF order [u1,w1,phi1,u2,w2,phi2] (segments.py:64-78)."""
    return (_MASS_INT * np.power(float(L), _MASS_LPOW)) * (rho * A * L / 420)


def element_stiffness_linear(L: float, E: float, I: float, A: float) -> np.ndarray:
    """6x6 linear stiffness (segments.py:32-62)."""
    EI = E * I
    EA = E * A
    a = EA / L
    c1 = EI / L
    c2 = EI / L**2
    c3 = EI / L**3
    K = np.zeros((6, 6))
    K[0, 0] = a
    K[0, 3] = -a
    K[3, 0] = -a
    K[3, 3] = a
    K[1, :] = [0, 12 * c3, -6 * c2, 0, -12 * c3, -6 * c2]
    K[2, :] = [0, -6 * c2, 4 * c1, 0, 6 * c2, 2 * c1]
    K[4, :] = [0, -12 * c3, 6 * c2, 0, 12 * c3, 6 * c2]
    K[5, :] = [0, -6 * c2, 2 * c1, 0, 6 * c2, 4 * c1]
    return K


# Monomial tables for the reference's f3..f6 (segments.py:262-472).  Each row:
#   (coefficient literal, 'A'|'D', (p_u1, p_w1, p_t1, p_u2, p_w2, p_t2), power of L)
# The decimal literals are DATA taken from those lines (they are not exact fractions and parity
# depends on them).  f5 is the exact negation of f3 term by term (segments.py:378-424).
_F3_TERMS = [
    (0.0357142857143344, "A", (0, 0, 3, 0, 0, 0), 3),
    (-0.107142857143003, "A", (0, 0, 2, 0, 0, 1), 3),
    (1.28571428571433, "A", (0, 1, 2, 0, 0, 0), 2),
    (-1.28571428571433, "A", (0, 0, 2, 0, 1, 0), 2),
    (-0.107142857143003, "A", (0, 0, 1, 0, 0, 2), 3),
    (1.0, "A", (1, 0, 1, 0, 0, 0), 2),
    (-1.0, "A", (0, 0, 1, 1, 0, 0), 2),
    (-3.8571428571413, "A", (0, 2, 1, 0, 0, 0), 1),
    (7.7142857142826, "A", (0, 1, 1, 0, 1, 0), 1),
    (-3.8571428571413, "A", (0, 0, 1, 0, 2, 0), 1),
    (0.0357142857143344, "A", (0, 0, 0, 0, 0, 3), 3),
    (1.28571428571433, "A", (0, 1, 0, 0, 0, 2), 2),
    (-1.28571428571433, "A", (0, 0, 0, 0, 1, 2), 2),
    (1.0, "A", (1, 0, 0, 0, 0, 1), 2),
    (-1.0, "A", (0, 0, 0, 1, 0, 1), 2),
    (-3.857142857143, "A", (0, 2, 0, 0, 0, 1), 1),
    (7.71428571428601, "A", (0, 1, 0, 0, 1, 1), 1),
    (-3.857142857143, "A", (0, 0, 0, 0, 2, 1), 1),
    (-12.0, "A", (1, 1, 0, 0, 0, 0), 1),
    (12.0, "A", (1, 0, 0, 0, 1, 0), 1),
    (12.0, "A", (0, 1, 0, 1, 0, 0), 1),
    (-12.0, "A", (0, 0, 0, 1, 1, 0), 1),
    (10.2857142857147, "A", (0, 3, 0, 0, 0, 0), 0),
    (-30.857142857144, "A", (0, 2, 0, 0, 1, 0), 0),
    (30.857142857144, "A", (0, 1, 0, 0, 2, 0), 0),
    (-10.2857142857147, "A", (0, 0, 0, 0, 3, 0), 0),
    (-60.0, "D", (0, 0, 1, 0, 0, 0), 1),
    (-60.0, "D", (0, 0, 0, 0, 0, 1), 1),
    (120.0, "D", (0, 1, 0, 0, 0, 0), 0),
    (-120.0, "D", (0, 0, 0, 0, 1, 0), 0),
]
# f3 = 0.1 * sum / L**3 ; f5 = 0.1 * (-sum) / L**3

_F4_TERMS = [
    (0.0285714285714391, "A", (0, 0, 3, 0, 0, 0), 1),
    (-0.0107142857142861, "A", (0, 0, 2, 0, 0, 1), 1),
    (0.0107142857142719, "A", (0, 1, 2, 0, 0, 0), 0),
    (-0.0107142857142719, "A", (0, 0, 2, 0, 1, 0), 0),
    (0.00714285714286444, "A", (0, 0, 1, 0, 0, 2), 1),
    (-0.0214285714286007, "A", (0, 1, 1, 0, 0, 1), 0),
    (0.0214285714286007, "A", (0, 0, 1, 0, 1, 1), 0),
    (-0.133333333333333, "A", (1, 0, 1, 0, 0, 0), 0),
    (0.133333333333333, "A", (0, 0, 1, 1, 0, 0), 0),
    (0.128571428571433, "A", (0, 2, 1, 0, 0, 0), -1),
    (-0.257142857142867, "A", (0, 1, 1, 0, 1, 0), -1),
    (0.128571428571433, "A", (0, 0, 1, 0, 2, 0), -1),
    (-0.00357142857143344, "A", (0, 0, 0, 0, 0, 3), 1),
    (-0.0107142857142719, "A", (0, 1, 0, 0, 0, 2), 0),
    (0.0107142857142719, "A", (0, 0, 0, 0, 1, 2), 0),
    (0.0333333333333333, "A", (1, 0, 0, 0, 0, 1), 0),
    (-0.0333333333333333, "A", (0, 0, 0, 1, 0, 1), 0),
    (0.1, "A", (1, 1, 0, 0, 0, 0), -1),
    (-0.1, "A", (1, 0, 0, 0, 1, 0), -1),
    (-0.1, "A", (0, 1, 0, 1, 0, 0), -1),
    (0.1, "A", (0, 0, 0, 1, 1, 0), -1),
    (-0.128571428571377, "A", (0, 3, 0, 0, 0, 0), -2),
    (0.38571428571413, "A", (0, 2, 0, 0, 1, 0), -2),
    (-0.38571428571413, "A", (0, 1, 0, 0, 2, 0), -2),
    (0.128571428571377, "A", (0, 0, 0, 0, 3, 0), -2),
    (4.0, "D", (0, 0, 1, 0, 0, 0), -1),
    (2.0, "D", (0, 0, 0, 0, 0, 1), -1),
    (-6.0, "D", (0, 1, 0, 0, 0, 0), -2),
    (6.0, "D", (0, 0, 0, 0, 1, 0), -2),
]

_F6_TERMS = [
    (-0.00357142857143344, "A", (0, 0, 3, 0, 0, 0), 1),
    (0.00714285714286356, "A", (0, 0, 2, 0, 0, 1), 1),
    (-0.0107142857143003, "A", (0, 1, 2, 0, 0, 0), 0),
    (0.0107142857143003, "A", (0, 0, 2, 0, 1, 0), 0),
    (-0.0107142857142932, "A", (0, 0, 1, 0, 0, 2), 1),
    (-0.021428571428558, "A", (0, 1, 1, 0, 0, 1), 0),
    (0.021428571428558, "A", (0, 0, 1, 0, 1, 1), 0),
    (0.0333333333333333, "A", (1, 0, 1, 0, 0, 0), 0),
    (-0.0333333333333333, "A", (0, 0, 1, 1, 0, 0), 0),
    (0.0285714285714271, "A", (0, 0, 0, 0, 0, 3), 1),
    (0.0107142857142932, "A", (0, 1, 0, 0, 0, 2), 0),
    (-0.0107142857142932, "A", (0, 0, 0, 0, 1, 2), 0),
    (-0.133333333333333, "A", (1, 0, 0, 0, 0, 1), 0),
    (0.133333333333333, "A", (0, 0, 0, 1, 0, 1), 0),
    (0.128571428571428, "A", (0, 2, 0, 0, 0, 1), -1),
    (-0.257142857142856, "A", (0, 1, 0, 0, 1, 1), -1),
    (0.128571428571428, "A", (0, 0, 0, 0, 2, 1), -1),
    (0.1, "A", (1, 1, 0, 0, 0, 0), -1),
    (-0.1, "A", (1, 0, 0, 0, 1, 0), -1),
    (-0.1, "A", (0, 1, 0, 1, 0, 0), -1),
    (0.1, "A", (0, 0, 0, 1, 1, 0), -1),
    (-0.128571428571433, "A", (0, 3, 0, 0, 0, 0), -2),
    (0.3857142857143, "A", (0, 2, 0, 0, 1, 0), -2),
    (-0.3857142857143, "A", (0, 1, 0, 0, 2, 0), -2),
    (0.128571428571433, "A", (0, 0, 0, 0, 3, 0), -2),
    (2.0, "D", (0, 0, 1, 0, 0, 0), -1),
    (4.0, "D", (0, 0, 0, 0, 0, 1), -1),
    (-6.0, "D", (0, 1, 0, 0, 0, 0), -2),
    (6.0, "D", (0, 0, 0, 0, 1, 0), -2),
]

# Constants of the two nested axial expressions f1 / f2 (segments.py:178-205, :225-252).
_AX_C11 = 0.0666666666666665
_AX_C12 = 0.0166666666666667
_AX_C22 = 0.0666666666666667


def _eval_terms(terms, q6: np.ndarray, EA, EI, L):
    """Left-to-right sum of monomials, the way the reference's expression is evaluated.

    ``q6`` may carry leading batch axes: shape (..., 6).
    """
    acc = 0.0
    for coef, kind, pw, lp in terms:
        term = coef * (EA if kind == "A" else EI)
        for k in range(6):
            if pw[k]:
                term = term * q6[..., k] ** pw[k]
        if lp > 0:
            term = term * L**lp
        elif lp < 0:
            term = term / L ** (-lp)
        acc = acc + term
    return acc


def element_force_nonlinear(q6: np.ndarray, L, E, I, A) -> np.ndarray:
    """Nonlinear element force in nodal order [u1,w1,phi1,u2,w2,phi2].

    Follows segments.py:121-157 (ordering [f1,f3,f4,f2,f5,f6]) and the six expressions at
    :159-472, including the one-sided axial coupling of ``_f_1_expr`` (:197-205).
    """
    q6 = np.asarray(q6, dtype=np.float64)
    EA = E * A
    EI = E * I
    u1, w1, t1, u2, w2, t2 = (q6[..., k] for k in range(6))
    # shared brackets of f1/f2
    g1 = _AX_C11 * t1 * L - _AX_C12 * t2 * L - 0.05 * w1 + 0.05 * w2
    g2 = _AX_C12 * t1 * L - _AX_C22 * t2 * L + 0.05 * w1 - 0.05 * w2
    g3 = -0.05 * t1 * L - 0.05 * t2 * L + 0.6 * w1 - 0.6 * w2
    f1 = EA * (L * (-t1 * g1 + t2 * g2 + u1) + (-u2 - w1 + w2) * g3) / L**2
    f2 = EA * (L * (t1 * g1 - t2 * g2 - u1 + u2) + (w1 - w2) * g3) / L**2
    s3 = _eval_terms(_F3_TERMS, q6, EA, EI, L)
    f3 = 0.1 * s3 / L**3
    f5 = 0.1 * (-s3) / L**3
    f4 = _eval_terms(_F4_TERMS, q6, EA, EI, L)
    f6 = _eval_terms(_F6_TERMS, q6, EA, EI, L)
    return np.stack([f1, f3, f4, f2, f5, f6], axis=-1)


# ----------------------------------------------------------------------------------------------
# Beam level (models/euler_bernoulli_beam.py, models/dynamic_beam_model.py)
# ----------------------------------------------------------------------------------------------


@dataclass
class BeamSpec:
    """Numeric description of one beam (the parsed CSV / DataFrame of the reference).

    Columns follow dynamic_beam_model.py:78-93.  ``bc[i]`` applies to node i for i < N
    (row i of the CSV, :205-218); node N cannot be constrained through the CSV.
    """

    length: np.ndarray
    elastic_modulus: np.ndarray
    moment_inertia: np.ndarray
    density: np.ndarray
    cross_area: np.ndarray
    elem_type: np.ndarray  # LINEAR / NONLINEAR per element
    bc: np.ndarray  # BC_* per CSV row (node i)
    wetted_area: Optional[np.ndarray] = None
    drag_coef: Optional[np.ndarray] = None

    @property
    def n_elements(self) -> int:
        return int(len(self.length))

    @staticmethod
    def uniform(
        n_elements: int,
        *,
        length=0.25,
        E=75e9,
        r=0.005,
        rho=6450.0,
        drag_coef=0.82,
        elem_type=LINEAR,
        bc_first=BC_FIXED,
    ) -> "BeamSpec":
        """Nitinol rod of examples/example_utilities.py:25-34 replicated over N elements."""
        N = n_elements
        I = np.pi * r**4 / 4
        A = np.pi * r**2
        wet = 2 * np.pi * r * length
        bc = np.full(N, BC_NONE, dtype=np.int64)
        bc[0] = bc_first
        et = np.full(N, elem_type, dtype=np.int64) if np.isscalar(elem_type) else np.asarray(elem_type)
        return BeamSpec(
            length=np.full(N, length, dtype=np.float64),
            elastic_modulus=np.full(N, E, dtype=np.float64),
            moment_inertia=np.full(N, I, dtype=np.float64),
            density=np.full(N, rho, dtype=np.float64),
            cross_area=np.full(N, A, dtype=np.float64),
            elem_type=et.astype(np.int64),
            bc=bc,
            wetted_area=np.full(N, wet, dtype=np.float64),
            drag_coef=np.full(N, drag_coef, dtype=np.float64),
        )


@dataclass
class ForceSpec:
    """models/force_params.py:6-69."""

    fluid_density: float = 0.0
    enable_fluid_effects: bool = False
    gravity_vector: Sequence[float] = (0.0, -9.81, 0.0)
    enable_gravity_effects: bool = False

    def __post_init__(self):
        self.gravity_vector = np.array(self.gravity_vector, dtype=float)
        if np.allclose(self.gravity_vector, 0.0):  # force_params.py:29-31
            self.enable_gravity_effects = False


def constrained_dofs(spec: BeamSpec) -> List[int]:
    """FIXED removes (u,w,phi); PINNED removes (u,w) (euler_bernoulli_beam.py:240-254)."""
    out = []
    for node, bc in enumerate(spec.bc):
        if bc == BC_FIXED:
            out += [3 * node, 3 * node + 1, 3 * node + 2]
        elif bc == BC_PINNED:
            out += [3 * node, 3 * node + 1]
    return sorted(out)


def unconstrained_dofs(spec: BeamSpec) -> np.ndarray:
    c = set(constrained_dofs(spec))
    return np.array([d for d in range(3 * (spec.n_elements + 1)) if d not in c], dtype=np.int64)


def assemble_mass_full(spec: BeamSpec) -> np.ndarray:
    """Global (3(N+1))^2 mass; element e -> rows/cols 3e..3e+5 (euler_bernoulli_beam.py:139-161)."""
    N = spec.n_elements
    M = np.zeros((3 * (N + 1), 3 * (N + 1)))
    for e in range(N):
        M[3 * e : 3 * e + 6, 3 * e : 3 * e + 6] += element_mass(
            spec.length[e], spec.density[e], spec.cross_area[e]
        )
    return M


def assemble_stiffness_full(spec: BeamSpec) -> np.ndarray:
    """Dense K for all-linear beams (euler_bernoulli_beam.py:458-500)."""
    N = spec.n_elements
    if np.any(spec.elem_type != LINEAR):
        raise ValueError("Cannot extract stiffness matrix from beam with nonlinear segments")
    K = np.zeros((3 * (N + 1), 3 * (N + 1)))
    for e in range(N):
        K[3 * e : 3 * e + 6, 3 * e : 3 * e + 6] += element_stiffness_linear(
            spec.length[e], spec.elastic_modulus[e], spec.moment_inertia[e], spec.cross_area[e]
        )
    return K


class BeamOracle:
    """One beam: reduced M, M^-1, k(q), built-in forces and the state-space RHS.

    Mirrors DynamicEulerBernoulliBeam (dynamic_beam_model.py:16-364): ``rhs(t, x, u)`` is
    ``get_dynamic_system()(t, x, u)``.
    """

    def __init__(self, spec: BeamSpec, forces: Optional[ForceSpec] = None):
        self.spec = spec
        self.forces = forces or ForceSpec()
        self.N = spec.n_elements
        self.unc = unconstrained_dofs(spec)
        self.n = len(self.unc)
        M_full = assemble_mass_full(spec)
        self.M = M_full[np.ix_(self.unc, self.unc)]  # euler_bernoulli_beam.py:265
        self.M_inv = np.linalg.inv(self.M)  # dynamic_beam_model.py:60 (scipy.sparse inv there)
        self._Ke = [
            element_stiffness_linear(
                spec.length[e], spec.elastic_modulus[e], spec.moment_inertia[e], spec.cross_area[e]
            )
            if spec.elem_type[e] == LINEAR
            else None
            for e in range(self.N)
        ]
        # state map: reduced dof index -> (param, node)  (dynamic_beam_model.py:120-149)
        names = ("u", "w", "phi")
        self.dof_to_node_param = {i: (names[d % 3], d // 3) for i, d in enumerate(self.unc)}
        # drag: per node with both w and dw_dt, factor from CSV row min(node, N-1)
        # (fluid_forces.py:56-61, 83-90)
        self._drag_idx = []
        self._drag_fac = []
        if self.forces.enable_fluid_effects:
            wet = np.append(spec.wetted_area, spec.wetted_area[-1])
            cd = np.append(spec.drag_coef, spec.drag_coef[-1])
            for i, (p, node) in self.dof_to_node_param.items():
                if p == "w":
                    self._drag_idx.append(i)
                    self._drag_fac.append(0.5 * self.forces.fluid_density * cd[node] * wet[node])
        # gravity: per-segment masses (gravity_forces.py:54-64)
        self._seg_mass = spec.density * spec.cross_area * spec.length

    # -- stiffness ---------------------------------------------------------------------------
    def stiffness(self, q: np.ndarray) -> np.ndarray:
        """k(q) on reduced DOFs (euler_bernoulli_beam.py:163-219, 270-289)."""
        s = self.spec
        q_full = np.zeros(3 * (self.N + 1))
        q_full[self.unc] = q
        k_full = np.zeros_like(q_full)
        for e in range(self.N):
            qe = q_full[3 * e : 3 * e + 6]
            if self._Ke[e] is not None:
                fe = self._Ke[e] @ qe
            else:
                fe = element_force_nonlinear(
                    qe, s.length[e], s.elastic_modulus[e], s.moment_inertia[e], s.cross_area[e]
                )
            k_full[3 * e : 3 * e + 6] += fe
        return k_full[self.unc]

    # -- built-in forces ---------------------------------------------------------------------
    def drag(self, x: np.ndarray) -> np.ndarray:
        """fluid_forces.py:103-142."""
        f = np.zeros(self.n)
        for i, fac in zip(self._drag_idx, self._drag_fac):
            vel = x[self.n + i]
            f[i] = -fac * vel * abs(vel)
        return f

    def gravity(self, x: np.ndarray) -> np.ndarray:
        """gravity_forces.py:66-148, indices taken in the REDUCED vector."""
        n = self.n
        f = np.zeros(n)
        pos = x[:n]
        gx, gy = self.forces.gravity_vector[0], self.forces.gravity_vector[1]
        for i, m in enumerate(self._seg_mass):
            a = 3 * i + 2
            b = 3 * (i + 1) + 2
            if a < n and b < n:
                phi = 0.5 * (pos[a] + pos[b])
            elif a < n:
                phi = pos[a]
            elif b < n:
                phi = pos[b]
            else:
                phi = 0.0
            c, s_ = math.cos(phi), math.sin(phi)
            fa = (c * gx + s_ * gy) * m * 0.5
            ft = (-s_ * gx + c * gy) * m * 0.5
            for idx, val in ((3 * i, fa), (3 * i + 1, ft), (3 * (i + 1), fa), (3 * (i + 1) + 1, ft)):
                if idx < n:
                    f[idx] += val
        return f

    def builtin_forces(self, x: np.ndarray) -> np.ndarray:
        """Registry order: drag then gravity (dynamic_beam_model.py:220-241)."""
        f = np.zeros(self.n)
        if self.forces.enable_fluid_effects:
            f = f + self.drag(x)
        if self.forces.enable_gravity_effects:
            f = f + self.gravity(x)
        return f

    # -- state space -------------------------------------------------------------------------
    def system(self, x: np.ndarray, forces_func: Optional[Callable] = None) -> np.ndarray:
        """dynamic_beam_model.py:256-272; forces are evaluated at t = 0.0."""
        n = self.n
        k = self.stiffness(x[:n])
        f = self.builtin_forces(x) if forces_func is None else forces_func(x, 0.0)
        return np.concatenate([x[n:], -self.M_inv.dot(k) + self.M_inv.dot(f)])

    def input(self, u: np.ndarray) -> np.ndarray:
        """[0 ; M^-1 u] (dynamic_beam_model.py:294-328)."""
        return np.concatenate([np.zeros(self.n), self.M_inv.dot(u)])

    def rhs(self, t: float, x: np.ndarray, u) -> np.ndarray:
        """dynamic_beam_model.py:343-362."""
        force = u(t) if callable(u) else u
        return self.system(x) + self.input(force)

    def stiffness_matrix(self) -> np.ndarray:
        """BC-reduced dense K (euler_bernoulli_beam.py:502-511)."""
        return assemble_stiffness_full(self.spec)[np.ix_(self.unc, self.unc)]


# ----------------------------------------------------------------------------------------------
# Integrators
# ----------------------------------------------------------------------------------------------


def rk4_solve(f: Callable, x0: np.ndarray, t0: float, h: float, nsteps: int, save_every: int = 0):
    """Classical RK4 (north_star row R1): t_k = t0 + k*h, input evaluated at stage times."""
    x = np.array(x0, dtype=np.float64)
    out = []
    for k in range(nsteps):
        t = t0 + k * h
        k1 = f(t, x)
        k2 = f(t + 0.5 * h, x + (0.5 * h) * k1)
        k3 = f(t + 0.5 * h, x + (0.5 * h) * k2)
        k4 = f(t + h, x + h * k3)
        x = x + (h / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        if save_every and (k + 1) % save_every == 0:
            out.append(x.copy())
    return (x, np.array(out)) if save_every else x


def dense_stiffness(beam: "BeamOracle") -> np.ndarray:
    """BC-reduced K of an all-linear beam, column by column from k(e_j) (euler_bernoulli_beam.py:422-511)."""
    n = beam.n
    K = np.zeros((n, n))
    for j in range(n):
        e = np.zeros(n)
        e[j] = 1.0
        K[:, j] = beam.stiffness(e)
    return K


def midpoint_solve(beam: "BeamOracle", u: Callable, x0: np.ndarray, t0: float, h: float, nsteps: int, save_every: int = 0):
    """Implicit midpoint rule (= Newmark average acceleration) on M q'' = -K q + u(t) for an all-linear,
    force-free beam.  NOT a reference code path (the reference integrates with SciPy's solve_ivp): this
    restates the textbook rule on the reference's M and K so that crb_midpoint has a CPU checker.
        dv = h (M + h^2/4 K)^-1 (u(t + h/2) - K (q + h/2 v)),  v+ = v + dv,  q+ = q + h v + h/2 dv"""
    n = beam.n
    K = dense_stiffness(beam)
    A = beam.M + 0.25 * h * h * K
    x = np.array(x0, dtype=np.float64)
    out = []
    for k in range(nsteps):
        q, v = x[:n], x[n:]
        tm = t0 + (k + 0.5) * h
        dv = h * np.linalg.solve(A, u(tm) - K @ (q + 0.5 * h * v))
        x = np.concatenate([q + h * v + 0.5 * h * dv, v + dv])
        if save_every and (k + 1) % save_every == 0:
            out.append(x.copy())
    return (x, np.array(out)) if save_every else x


# Dormand-Prince 5(4) tableau as used by SciPy (scipy/integrate/_ivp/rk.py:538-565).
DP_C = np.array([0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1])
DP_A = np.array(
    [
        [0, 0, 0, 0, 0],
        [1 / 5, 0, 0, 0, 0],
        [3 / 40, 9 / 40, 0, 0, 0],
        [44 / 45, -56 / 15, 32 / 9, 0, 0],
        [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729, 0],
        [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    ]
)
DP_B = np.array([35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84])
DP_E = np.array([-71 / 57600, 0, 71 / 16695, -71 / 1920, 17253 / 339200, -22 / 525, 1 / 40])
DP_P = np.array(
    [
        [1, -8048581381 / 2820520608, 8663915743 / 2820520608, -12715105075 / 11282082432],
        [0, 0, 0, 0],
        [0, 131558114200 / 32700410799, -68118460800 / 10900136933, 87487479700 / 32700410799],
        [0, -1754552775 / 470086768, 14199869525 / 1410260304, -10690763975 / 1880347072],
        [0, 127303824393 / 49829197408, -318862633887 / 49829197408, 701980252875 / 199316789632],
        [0, -282668133 / 205662961, 2019193451 / 616988883, -1453857185 / 822651844],
        [0, 40617522 / 29380423, -110615467 / 29380423, 69997945 / 29380423],
    ]
)

SAFETY = 0.9
MIN_FACTOR = 0.2
MAX_FACTOR = 10.0


def _rms(v: np.ndarray) -> float:
    return float(np.linalg.norm(v) / math.sqrt(v.size))  # common.py:63-65


def select_initial_step(f, t0, y0, t_bound, f0, order, rtol, atol) -> float:
    """scipy/integrate/_ivp/common.py:68-135 (direction = +1, max_step = inf)."""
    if y0.size == 0:
        return math.inf
    interval_length = abs(t_bound - t0)
    if interval_length == 0.0:
        return 0.0
    scale = atol + np.abs(y0) * rtol
    d0 = _rms(y0 / scale)
    d1 = _rms(f0 / scale)
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    h0 = min(h0, interval_length)
    y1 = y0 + h0 * f0
    f1 = f(t0 + h0, y1)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1 <= 1e-15 and d2 <= 1e-15:
        h1 = max(1e-6, h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1 / (order + 1))
    return min(100 * h0, h1, interval_length)


@dataclass
class Rk45Result:
    t: np.ndarray
    y: np.ndarray  # (n_state, len(t)) like OdeResult.y
    nfev: int
    naccept: int
    nreject: int
    status: int
    h_last: float = 0.0


def rk45_solve(f, t_span, y0, t_eval=None, rtol=1e-3, atol=1e-6) -> Rk45Result:
    """Restatement of solve_ivp(method="RK45") for forward integration.

    Follows scipy/integrate/_ivp/rk.py:86-180 (step control), :14-72 (rk_step) and
    ivp.py (t_eval handling through the quartic dense output, rk.py:178-180, base.py).
    """
    t0, tf = float(t_span[0]), float(t_span[1])
    y = np.array(y0, dtype=np.float64)
    n = y.size
    nfev = 0

    def fun(t, yy):
        nonlocal nfev
        nfev += 1
        return f(t, yy)

    fcur = fun(t0, y)
    h_abs = select_initial_step(fun, t0, y, tf, fcur, 4, rtol, atol)
    err_exp = -1.0 / 5.0
    t = t0
    ts, ys = [], []
    if t_eval is None:
        ts.append(t0)
        ys.append(y.copy())
        t_eval_i = None
    else:
        t_eval = np.asarray(t_eval, dtype=np.float64)
        t_eval_i = 0
    naccept = nreject = 0
    status = 0
    K = np.zeros((7, n))
    while t < tf:
        min_step = 10 * abs(np.nextafter(t, np.inf) - t)
        if h_abs < min_step:
            h_abs = min_step
        step_accepted = False
        step_rejected = False
        while not step_accepted:
            if h_abs < min_step:
                status = -1
                break
            h = h_abs
            t_new = t + h
            if t_new - tf > 0:
                t_new = tf
            h = t_new - t
            h_abs = abs(h)
            K[0] = fcur
            for s in range(1, 6):
                dy = np.dot(K[:s].T, DP_A[s, :s]) * h
                K[s] = fun(t + DP_C[s] * h, y + dy)
            y_new = y + h * np.dot(K[:6].T, DP_B)
            f_new = fun(t + h, y_new)
            K[6] = f_new
            scale = atol + np.maximum(np.abs(y), np.abs(y_new)) * rtol
            err = _rms(np.dot(K.T, DP_E) * h / scale)
            if err < 1:
                if err == 0:
                    factor = MAX_FACTOR
                else:
                    factor = min(MAX_FACTOR, SAFETY * err**err_exp)
                if step_rejected:
                    factor = min(1, factor)
                h_abs *= factor
                step_accepted = True
                naccept += 1
            else:
                h_abs *= max(MIN_FACTOR, SAFETY * err**err_exp)
                step_rejected = True
                nreject += 1
        if status < 0:
            break
        t_old, y_old = t, y
        t, y, fcur = t_new, y_new, f_new
        if t_eval is None:
            ts.append(t)
            ys.append(y.copy())
        else:
            # ivp.py: t_eval points with t_old < te <= t via dense output
            Q = K.T.dot(DP_P)
            hh = t - t_old
            while t_eval_i < len(t_eval) and t_eval[t_eval_i] <= t:
                te = t_eval[t_eval_i]
                if te == t_old and naccept == 1:
                    pass
                xx = (te - t_old) / hh
                p = np.cumprod(np.tile(xx, 4))
                ys.append(y_old + hh * np.dot(Q, p))
                ts.append(te)
                t_eval_i += 1
    return Rk45Result(
        t=np.array(ts),
        y=np.array(ys).T if ys else np.zeros((n, 0)),
        nfev=nfev,
        naccept=naccept,
        nreject=nreject,
        status=status,
        h_last=h_abs,
    )


# ----------------------------------------------------------------------------------------------
# LQR (control/linear_quadratic_regulator.py, control/full_state_linear.py)
# ----------------------------------------------------------------------------------------------


def lqr_matrices(K_beam: np.ndarray, M_beam: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """A = [[0, I], [-M^-1 K, 0]], B = [[0], [M^-1]] (linear_quadratic_regulator.py:84-146)."""
    n = M_beam.shape[0]
    Minv = np.linalg.inv(M_beam)
    A = np.zeros((2 * n, 2 * n))
    A[:n, n:] = np.eye(n)
    A[n:, :n] = -Minv @ K_beam
    B = np.zeros((2 * n, n))
    B[n:, :] = Minv
    return A, B


def lqr_gain(K_beam, M_beam, Q, R) -> np.ndarray:
    """Gain of control.lqr (python-control, absent here): K = R^-1 B^T S with S from the CARE.

    The CARE solution is unique, so SciPy's solver agrees with python-control/slycot to solver
    tolerance; gain VALUES are "parity unpinned" by the reference's tests (tests/test_control.py
    pins only shape and closed-loop stability), so rollout parity fixes K in the golden file.
    """
    from scipy.linalg import solve_continuous_are

    A, B = lqr_matrices(K_beam, M_beam)
    S = solve_continuous_are(A, B, Q, R)
    return np.linalg.solve(R, B.T @ S)


def lqr_gain_refined(K_beam, M_beam, Q, R, steps: int = 3) -> np.ndarray:
    """`lqr_gain` polished by Newton-Kleinman steps (each one Lyapunov solve on the closed loop): the checker for
    the device solver crb_lqr_gains.  On the example beams SciPy's CARE gain is ~1.5e-8 (relative, max-norm) away
    from this fixed point, successive refinements agree to ~2e-10."""
    from scipy.linalg import solve_continuous_are, solve_continuous_lyapunov

    A, B = lqr_matrices(K_beam, M_beam)
    G = B @ np.linalg.solve(R, B.T)
    S = solve_continuous_are(A, B, Q, R)
    for _ in range(steps):
        res = A.T @ S + S @ A - S @ G @ S + Q
        S = S + solve_continuous_lyapunov((A - G @ S).T, -res)
        S = 0.5 * (S + S.T)
    return np.linalg.solve(R, B.T @ S)


def full_state_feedback(gain: np.ndarray, x: np.ndarray, r: np.ndarray) -> np.ndarray:
    """u_c = K (r - x)  (control/full_state_linear.py:58)."""
    return gain @ (r - x)
