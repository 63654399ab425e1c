"""Synthetic ensemble definitions for the BASELINE.json configs (NumPy only, deterministic).

Used by ``bench.py``, the tests and ``tests/golden/make_golden.py`` so that the members the
reference ran in the build container are the same members the GPU runs on the B200 box.
Material set: examples/example_utilities.py:25-34 of the reference (Nitinol rod).
All RNG is ``numpy.random.default_rng(seed)`` (PCG64), draws in the documented order.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

NITINOL = {"length": 0.25, "E": 75e9, "r": 0.005, "rho": 6450.0, "drag_coef": 0.82}

LINEAR = 0
NONLINEAR = 1


def material():
    """Derived properties, examples/example_utilities.py:28-34."""
    p = dict(NITINOL)
    p["I"] = np.pi * p["r"] ** 4 / 4
    p["A"] = np.pi * p["r"] ** 2
    p["wetted_area"] = 2 * np.pi * p["r"] * p["length"]
    return p


@dataclass
class Ensemble:
    """Plain arrays describing an ensemble of cantilevers (node 0 FIXED)."""

    name: str
    n_elements: int
    elem_type: int  # all elements share a type in the BASELINE configs
    E: np.ndarray  # [B, N] elastic modulus per member and element
    q0: np.ndarray  # [B, n]
    v0: np.ndarray  # [B, n]
    h: float  # RK4 step (or 0 for adaptive)
    fluid_density: float = 0.0
    gravity: bool = False
    impulse_amp: Optional[np.ndarray] = None  # [B] tip-w force for t < impulse_duration
    impulse_duration: float = 0.01

    @property
    def n_members(self) -> int:
        return int(self.E.shape[0])

    @property
    def n_free(self) -> int:
        return 3 * self.n_elements


def config1() -> Ensemble:
    """cfg 1: linear cantilever, 10 elements, gravity, 0.1 N tip impulse (one member)."""
    N = 10
    return Ensemble(
        "cfg1_linear10_gravity", N, LINEAR, np.full((1, N), NITINOL["E"]),
        np.zeros((1, 3 * N)), np.zeros((1, 3 * N)), 2.5e-5,
        gravity=True, impulse_amp=np.array([0.1]),
    )


def config2() -> Ensemble:
    """cfg 2: nonlinear, 20 elements, fluid drag (rho_f = 1000), 0.1 N tip impulse."""
    N = 20
    return Ensemble(
        "cfg2_nonlinear20_fluid", N, NONLINEAR, np.full((1, N), NITINOL["E"]),
        np.zeros((1, 3 * N)), np.zeros((1, 3 * N)), 2.5e-5,
        fluid_density=1000.0, impulse_amp=np.array([0.1]),
    )


def config3(n_members: int = 65536, n_elements: int = 32, seed: int = 1234) -> Ensemble:
    """cfg 3: linear ensemble, per-member-per-element E, random IC, u = 0, RK4 h = 2e-5.

    Draw order: z_E [B,N], q0 [B,n], v0 [B,n] (all standard normal).
    """
    rng = np.random.default_rng(seed)
    B, N = n_members, n_elements
    E = NITINOL["E"] * np.exp(0.2 * rng.standard_normal((B, N)))
    q0 = 1e-3 * rng.standard_normal((B, 3 * N))
    v0 = 1e-1 * rng.standard_normal((B, 3 * N))
    return Ensemble("cfg3_linear32_ensemble", N, LINEAR, E, q0, v0, 2e-5)


def config4(n_members: int = 4096, n_elements: int = 64, seed: int = 4321) -> Ensemble:
    """cfg 4: nonlinear ensemble, drag + gravity, per-member scalar E and tip impulse; RK45.

    Draw order: z_E [B], amp [B] ~ U(0.05, 0.5).
    """
    rng = np.random.default_rng(seed)
    B, N = n_members, n_elements
    Em = NITINOL["E"] * np.exp(0.2 * rng.standard_normal(B))
    amp = rng.uniform(0.05, 0.5, B)
    return Ensemble(
        "cfg4_nonlinear64_rk45", N, NONLINEAR, np.repeat(Em[:, None], N, axis=1),
        np.zeros((B, 3 * N)), np.zeros((B, 3 * N)), 0.0,
        fluid_density=1000.0, gravity=True, impulse_amp=amp,
    )


def config5(n_members: int = 1048576, seed: int = 555) -> Ensemble:
    """cfg 5: LQR rollout, one shared N = 6 linear design with gravity; disturbance U(1, 20) N."""
    rng = np.random.default_rng(seed)
    B, N = n_members, 6
    amp = rng.uniform(1.0, 20.0, B)
    return Ensemble(
        "cfg5_lqr6_rollout", N, LINEAR, np.full((B, N), NITINOL["E"]),
        np.zeros((B, 3 * N)), np.zeros((B, 3 * N)), 5e-6,
        gravity=True, impulse_amp=amp,
    )


def sample_members(n_members: int, k: int, seed: int) -> np.ndarray:
    """Sorted sample of member indices used for CPU parity (rng.choice without replacement)."""
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(n_members, size=min(k, n_members), replace=False))
