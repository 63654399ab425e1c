"""API contract shared with the reference (models/abstractions.py:9-197), batched.

Same names and meaning as the reference so user code ports one-to-one; the only difference
is that state arrays are torch CUDA FP64 tensors with a leading member axis:
``x[B, 2n]`` instead of ``x[2n]``.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from enum import Enum
from typing import Optional


class ElementType(Enum):
    """models/abstractions.py:9-13."""

    LINEAR = "linear"
    NONLINEAR = "nonlinear"


class BoundaryConditionType(Enum):
    """models/abstractions.py:16-20.  FIXED removes (u, w, phi); PINNED removes (u, w)."""

    FIXED = "fixed"
    PINNED = "pinned"


_POSITIVE = ("length", "elastic_modulus", "moment_inertia", "density", "cross_area")
_LABEL = {
    "length": "Length",
    "elastic_modulus": "Elastic modulus",
    "moment_inertia": "Moment of inertia",
    "density": "Density",
    "cross_area": "Cross area",
}


@dataclass
class Properties:
    """Per-segment properties with the reference's validation (models/abstractions.py:23-67)."""

    length: float
    elastic_modulus: float
    moment_inertia: float
    density: float
    cross_area: float
    segment_id: int
    element_type: str
    wetted_area: Optional[float] = None
    drag_coef: Optional[float] = None

    def __post_init__(self):
        for name in _POSITIVE:
            val = getattr(self, name)
            if val <= 0:
                raise ValueError(f"{_LABEL[name]} must be positive, got {val}")
        if self.element_type.lower() not in {t.value for t in ElementType}:
            raise ValueError(f"Invalid element type: {self.element_type}")

    def get_element_type(self) -> ElementType:
        return ElementType(self.element_type.lower())

    def has_fluid_properties(self) -> bool:
        return self.wetted_area is not None and self.drag_coef is not None


class AbstractForce(ABC):
    """Force plug-in (models/abstractions.py:153-173).

    ``compute_forces(x, t)`` receives ``x[B, 2n]`` (torch, CUDA, float64) and returns
    ``f[B, n]`` on the same device.  A NumPy-only implementation is rejected with TypeError
    by the integrator: there is no CPU path.
    """

    @abstractmethod
    def compute_forces(self, x, t):
        ...

    @abstractmethod
    def is_enabled(self) -> bool:
        ...


class AbstractInputHandler(ABC):
    """Input plug-in (models/abstractions.py:176-197): returns a DELTA added to ``u``."""

    @abstractmethod
    def compute_input(self, x, r, t):
        ...

    @abstractmethod
    def is_enabled(self) -> bool:
        ...
