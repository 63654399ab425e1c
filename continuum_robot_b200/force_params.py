"""ForceParams -- same fields and validation as the reference (models/force_params.py:6-69)."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class ForceParams:
    fluid_density: float = 0.0
    enable_fluid_effects: bool = False
    gravity_vector: List[float] = field(default_factory=lambda: [0.0, -9.81, 0.0])
    enable_gravity_effects: bool = False

    def __post_init__(self):
        self.gravity_vector = np.array(self.gravity_vector, dtype=float)
        if len(self.gravity_vector) != 3:
            raise ValueError("gravity_vector must have exactly 3 components [gx, gy, gz]")
        if np.allclose(self.gravity_vector, 0.0):  # a zero vector switches gravity off
            self.enable_gravity_effects = False
        if self.enable_fluid_effects and self.fluid_density <= 0:
            raise ValueError("fluid_density must be positive when fluid effects are enabled")

    def __bool__(self) -> bool:
        return bool(self.enable_fluid_effects or self.enable_gravity_effects)

    def get_gravity_vector(self) -> np.ndarray:
        return self.gravity_vector.copy()

    def set_gravity_vector(self, gravity_vector: List[float]) -> None:
        if len(gravity_vector) != 3:
            raise ValueError("gravity_vector must have exactly 3 components [gx, gy, gz]")
        self.gravity_vector = np.array(gravity_vector, dtype=float)
        if np.allclose(self.gravity_vector, 0.0):
            self.enable_gravity_effects = False
