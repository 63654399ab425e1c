"""Functional-composition registries (reference: models/force_registry.py:6-173).

Semantics kept: ``register`` ignores disabled components, ``unregister`` returns a bool,
``get_registered_*`` return copies, ``is_enabled()`` is polled on every evaluation so that
toggling a component after ``create_system_func`` takes effect on the next call
(tests/test_advanced_composition.py:368-398 of the reference).

Difference: the built-in FluidDragForce / GravityForce are *fused* into the RHS kernel, so the
aggregated function only sums the components that are not fused (user plug-ins, evaluated as
torch ops on the device); ``split()`` tells the beam which built-ins are currently enabled.
"""

from __future__ import annotations

from typing import Callable, List, Tuple

from .abstractions import AbstractForce, AbstractInputHandler


class ForceRegistry:
    def __init__(self):
        self._forces: List[AbstractForce] = []

    def register(self, force_instance: AbstractForce) -> None:
        if force_instance.is_enabled():
            self._forces.append(force_instance)

    def unregister(self, force_instance: AbstractForce) -> bool:
        if force_instance in self._forces:
            self._forces.remove(force_instance)
            return True
        return False

    def clear(self) -> None:
        self._forces.clear()

    def get_registered_forces(self) -> List[AbstractForce]:
        return list(self._forces)

    def split(self) -> Tuple[List[AbstractForce], List[AbstractForce]]:
        """(enabled fused built-ins, enabled user plug-ins), registration order preserved."""
        fused, user = [], []
        for f in self._forces:
            if f.is_enabled():
                (fused if getattr(f, "fused_kind", None) else user).append(f)
        return fused, user

    def create_aggregated_function(self) -> Callable:
        """forces(x[B,2n], t) -> f[B,n]: sum of every enabled component (built-ins included)."""
        import torch

        def aggregate_forces(x, t: float = 0.0):
            total = None
            for force in self._forces:
                if force.is_enabled():
                    contrib = force.compute_forces(x, t)
                    total = contrib.clone() if total is None else total + contrib
            if total is None:
                return torch.zeros(x.shape[:-1] + (x.shape[-1] // 2,), dtype=x.dtype, device=x.device)
            return total

        return aggregate_forces

    def __len__(self) -> int:
        return len(self._forces)

    def __contains__(self, force_instance) -> bool:
        return force_instance in self._forces


class InputRegistry:
    def __init__(self):
        self._input_handlers: List[AbstractInputHandler] = []

    def register(self, input_handler: AbstractInputHandler) -> None:
        if input_handler.is_enabled():
            self._input_handlers.append(input_handler)

    def unregister(self, input_handler: AbstractInputHandler) -> bool:
        if input_handler in self._input_handlers:
            self._input_handlers.remove(input_handler)
            return True
        return False

    def clear(self) -> None:
        self._input_handlers.clear()

    def get_registered_handlers(self) -> List[AbstractInputHandler]:
        return list(self._input_handlers)

    def create_aggregated_function(self) -> Callable:
        """process_input(x, u, t) -> u + sum of handler deltas (force_registry.py:137-165)."""

        def aggregate_input_processing(x, u, t: float = 0.0):
            total = u.clone()
            for handler in self._input_handlers:
                if handler.is_enabled():
                    total = total + handler.compute_input(x, u, t)
            return total

        return aggregate_input_processing

    def __len__(self) -> int:
        return len(self._input_handlers)

    def __contains__(self, input_handler) -> bool:
        return input_handler in self._input_handlers
