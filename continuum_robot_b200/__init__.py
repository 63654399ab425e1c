"""continuum_robot_b200 -- B200-native batched integrator for continuum-robot beam dynamics.

Drop-in (batched) mirror of the reference's beam-system and functional-composition API; all
numerics run in libcrb.so (hand-written sm_100a CUDA behind the C ABI of include/crb.h).
See DESIGN.md.
"""

from .abstractions import (  # noqa: F401
    AbstractForce,
    AbstractInputHandler,
    BoundaryConditionType,
    ElementType,
    Properties,
)
from .force_params import ForceParams  # noqa: F401
from .force_registry import ForceRegistry, InputRegistry  # noqa: F401


def __getattr__(name):
    # torch-dependent modules are imported lazily so that `import continuum_robot_b200.ensembles`
    # stays NumPy-only.
    if name in ("BatchedDynamicEulerBernoulliBeam", "BatchedEulerBernoulliBeam", "FluidDragForce", "GravityForce", "TipImpulse",
                "SinusoidInput", "PiecewiseLinearInput"):
        from . import dynamic_beam

        return getattr(dynamic_beam, name)
    if name in ("solve_ensemble", "rk4_steps", "midpoint_steps", "EnsembleResult", "HostPipeline"):
        from . import integrate

        return getattr(integrate, name)
    if name in ("FullStateLinear", "LinearQuadraticRegulator", "BatchedLinearQuadraticRegulator"):
        from . import control

        return getattr(control, name)
    if name in ("tip_displacement", "beam_shapes", "cantilever_frequencies", "output_selection"):
        from . import outputs

        return getattr(outputs, name)
    raise AttributeError(name)
