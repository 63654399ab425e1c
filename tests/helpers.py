"""Shared helpers for the test-suite (oracle <-> product glue; test infrastructure only)."""

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BC_NAMES = {0: "NONE", 1: "FIXED", 2: "PINNED"}


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def case_names(g):
    return sorted({k.split("/")[0] for k in g.files})


def params_array(g, prefix=""):
    """[N,7] parameter block from golden arrays (values as parsed by the reference)."""
    cols = ("length", "elastic_modulus", "moment_inertia", "density", "cross_area", "wetted_area", "drag_coef")
    return np.stack([np.asarray(g[prefix + c], dtype=np.float64) for c in cols], axis=1)


def oracle_spec(g, prefix=""):
    from oracle import beam_oracle as bo

    return bo.BeamSpec(
        g[prefix + "length"], g[prefix + "elastic_modulus"], g[prefix + "moment_inertia"], g[prefix + "density"],
        g[prefix + "cross_area"], g[prefix + "elem_type"], g[prefix + "bc"], g[prefix + "wetted_area"], g[prefix + "drag_coef"],
    )


def block_err(got, ref, n):
    """SURVEY 8(d) parity metric: block inf-norm relative error of q and of v."""
    out = []
    for sl in (slice(0, n), slice(n, 2 * n)):
        den = np.abs(ref[..., sl]).max()
        num = np.abs(got[..., sl] - ref[..., sl]).max()
        out.append(num / den if den > 1e-300 else num)
    return max(out)


def make_gpu_beam(par, elem_type, bc, fluid_density=0.0, gravity=False, gravity_vector=(0.0, -9.81, 0.0), **kw):
    from continuum_robot_b200 import ForceParams
    from continuum_robot_b200.dynamic_beam import BatchedDynamicEulerBernoulliBeam

    fp = ForceParams(fluid_density=float(fluid_density), enable_fluid_effects=float(fluid_density) > 0,
                     gravity_vector=list(gravity_vector), enable_gravity_effects=bool(gravity))
    N = len(elem_type)
    beam = BatchedDynamicEulerBernoulliBeam(
        {"params": par, "type": [int(t) for t in elem_type], "boundary_condition": [int(b) for b in bc[:N]]},
        force_params=fp, **kw)
    beam.create_system_func()
    beam.create_input_func()
    return beam
