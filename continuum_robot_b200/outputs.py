"""Trajectory consumers that run on the device: what the reference's callers
do with ``OdeResult.y`` right after the step loop.

  * ``tip_displacement``   examples/lqr_control.py:166-183 reads ``sol.y[n_pos - 2]`` (tip w)
  * ``beam_shapes``        examples/example_utilities.py:173-205 ``extract_beam_shapes``; like the
    reference it reads ``y[n_pos + 1 :: 3]`` -- the transverse VELOCITIES, not the displacements
    (SURVEY quirk Q7).  ``displacements=True`` gives the physically meant ``y[1 : n_pos : 3]``.
  * ``cantilever_frequencies``  the analytic known-answer of example_utilities.py:208-240, used as a
    physics check of the linear operator.
"""

from __future__ import annotations

import numpy as np


def output_selection(beam, what):
    """State indices a lean recording keeps (``solve_ensemble(..., outputs=...)`` / ``rk4_steps(..., out_sel=...)``):
    the kernels then write frames ``[T, B, len(selection)]`` instead of the whole state ``[T, B, 2n]``.

      * ``"tip"``             ``[n - 2]``: the tip trace ``sol.y[n_pos - 2]`` of examples/lqr_control.py:166-183
      * ``"shape"``           transverse displacement ``w`` of every node that has one, in node order
      * ``"shape_velocity"``  ``y[n_pos + 1 :: 3]``: what examples/example_utilities.py:173-205 actually reads (Q7)
      * a sequence of state indices (0 <= r < 2n, no repeats)
    """
    n = beam.n_free
    if isinstance(what, str):
        if what == "tip":
            return [n - 2]
        if what == "shape":
            nodes = sorted(nd for (prm, nd) in beam.node_param_to_state if prm == "w")
            return [beam.node_param_to_state[("w", nd)] for nd in nodes]
        if what == "shape_velocity":
            return list(range(n + 1, 2 * n, 3))
        raise ValueError(f"unknown output selection {what!r} (expected 'state', 'tip', 'shape', 'shape_velocity' or indices)")
    idx = [int(r) for r in what]
    if not idx or min(idx) < 0 or max(idx) >= 2 * n or len(set(idx)) != len(idx):
        raise ValueError(f"output selection must be distinct state indices in [0, {2 * n})")
    return idx


def tip_displacement(y):
    """y[B, 2n, T] (or [2n, T]) -> tip transverse displacement [B, T] (reduced DOF n-2)."""
    n = y.shape[-2] // 2
    return y[..., n - 2, :]


def beam_shapes(y, n_segments: int, dx: float, displacements: bool = False):
    """(x, yy) node coordinates over time, each [B, T, n_segments + 1], computed with torch ops on
    the device holding ``y``.  Row layout mirrors ``extract_beam_shapes`` (fixed base at 0)."""
    import torch

    squeeze = y.ndim == 2
    if squeeze:
        y = y.unsqueeze(0)
    B, n2, T = y.shape
    n = n2 // 2
    pos = y[:, 1:n:3, :] if displacements else y[:, n + 1 :: 3, :]  # [B, nodes, T]
    k = min(pos.shape[1], n_segments)
    yy = torch.zeros((B, T, n_segments + 1), dtype=y.dtype, device=y.device)
    yy[:, :, 1 : k + 1] = pos[:, :k, :].permute(0, 2, 1)
    x = (dx * torch.arange(n_segments + 1, dtype=y.dtype, device=y.device)).expand(B, T, n_segments + 1).clone()
    return (x[0], yy[0]) if squeeze else (x, yy)


def cantilever_frequencies(length, elastic_modulus, moment_inertia, density, cross_area):
    """First four natural frequencies (Hz) of a uniform cantilever (example_utilities.py:208-240)."""
    beta_l = np.array([0.596864, 1.49418, 2.50025, 3.49999]) * np.pi
    return beta_l**2 * np.sqrt(elastic_modulus * moment_inertia / (density * cross_area * length**4)) / (2 * np.pi)
