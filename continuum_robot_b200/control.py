"""Control layer of the hot path: state feedback inside the RHS, gain synthesis on the host.

  * ``FullStateLinear``           control/full_state_linear.py:5-64 -- u_c = K (r - x); the product
    ``K (r - x)`` is evaluated INSIDE the RHS kernel at every stage (crb_system_t.gain).
  * ``LinearQuadraticRegulator``  control/linear_quadratic_regulator.py:5-200 -- builds A, B and solves
    the Riccati equation ONCE on the host.  The reference calls python-control's ``ct.lqr``
    (absent in this image); the CARE solution is unique, so ``scipy.linalg.solve_continuous_are``
    gives the same gain to solver tolerance.  One-off host synthesis is outside the hot path
    (SURVEY section 2 row 7).
"""

from __future__ import annotations

import numpy as np

from .abstractions import AbstractInputHandler


class FullStateLinear(AbstractInputHandler):
    def __init__(self, gain_matrix, enabled: bool = True, reference=None):
        g = gain_matrix
        if getattr(g, "ndim", None) != 2:
            raise ValueError("Gain matrix must be a 2D array.")
        self.gain_matrix = g
        self.enabled = enabled
        self.reference = reference

    def compute_input(self, x, r, t):
        """K @ (r - x) for x[B,2n] / r[B,2n] or [2n] (torch, on device)."""
        import torch

        if not isinstance(x, torch.Tensor):
            raise TypeError("FullStateLinear.compute_input expects torch tensors (no CPU path)")
        K = self.gain_matrix
        if not isinstance(K, torch.Tensor):
            K = torch.as_tensor(np.asarray(K, dtype=np.float64), device=x.device)
        if x.shape[-1] != r.shape[-1]:
            raise ValueError("State vector and refrence vector must have the same length.")
        if K.shape[1] != x.shape[-1]:
            raise ValueError("Gain matrix column dimension must match state vector length.")
        return (r - x) @ K.T

    def is_enabled(self) -> bool:
        return self.enabled


class LinearQuadraticRegulator:
    """Same constructor contract and error messages as the reference class."""

    def __init__(self, K_beam, M_beam, Q, R):
        K_beam, M_beam, Q, R = (np.asarray(a, dtype=np.float64) for a in (K_beam, M_beam, Q, R))
        for name, a in (("K_beam", K_beam), ("M_beam", M_beam), ("Q", Q), ("R", R)):
            if a.ndim != 2 or a.shape[0] != a.shape[1]:
                raise ValueError(f"{name} must be a square matrix")
        if K_beam.shape != M_beam.shape:
            raise ValueError("K_beam and M_beam must have the same dimensions")
        self.K_beam, self.M_beam, self.Q, self.R = K_beam, M_beam, Q, R
        self._A = self._B = self._K = self._S = self._E = None

    def get_A(self) -> np.ndarray:
        if self._A is None:
            n = self.M_beam.shape[0]
            try:
                Minv = np.linalg.inv(self.M_beam)
            except np.linalg.LinAlgError:
                raise ValueError("Mass matrix is singular and cannot be inverted")
            A = np.zeros((2 * n, 2 * n))
            A[:n, n:] = np.eye(n)
            A[n:, :n] = -Minv @ self.K_beam
            self._A = A
        return self._A

    def get_B(self) -> np.ndarray:
        if self._B is None:
            n = self.M_beam.shape[0]
            try:
                Minv = np.linalg.inv(self.M_beam)
            except np.linalg.LinAlgError:
                raise ValueError("Mass matrix is singular and cannot be inverted")
            B = np.zeros((2 * n, n))
            B[n:, :] = Minv
            self._B = B
        return self._B

    def compute_gain_matrix(self) -> np.ndarray:
        if self._K is not None:
            return self._K
        from scipy.linalg import solve_continuous_are

        A, B = self.get_A(), self.get_B()
        if self.Q.shape[0] != A.shape[0]:
            raise ValueError(f"Q matrix dimension {self.Q.shape[0]} must match state dimension {A.shape[0]}")
        if self.R.shape[0] != B.shape[1]:
            raise ValueError(f"R matrix dimension {self.R.shape[0]} must match input dimension {B.shape[1]}")
        try:
            S = solve_continuous_are(A, B, self.Q, self.R)
            K = np.linalg.solve(self.R, B.T @ S)
        except Exception as e:  # same wrapping as the reference
            raise ValueError(f"Failed to solve LQR problem: {e}")
        eig = np.linalg.eigvals(A - B @ K)
        if np.any(eig.real >= 0):
            raise ValueError("LQR solution results in unstable closed-loop system")
        self._K, self._S, self._E = K, S, eig
        return self._K

    def get_K(self) -> np.ndarray:
        return self.compute_gain_matrix()
