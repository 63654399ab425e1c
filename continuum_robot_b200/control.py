"""Control layer of the hot path: state feedback inside the RHS, gain synthesis on the host.

  * ``FullStateLinear``           control/full_state_linear.py:5-64 -- u_c = K (r - x); the product
    ``K (r - x)`` is evaluated INSIDE the RHS kernel at every stage (crb_system_t.gain).
  * ``LinearQuadraticRegulator``  control/linear_quadratic_regulator.py:5-200 -- builds A, B and solves
    the Riccati equation ONCE on the host.  The reference calls python-control's ``ct.lqr``
    (absent in this image); the CARE solution is unique, so ``scipy.linalg.solve_continuous_are``
    gives the same gain to solver tolerance.  One-off host synthesis is outside the hot path
    (SURVEY section 2 row 7).
  * ``BatchedLinearQuadraticRegulator`` -- the same contract for an ENSEMBLE of designs, on the device
    (crb_lqr_gains): per-member A / B build and Riccati solve, one thread block
    per design; the gains feed the rollout as ``FullStateLinear(K[B,n,2n])``.
"""

from __future__ import annotations

import numpy as np

from .abstractions import AbstractInputHandler


class FullStateLinear(AbstractInputHandler):
    def __init__(self, gain_matrix, enabled: bool = True, reference=None):
        g = gain_matrix
        # 2-D: one gain for the ensemble (the reference's contract); 3-D [B,n,2n]: one gain per member
        # (BatchedLinearQuadraticRegulator)
        if getattr(g, "ndim", None) not in (2, 3):
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
        if K.shape[-1] != x.shape[-1]:
            raise ValueError("Gain matrix column dimension must match state vector length.")
        if K.ndim == 3:
            return torch.einsum("bij,bj->bi", K, (r - x).expand(K.shape[0], -1))
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


class BatchedLinearQuadraticRegulator:
    """``LinearQuadraticRegulator`` (control/linear_quadratic_regulator.py:5-200) for B designs at once.

    ``K_beam`` / ``M_beam``: CUDA FP64 tensors ``[B,n,n]`` or ``[n,n]`` (shared); ``Q [2n,2n]``, ``R [n,n]`` shared.
    ``compute_gain_matrix()`` returns ``K[B,n,2n]`` on the device and raises ``ValueError`` with the reference's
    messages if any member has no stabilising solution (:182-189).  There is no CPU path.
    """

    STATUS_TEXT = {1: "Mass matrix is singular and cannot be inverted",
                   2: "Failed to solve LQR problem: no stabilising Riccati solution",
                   3: "LQR solution results in unstable closed-loop system"}

    def __init__(self, K_beam, M_beam, Q, R, refine_passes: int = 1, decouple: bool = True):
        import torch

        for name, a in (("K_beam", K_beam), ("M_beam", M_beam), ("Q", Q), ("R", R)):
            if not isinstance(a, torch.Tensor) or not a.is_cuda:
                raise TypeError(f"{name} must be a CUDA torch tensor (no CPU path)")
            if a.ndim < 2 or a.shape[-1] != a.shape[-2]:
                raise ValueError(f"{name} must be a square matrix")
        if K_beam.shape[-1] != M_beam.shape[-1]:
            raise ValueError("K_beam and M_beam must have the same dimensions")
        n = int(M_beam.shape[-1])
        if Q.ndim != 2 or Q.shape[0] != 2 * n:
            raise ValueError(f"Q matrix dimension {Q.shape[0]} must match state dimension {2 * n}")
        if R.ndim != 2 or R.shape[0] != n:
            raise ValueError(f"R matrix dimension {R.shape[0]} must match input dimension {n}")
        f64 = lambda a: a.to(dtype=torch.float64).contiguous()
        self.K_beam = f64(K_beam if K_beam.ndim == 3 else K_beam[None])
        self.M_beam = f64(M_beam if M_beam.ndim == 3 else M_beam[None])
        self.Q, self.R = f64(Q), f64(R)
        self.n = n
        self.B = max(self.K_beam.shape[0], self.M_beam.shape[0])
        for name, a in (("K_beam", self.K_beam), ("M_beam", self.M_beam)):
            if a.shape[0] not in (1, self.B):
                raise ValueError(f"{name} has {a.shape[0]} members, expected 1 or {self.B}")
        self.refine_passes = int(refine_passes)
        self.decouple = bool(decouple)  # solve DOF groups that nothing couples as separate, smaller problems
        self._K = self._S = self.status = self.residual = None

    def _components(self):
        """Groups of position DOFs that M, K (of any member), R and the four n x n blocks of Q do not couple with each
        other.  The Riccati equation of such a design is reducible: S and the gain are block diagonal over the groups
        (a dense CARE solver returns exact zeros there, tests/golden/cfg5_samples.npz), and every group is a smaller
        synthesis problem.  A straight beam always splits into its axial and its bending DOFs with the diagonal weights
        of examples/lqr_control.py:61-66: 4n x 4n Hamiltonian -> one of 4n/3 and one of 8n/3, a third of the work."""
        import torch
        from scipy.sparse.csgraph import connected_components

        n = self.n
        P = (self.M_beam != 0).any(0) | (self.K_beam != 0).any(0) | (self.R != 0)
        for blk in (self.Q[:n, :n], self.Q[:n, n:], self.Q[n:, :n], self.Q[n:, n:]):
            P |= blk != 0
        ncomp, labels = connected_components((P | P.T).cpu().numpy(), directed=False)
        dev = self.M_beam.device
        return [torch.from_numpy((labels == c).nonzero()[0]).to(dev) for c in range(ncomp)]

    def _solve(self, M_beam, K_beam, Q, R):
        """crb_lqr_gains on one (sub-)problem: gain [B,k,2k], S [B,2k,2k], residual [B], status [B]."""
        import ctypes as C

        import torch

        from . import _lib

        lib = _lib.load()
        dev, n, B = M_beam.device, int(M_beam.shape[-1]), self.B
        need = C.c_size_t(0)
        _lib.check(lib.crb_lqr_workspace_bytes(n, B, C.byref(need)), ValueError)
        ws = torch.empty((need.value + 7) // 8, dtype=torch.float64, device=dev)
        gain = torch.empty((B, n, 2 * n), dtype=torch.float64, device=dev)
        S = torch.empty((B, 2 * n, 2 * n), dtype=torch.float64, device=dev)
        resid = torch.empty(B, dtype=torch.float64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        rc = lib.crb_lqr_gains(
            n, B, M_beam.data_ptr(), int(M_beam.shape[0] == 1), K_beam.data_ptr(), int(K_beam.shape[0] == 1),
            Q.data_ptr(), R.data_ptr(), self.refine_passes, gain.data_ptr(), S.data_ptr(), resid.data_ptr(),
            status.data_ptr(), ws.data_ptr(), need.value, torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, ValueError)
        return gain, S, resid, status

    def compute_gain_matrix(self):
        if self._K is not None:
            return self._K
        import torch

        dev, n, B = self.M_beam.device, self.n, self.B
        with torch.cuda.device(dev):
            comps = self._components() if self.decouple else []
            if len(comps) < 2:
                gain, S, resid, status = self._solve(self.M_beam, self.K_beam, self.Q, self.R)
            else:
                gain = torch.zeros((B, n, 2 * n), dtype=torch.float64, device=dev)
                S = torch.zeros((B, 2 * n, 2 * n), dtype=torch.float64, device=dev)
                status = torch.zeros(B, dtype=torch.int32, device=dev)
                res2 = torch.zeros(B, dtype=torch.float64, device=dev)
                for idx in comps:
                    idx2 = torch.cat([idx, idx + n])
                    sub = lambda a, i: a[:, i][:, :, i].contiguous()
                    Qc = self.Q[idx2][:, idx2].contiguous()
                    g, s, r, st = self._solve(sub(self.M_beam, idx), sub(self.K_beam, idx), Qc, self.R[idx][:, idx].contiguous())
                    gain[:, idx[:, None], idx2[None, :]] = g
                    S[:, idx2[:, None], idx2[None, :]] = s
                    status = torch.maximum(status, st)
                    res2 += (r * torch.linalg.matrix_norm(Qc)) ** 2  # residuals are relative to ||Q_c||_F
                resid = torch.sqrt(res2) / torch.linalg.matrix_norm(self.Q)
            bad = torch.nonzero(status).flatten()
        self.status, self.residual = status, resid
        if bad.numel():
            i = int(bad[0])
            raise ValueError(f"{self.STATUS_TEXT[int(status[i])]} (member {i}; {bad.numel()} of {B} members failed)")
        self._K, self._S = gain, S
        return gain

    def get_K(self):
        return self.compute_gain_matrix()

    def get_S(self):
        self.compute_gain_matrix()
        return self._S
