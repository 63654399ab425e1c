// crb_device.cuh -- device-side building blocks of the batched beam RHS (sm_100a).
//
// Work decomposition ("lane group"): one ensemble member is owned by G lanes of a warp; lane g
// owns M consecutive node slots (slot s = g*M + j) and the element LEFT of each of its slots.
// State, stage vectors and accumulators live in registers; neighbouring lanes exchange one
// node of halo by warp shuffles; the banded mass solve is a partitioned (SPIKE-type) block
// LDL^T whose constants are precomputed by crb_assemble (see crb_assemble.cuh).
//
// Reference behaviour restated here (paths under /root/reference/src/continuum_robot/):
//   element forces     models/segments.py:32-62 (linear), :121-472 (nonlinear polynomials)
//   scatter / BCs      models/euler_bernoulli_beam.py:163-219, 221-298
//   drag, gravity      models/fluid_forces.py:103-142, models/gravity_forces.py:66-148
//   RHS composition    models/dynamic_beam_model.py:256-272, 294-328, 343-362
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "crb.h"

#define CRB_FULL_MASK 0xffffffffu
#define CRB_SLOT_PAIRS 13  // 26 doubles of mass-solve constants per slot (as double2 pairs)
#define CRB_SCAN_PAIRS 6   // 12 doubles of scan constants per lane per level

// Scalars of crb_plan_t that kernels need (the red_index table travels as a device array).
struct KPlan {
  int N, n_free, n0, p_act, m, g, p, levels, contiguous, has_mask;
  int n_sm;  // SMs of the device (persistent kernels)
  long long mfac_doubles;
};

__device__ __forceinline__ double shfl_up_d(double v, int d, int w) {
  return __shfl_up_sync(CRB_FULL_MASK, v, d, w);
}
__device__ __forceinline__ double shfl_down_d(double v, int d, int w) {
  return __shfl_down_sync(CRB_FULL_MASK, v, d, w);
}

// ------------------------------------------------------------------------------------------
// Mass-solve constants: layout [pair][j][g] of double2 so the G lanes of a member read one
// contiguous 16*G-byte run (conflict-free LDS.128 when staged in shared memory).
//   pairs 0,1 Lm | 2,3 Sinv | 4,5 U | 6,7 Phi | 8,9 Psi | 10 (lu, sinv_u) | 11 (uu, phi_u) | 12 (psi_u, -)
// ------------------------------------------------------------------------------------------
template <int M>
struct MassConsts {
  const double* slot;  // base of the slot part
  const double* scan;  // base of the scan part
  int G, g;
  __device__ __forceinline__ double2 ld(int pair, int j) const {
    return *reinterpret_cast<const double2*>(slot + ((((pair * M) + j) * G + g) << 1));
  }
  __device__ __forceinline__ double2 lds(int level, int pair) const {
    return *reinterpret_cast<const double2*>(scan + (((level * CRB_SCAN_PAIRS + pair) * G + g) << 1));
  }
};

__device__ __forceinline__ void mv2(const double2 r0, const double2 r1, double x0, double x1,
                                    double& y0, double& y1) {
  // y += [r0; r1] * x
  y0 = fma(r0.x, x0, fma(r0.y, x1, y0));
  y1 = fma(r1.x, x0, fma(r1.y, x1, y1));
}

// Solve M a = b in place (b[j][d], d = u,w,phi).  Algebraically exact partitioned solve:
//   local forward  y~_s = b_s - Lm_s y~_{s-1}          (incoming boundary value ignored)
//   forward scan   z_g  = y~_last(g) + Tl_g z_{g-1}     (Kogge-Stone with precomputed products)
//   local backward x~_s = Sinv_s y~_s - U_s x~_{s+1}
//   backward scan  w_g  = x~_first(g) + Phi_first y_in + Psi_first w_{g+1}
//   correction     x_s  = x~_s + Phi_s y_in(g) + Psi_s x_in(g)
template <int M>
__device__ __forceinline__ void mass_solve(double (&b)[M][3], const MassConsts<M>& C, int levels) {
  const int G = C.G;
  // ---- local forward ----
#pragma unroll
  for (int j = 1; j < M; ++j) {
    const double2 l0 = C.ld(0, j), l1 = C.ld(1, j);
    const double lu = C.ld(10, j).x;
    b[j][0] = fma(-lu, b[j - 1][0], b[j][0]);
    const double p0 = b[j - 1][1], p1 = b[j - 1][2];
    b[j][1] = fma(-l0.x, p0, fma(-l0.y, p1, b[j][1]));
    b[j][2] = fma(-l1.x, p0, fma(-l1.y, p1, b[j][2]));
  }
  double yin0 = 0.0, yin1 = 0.0, yin2 = 0.0;
  if (G > 1) {
    double z0 = b[M - 1][0], z1 = b[M - 1][1], z2 = b[M - 1][2];
    for (int l = 0; l < levels; ++l) {
      const int d = 1 << l;
      const double s0 = shfl_up_d(z0, d, G), s1 = shfl_up_d(z1, d, G), s2 = shfl_up_d(z2, d, G);
      const double2 c0 = C.lds(l, 0), c1 = C.lds(l, 1), cu = C.lds(l, 2);
      z0 = fma(cu.x, s0, z0);
      mv2(c0, c1, s1, s2, z1, z2);
    }
    yin0 = shfl_up_d(z0, 1, G);
    yin1 = shfl_up_d(z1, 1, G);
    yin2 = shfl_up_d(z2, 1, G);
    if (C.g == 0) { yin0 = 0.0; yin1 = 0.0; yin2 = 0.0; }
  }
  // ---- local backward on the uncorrected y~ ----
  {
    const double2 s0 = C.ld(2, M - 1), s1 = C.ld(3, M - 1);
    const double su = C.ld(10, M - 1).y;
    const double y1 = b[M - 1][1], y2 = b[M - 1][2];
    b[M - 1][0] *= su;
    b[M - 1][1] = fma(s0.x, y1, s0.y * y2);
    b[M - 1][2] = fma(s1.x, y1, s1.y * y2);
  }
#pragma unroll
  for (int j = M - 2; j >= 0; --j) {
    const double2 s0 = C.ld(2, j), s1 = C.ld(3, j), u0 = C.ld(4, j), u1 = C.ld(5, j);
    const double2 cu = C.ld(10, j);
    const double uu = C.ld(11, j).x;
    const double y1 = b[j][1], y2 = b[j][2];
    const double n1 = b[j + 1][1], n2 = b[j + 1][2];
    b[j][0] = fma(cu.y, b[j][0], -uu * b[j + 1][0]);
    b[j][1] = fma(s0.x, y1, fma(s0.y, y2, -fma(u0.x, n1, u0.y * n2)));
    b[j][2] = fma(s1.x, y1, fma(s1.y, y2, -fma(u1.x, n1, u1.y * n2)));
  }
  if (G > 1) {
    // ---- backward scan on the first slot of every chunk ----
    double w0, w1, w2;
    {
      const double2 f0 = C.ld(6, 0), f1 = C.ld(7, 0);
      const double fu = C.ld(11, 0).y;
      w0 = fma(fu, yin0, b[0][0]);
      w1 = b[0][1];
      w2 = b[0][2];
      mv2(f0, f1, yin1, yin2, w1, w2);
    }
    for (int l = 0; l < levels; ++l) {
      const int d = 1 << l;
      const double s0 = shfl_down_d(w0, d, G), s1 = shfl_down_d(w1, d, G), s2 = shfl_down_d(w2, d, G);
      const double2 c0 = C.lds(l, 3), c1 = C.lds(l, 4), cu = C.lds(l, 5);
      w0 = fma(cu.x, s0, w0);
      mv2(c0, c1, s1, s2, w1, w2);
    }
    double xin0 = shfl_down_d(w0, 1, G), xin1 = shfl_down_d(w1, 1, G), xin2 = shfl_down_d(w2, 1, G);
    if (C.g == G - 1) { xin0 = 0.0; xin1 = 0.0; xin2 = 0.0; }
    // ---- correction ----
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double2 f0 = C.ld(6, j), f1 = C.ld(7, j), p0 = C.ld(8, j), p1 = C.ld(9, j);
      const double fu = C.ld(11, j).y, pu = C.ld(12, j).x;
      b[j][0] = fma(fu, yin0, fma(pu, xin0, b[j][0]));
      mv2(f0, f1, yin1, yin2, b[j][1], b[j][2]);
      mv2(p0, p1, xin1, xin2, b[j][1], b[j][2]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Compact mass solve (fast kernels of crb_rk4_fast.cuh and the shape-specialised linear general
// kernels).  Same block-LDL^T factors as mass_solve, but the partition (SPIKE) corrections are
// RECOMPUTED by a second local sweep instead of being read as stored spike vectors: per slot only
// Sinv (4 doubles) and T = -Lm (5 doubles, the unit-lower block coupling a slot to its left
// neighbour) come from shared memory -- 4.5 LDS.128 per slot and solve instead of 13 -- and every
// sweep is a chain of pure DFMAs (5 per slot):
//
//   forward  A: y~_j = b_j + T_j y~_{j-1}  (zero incoming)  -> Kogge-Stone scan -> y_last of every lane
//   forward  B: y_j  = b_j + T_j y_{j-1}   (true incoming y from the left lane)
//   diagonal  : xhat_j = Sinv_j y_j
//   backward A: x~_j = xhat_j + T_{j+1}^T x~_{j+1}  (zero incoming) -> scan -> x_first of every lane
//   backward B: x_j  = xhat_j + T_{j+1}^T x_{j+1}   (the right lane sends T_0^T x_first)
//
// The factors are the true ones, so non-uniform mass, phantom slots (T = 0) and identity rows of
// constrained DOFs are handled like in mass_solve.  "Compact copy" layout appended by crb_assemble
// to every factor set:  slot part [pair 0..3][j][g] double2 = (s00,s01) (s11,sinv_u) (t00,t01) (t10,t11),
// then [jj < ceil(M/2)][g] double2 = (tu_{2jj}, tu_{2jj+1});  scan part [level][pair 0..4][g] double2.
// ------------------------------------------------------------------------------------------
__host__ __device__ constexpr int crb_compact_slot_doubles(int m, int g) { return 2 * g * (4 * m + (m + 1) / 2); }
__host__ __device__ constexpr int crb_compact_doubles(int m, int g, int levels) {
  return crb_compact_slot_doubles(m, g) + 10 * (levels > 0 ? levels : 1) * g;
}

// what the compact solve needs from a lane
struct FastMass {
  static constexpr int kPin = 0;   // slots whose 2x2 T block lives in registers for the whole launch
  static constexpr int kPinU = 0;  // 1: the tu values too
  static constexpr int kPinS = 0;  // slots whose Sinv lives in registers
  int g;                // lane within the member
  const double* fslot;  // slot part of the compact copy (shared memory)
  const double* fscan;  // scan part
  double pin[1][4], pinu[1], pins[1][4];  // unused (nothing pinned)
};

template <int M, int G, typename CT>
__device__ __forceinline__ double2 ld_fslot(const CT& C, int pair, int j) {
  return *reinterpret_cast<const double2*>(C.fslot + (((pair * M + j) * G + C.g) << 1));
}
template <int M, int G, typename CT>
__device__ __forceinline__ double2 ld_fscan(const CT& C, int level, int pair) {
  return *reinterpret_cast<const double2*>(C.fscan + (((level * 5 + pair) * G + C.g) << 1));
}

// Mass solve of R right-hand sides at once (they share every constant read from shared memory).
template <int M, int LV, int R, typename CT>
__device__ __forceinline__ void fast_solve_r(double (&b)[R][M][3], const CT& C) {
  constexpr int G = 1 << LV;
  // T_j = -Lm of the lane's own slots stays in registers for the four sweeps
  double t00[M], t01[M], t10[M], t11[M], tu[M + 1];
#pragma unroll
  for (int jj = 0; jj < (M + 1) / 2; ++jj) {
    if (CT::kPinU) {
      tu[2 * jj] = C.pinu[CT::kPinU ? 2 * jj : 0];
      tu[2 * jj + 1] = C.pinu[CT::kPinU ? 2 * jj + 1 : 0];
    } else {
      const double2 u = ld_fslot<M, G, CT>(C, 4, jj);  // address ((4 M + jj) G + g): the tu block follows pair 3
      tu[2 * jj] = u.x;
      tu[2 * jj + 1] = u.y;
    }
  }
#pragma unroll
  for (int j = 0; j < M; ++j) {
    if (j < CT::kPin) {  // pinned in registers by the kernel (saves 2 LDS.128 per slot and solve)
      t00[j] = C.pin[j < CT::kPin ? j : 0][0];
      t01[j] = C.pin[j < CT::kPin ? j : 0][1];
      t10[j] = C.pin[j < CT::kPin ? j : 0][2];
      t11[j] = C.pin[j < CT::kPin ? j : 0][3];
    } else {
      const double2 a = ld_fslot<M, G, CT>(C, 2, j), c = ld_fslot<M, G, CT>(C, 3, j);
      t00[j] = a.x;
      t01[j] = a.y;
      t10[j] = c.x;
      t11[j] = c.y;
    }
  }
  double y0[R], y1[R], y2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { y0[r] = b[r][0][0]; y1[r] = b[r][0][1]; y2[r] = b[r][0][2]; }
  // ---- forward A: y~ at the chunk's last slot, zero incoming ----
#pragma unroll
  for (int j = 1; j < M; ++j)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double p1 = y1[r], p2 = y2[r];
      y0[r] = fma(tu[j], y0[r], b[r][j][0]);
      y1[r] = fma(t00[j], p1, fma(t01[j], p2, b[r][j][1]));
      y2[r] = fma(t10[j], p1, fma(t11[j], p2, b[r][j][2]));
    }
  double xi0[R], xi1[R], xi2[R];  // y of the left neighbour's last slot
#pragma unroll
  for (int r = 0; r < R; ++r) { xi0[r] = 0.0; xi1[r] = 0.0; xi2[r] = 0.0; }
  double cub[LV > 0 ? LV : 1];  // u-scan coefficients of the backward scan (loaded with the forward ones)
  if (G > 1) {
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      const int d = 1 << l;
      const double2 c0 = ld_fscan<M, G, CT>(C, l, 0), c1 = ld_fscan<M, G, CT>(C, l, 1), cu = ld_fscan<M, G, CT>(C, l, 4);
      cub[l] = cu.y;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double s0 = shfl_up_d(y0[r], d, G), s1 = shfl_up_d(y1[r], d, G), s2 = shfl_up_d(y2[r], d, G);
        y0[r] = fma(cu.x, s0, y0[r]);
        mv2(c0, c1, s1, s2, y1[r], y2[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      xi0[r] = shfl_up_d(y0[r], 1, G);
      xi1[r] = shfl_up_d(y1[r], 1, G);
      xi2[r] = shfl_up_d(y2[r], 1, G);
      if (C.g == 0) { xi0[r] = 0.0; xi1[r] = 0.0; xi2[r] = 0.0; }
    }
  }
  // ---- forward B (true incoming), then the block-diagonal solve: b[j] <- xhat_j = Sinv_j y_j ----
#pragma unroll
  for (int j = 0; j < M; ++j) {
    double2 sa, sc;
    if (j < CT::kPinS) {
      sa = make_double2(C.pins[j < CT::kPinS ? j : 0][0], C.pins[j < CT::kPinS ? j : 0][1]);
      sc = make_double2(C.pins[j < CT::kPinS ? j : 0][2], C.pins[j < CT::kPinS ? j : 0][3]);
    } else {
      sa = ld_fslot<M, G, CT>(C, 0, j);
      sc = ld_fslot<M, G, CT>(C, 1, j);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double p1 = xi1[r], p2 = xi2[r];
      xi0[r] = fma(tu[j], xi0[r], b[r][j][0]);
      xi1[r] = fma(t00[j], p1, fma(t01[j], p2, b[r][j][1]));
      xi2[r] = fma(t10[j], p1, fma(t11[j], p2, b[r][j][2]));
      b[r][j][0] = sc.y * xi0[r];
      b[r][j][1] = fma(sa.x, xi1[r], sa.y * xi2[r]);
      b[r][j][2] = fma(sa.y, xi1[r], sc.x * xi2[r]);
    }
  }
  // ---- backward A: x~ at the chunk's first slot, zero incoming ----
  double r0[R], r1[R], r2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { r0[r] = b[r][M - 1][0]; r1[r] = b[r][M - 1][1]; r2[r] = b[r][M - 1][2]; }
#pragma unroll
  for (int j = M - 2; j >= 0; --j)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double p1 = r1[r], p2 = r2[r];
      r0[r] = fma(tu[j + 1], r0[r], b[r][j][0]);
      r1[r] = fma(t00[j + 1], p1, fma(t10[j + 1], p2, b[r][j][1]));
      r2[r] = fma(t01[j + 1], p1, fma(t11[j + 1], p2, b[r][j][2]));
    }
  double n0[R], n1[R], n2[R];  // T_0^T x_first of the right neighbour
#pragma unroll
  for (int r = 0; r < R; ++r) { n0[r] = 0.0; n1[r] = 0.0; n2[r] = 0.0; }
  if (G > 1) {
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      const int d = 1 << l;
      const double2 c0 = ld_fscan<M, G, CT>(C, l, 2), c1 = ld_fscan<M, G, CT>(C, l, 3);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const double s0 = shfl_down_d(r0[r], d, G), s1 = shfl_down_d(r1[r], d, G), s2 = shfl_down_d(r2[r], d, G);
        r0[r] = fma(cub[l], s0, r0[r]);
        mv2(c0, c1, s1, s2, r1[r], r2[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double z0 = tu[0] * r0[r];
      const double z1 = fma(t00[0], r1[r], t10[0] * r2[r]);
      const double z2 = fma(t01[0], r1[r], t11[0] * r2[r]);
      n0[r] = shfl_down_d(z0, 1, G);
      n1[r] = shfl_down_d(z1, 1, G);
      n2[r] = shfl_down_d(z2, 1, G);
      if (C.g == G - 1) { n0[r] = 0.0; n1[r] = 0.0; n2[r] = 0.0; }
    }
  }
  // ---- backward B: true incoming; b[j] <- x_j ----
#pragma unroll
  for (int r = 0; r < R; ++r) {
    n0[r] += b[r][M - 1][0];
    n1[r] += b[r][M - 1][1];
    n2[r] += b[r][M - 1][2];
    b[r][M - 1][0] = n0[r];
    b[r][M - 1][1] = n1[r];
    b[r][M - 1][2] = n2[r];
  }
#pragma unroll
  for (int j = M - 2; j >= 0; --j)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const double p1 = n1[r], p2 = n2[r];
      n0[r] = fma(tu[j + 1], n0[r], b[r][j][0]);
      n1[r] = fma(t00[j + 1], p1, fma(t10[j + 1], p2, b[r][j][1]));
      n2[r] = fma(t01[j + 1], p1, fma(t11[j + 1], p2, b[r][j][2]));
      b[r][j][0] = n0[r];
      b[r][j][1] = n1[r];
      b[r][j][2] = n2[r];
    }
}

// ------------------------------------------------------------------------------------------
// Element forces.  `c` = 4 coefficients written by crb_assemble for the element's type.
// Both accumulate  bA -= f[0:3]  (node 1) and  bB -= f[3:6]  (node 2).
// ------------------------------------------------------------------------------------------

// linear: c = (EA/L, 12EI/L^3, 6EI/L^2, 2EI/L)   (models/segments.py:32-62)
__device__ __forceinline__ void elem_linear(const double4 c, const double (&qa)[3], const double (&qb)[3],
                                            double (&bA)[3], double (&bB)[3]) {
  const double du = qa[0] - qb[0];
  const double d = qa[1] - qb[1];
  const double s = qa[2] + qb[2];
  const double fu = c.x * du;                 // axial force on node 1 (node 2: -fu)
  const double V = fma(c.y, d, -c.z * s);     // shear on node 1 (node 2: -V)
  const double R = fma(c.w, s, -c.z * d);     // common part of both end moments
  bA[0] -= fu;
  bB[0] += fu;
  bA[1] -= V;
  bB[1] += V;
  bA[2] -= fma(c.w, qa[2], R);
  bB[2] -= fma(c.w, qb[2], R);
}

// Decimal literals of the reference's nonlinear polynomials (models/segments.py:159-472), kept in
// constant memory so that every DFMA takes its coefficient as a constant-bank operand (as 64-bit
// immediates they cost two uniform moves each, 18 % of the issued instructions of the kernel).
static __constant__ double kNlc[36] = {
    /* 0 g1 */ 0.0666666666666665, 0.0166666666666667, 0.05,
    /* 3 g2 */ 0.0666666666666667, 0.6,
    /* 5 f3 */ 0.0357142857143344, 0.107142857143003, 1.28571428571433, 3.8571428571413, 3.857142857143,
    /*10    */ 10.2857142857147, 0.1,
    /*12 f4 */ 0.0285714285714391, 0.0107142857142861, 0.0107142857142719, 0.00714285714286444,
    /*16    */ 0.0214285714286007, 0.133333333333333, 0.0333333333333333, 0.128571428571433,
    /*20    */ 0.00357142857143344, 0.128571428571377,
    /*22 f6 */ 0.00714285714286356, 0.0107142857143003, 0.0107142857142932, 0.021428571428558,
    /*26    */ 0.0285714285714271, 0.128571428571428,
    /*28    */ 0.0357142857143344 + 0.107142857143003, 0, 0, 0, 0, 0, 0, 0};

// nonlinear: c = (EA/L^2, EI/L^2, L, 1/L).  The six polynomials of models/segments.py:159-472
// rewritten in the scaled variables a = theta1*L, b = theta2*L, d = w1-w2, e = u1-u2 (the
// monomials group exactly; decimal literals are the reference's).  Node order
// [f1,f3,f4 | f2,f5,f6] (segments.py:146-155); f5 = -f3; f1 keeps the one-sided axial term of
// segments.py:197-205.
__device__ __forceinline__ void elem_nonlinear(const double4 c, const double (&qa)[3], const double (&qb)[3],
                                               double (&bA)[3], double (&bB)[3]) {
  const double* K = kNlc;
  const double al = c.x, de = c.y, L = c.z, iL = c.w;
  const double u1 = qa[0], u2 = qb[0];
  const double a = qa[2] * L, b = qb[2] * L;
  const double d = qa[1] - qb[1];
  const double e = u1 - u2;
  const double Le = L * e;
  const double a2 = a * a, b2 = b * b, d2 = d * d, ab = a * b;
  // axial pair f1 / f2
  const double g1 = fma(K[0], a, fma(-K[1], b, -K[2] * d));
  const double g2 = fma(K[1], a, fma(-K[3], b, K[2] * d));
  const double g3 = fma(-K[2], a + b, K[4] * d);
  const double T = fma(a, g1, -b * g2);
  const double f1 = al * (fma(L, u1, -T) - (u2 + d) * g3);
  const double f2 = al * (fma(d, g3, T) - Le);
  // transverse force f3 (= -f5).  The cubic forms are grouped by their common factors (a^3 + b^3 = (a + b)(a^2 - ab +
  // b^2); powers of d nested): 11 / 17 / 17 FP64 instructions for A3 / A4 / A6 instead of 16 / 23 / 24 term by term.
  const double apb = a + b, s2 = a2 + b2;
  const double P1 = fma(K[5], s2, fma(-K[28], ab, Le));                                     // K[28] = K[5] + K[6]
  const double P2 = fma(K[7], s2, fma(-12.0, Le, fma(K[10], d2, -d * fma(K[8], a, K[9] * b))));
  const double A3 = fma(apb, P1, d * P2);
  const double D3 = fma(120.0, d, -60.0 * apb);
  const double f3 = (K[11] * iL) * fma(al, A3, de * D3);
  // moment f4 (node 1)
  const double G1 = fma(K[12], a, -K[13] * b);
  const double G2 = fma(K[15], a, -K[20] * b);
  const double G3 = fma(K[14], a2 - b2, fma(-K[16], ab, fma(K[11], Le, d * fma(K[19], a, -K[21] * d))));
  const double G5 = fma(K[18], b, -K[17] * a);
  const double A4 = fma(a2, G1, fma(b2, G2, fma(d, G3, Le * G5)));
  const double f4 = fma(al, A4, de * fma(4.0, a, fma(2.0, b, -6.0 * d)));
  // moment f6 (node 2)
  const double H1 = fma(-K[20], a, fma(K[22], b, -K[23] * d));
  const double H2 = fma(-K[24], a, fma(K[26], b, K[24] * d));
  const double H3 = fma(-K[25], ab, fma(K[11], Le, d * fma(K[27], b, -K[19] * d)));
  const double H5 = fma(K[18], a, -K[17] * b);
  const double A6 = fma(a2, H1, fma(b2, H2, fma(d, H3, Le * H5)));
  const double f6 = fma(al, A6, de * fma(2.0, a, fma(4.0, b, -6.0 * d)));
  bA[0] -= f1;
  bA[1] -= f3;
  bA[2] -= f4;
  bB[0] -= f2;
  bB[1] += f3;
  bB[2] -= f6;
}

// ------------------------------------------------------------------------------------------
// Per-lane constant context of one member (loaded once per launch / per member batch).
// ------------------------------------------------------------------------------------------
template <int M>
struct LaneCtx {
  int G, g, levels;
  int lane;              // lane id in the warp
  int member;            // clamped member index
  bool active;           // member < B
  int n;                 // free position DOFs
  int nseg;              // number of CSV rows (= elements)
  int ri[M][3];          // reduced index of own DOFs (-1: constrained / phantom)
  int et[M];             // type of the element left of own slot j
  double4 kc[M];         // stiffness coefficients of that element
  double imp_amp;        // impulse amplitude of this member (0 if none) ...
  int imp_local;         // ... and the own DOF (3 j + d) it acts on, or -1
  double drag[M];        // drag factor of own slot (0 = none)
  double gl[M], gt[M];   // gravity half masses: two-ended pseudo-segment left of slot, tail
  MassConsts<M> mc;
  FastMass fm;           // uniform-mass solve constants (shape-specialised kernels)
  double* scratch;       // shared-memory scratch of this member (2n doubles) or nullptr
  const double* gain_sm; // this warp's per-member gains staged in shared memory as [c][3 j + d][lane], or nullptr
};

struct RhsFlags {
  bool drag, grav_slot, grav_generic, mask, uconst, impulse, gain, fext, utime;
};

// Compile-time feature sets: a kernel instantiated for a feature set contains only that code
// (the full RHS is > 60 KB of SASS; the instruction cache is 32 KB).  The host picks the smallest
// set that covers what the system enables.
enum : unsigned {
  CRB_F_LINEAR = 1, CRB_F_NONLIN = 2, CRB_F_DRAG = 4, CRB_F_GRAVS = 8, CRB_F_GRAVG = 16, CRB_F_MASK = 32,
  CRB_F_INPUT = 64, CRB_F_GAINM = 128, CRB_F_GAINS = 256, CRB_F_ALL = 511,
  // profile A: linear elements, slot gravity, inputs, tensor-core feedback (configs 1, 5)
  CRB_F_PROFILE_A = CRB_F_LINEAR | CRB_F_GRAVS | CRB_F_INPUT | CRB_F_GAINM,
  // profile B: nonlinear elements, drag, slot gravity, inputs (configs 2, 4)
  CRB_F_PROFILE_B = CRB_F_NONLIN | CRB_F_DRAG | CRB_F_GRAVS | CRB_F_INPUT,
  // profile C: profile A with the feedback product on the reduced-vector path (one gain PER MEMBER, or lane
  // layouts other than 4 lanes per member): LQR rollout of a design ensemble
  CRB_F_PROFILE_C = CRB_F_LINEAR | CRB_F_GRAVS | CRB_F_INPUT | CRB_F_GAINS,
};

__device__ __forceinline__ RhsFlags make_flags(const crb_system_t& s, const KPlan& p) {
  RhsFlags f;
  f.drag = s.drag != nullptr;
  f.grav_slot = s.grav_mode == 1;
  f.grav_generic = s.grav_mode == 2;
  f.mask = p.has_mask != 0;
  f.uconst = s.u_const != nullptr;
  f.impulse = s.imp_amp != nullptr;
  f.gain = s.gain != nullptr;
  f.fext = s.f_ext != nullptr;
  f.utime = s.u_sin_amp != nullptr || s.u_tab_v != nullptr;
  return f;
}

template <int M>
__device__ __forceinline__ void load_lane_ctx(LaneCtx<M>& L, const KPlan& P, const crb_system_t& S,
                                              int member, int g, const double* mfac_smem,
                                              double* scratch) {
  L.G = P.g;
  L.g = g;
  L.lane = threadIdx.x & 31;
  L.levels = P.levels;
  L.n = P.n_free;
  L.nseg = P.N;
  L.active = member < S.n_members;
  L.member = L.active ? member : S.n_members - 1;
  L.scratch = scratch;
  L.gain_sm = nullptr;
  const int s0 = g * M;
  const double* kc = S.kcoef + (S.stiff_shared ? 0ll : (long long)L.member * P.p * 4);
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const int s = s0 + j;
#pragma unroll
    for (int d = 0; d < 3; ++d) L.ri[j][d] = S.red_index[3 * s + d];
    L.et[j] = S.elem_type[s];
    const double2 k0 = *reinterpret_cast<const double2*>(kc + 4 * s);
    const double2 k1 = *reinterpret_cast<const double2*>(kc + 4 * s + 2);
    L.kc[j] = make_double4(k0.x, k0.y, k1.x, k1.y);
    const long long fo = S.force_shared ? 0ll : (long long)L.member * P.p;
    L.drag[j] = S.drag ? S.drag[fo + s] : 0.0;
    L.gl[j] = (S.grav && S.grav_mode == 1) ? S.grav[2 * (fo + s)] : 0.0;
    L.gt[j] = (S.grav && S.grav_mode == 1) ? S.grav[2 * (fo + s) + 1] : 0.0;
  }
  L.imp_amp = S.imp_amp ? S.imp_amp[L.member] : 0.0;
  L.imp_local = -1;
  if (S.imp_amp) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (L.ri[j][d] == S.imp_dof) L.imp_local = 3 * j + d;
  }
  const double* mf = S.mass_shared ? mfac_smem : S.mfac + (long long)L.member * P.mfac_doubles;
  L.mc.slot = mf;
  L.mc.scan = mf + 2 * CRB_SLOT_PAIRS * P.p;
  L.mc.G = P.g;
  L.mc.g = g;
}

// fdlibm kernel polynomial coefficients of crb_sincos (sin: S6..S1, cos: C6..C1, -1/2)
static __constant__ double kSinCos[13] = {
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04,
    8.33333333332248946124e-03, -1.66666666666666324348e-01,
    -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05,
    -1.38888888888741095749e-03, 4.16666666666666019037e-02, -0.5};

// sin and cos of a rotation angle, inlinable and branch-free on the common path: two-constant FMA
// Cody-Waite reduction by pi/2 (x - n c1 is exact in an FMA) and the fdlibm kernel polynomials
// (< 1 ulp on [-pi/4, pi/4]); about 25 FP64 instructions instead of a call into the library routine
// with its large-argument path (the gravity term evaluates one per segment and RHS,
// gravity_forces.py:117-125).  |x| >= 1e5 (never a beam rotation) falls back to the library.
// SMALL_PATH: a warp vote skips the reduction and the quadrant selects when |x| <= pi/4 on every active lane (same
// polynomials on the same argument: bitwise the same result).  Worth it only where the sincos is a large share of a
// short right-hand side: the shared-operator kernel gains 4 % (8.18 -> 7.87 ms on config 5); in the banded kernels
// the extra branch costs more than it saves (config 1 -8 %, config 4 -6 %, dense rollout -8 %: measured), so it is opt-in.
template <bool SMALL_PATH = false, bool CONST_COEF = false>
__device__ __forceinline__ void crb_sincos(double x, double& sn, double& cs) {
  if (!(fabs(x) < 1.0e5)) {
    sincos(x, &sn, &cs);
    return;
  }
  // |x| <= pi/4 on every active lane (beam rotations: the common case): the reduction below would find n = 0, r = x
  // and quadrant 0
  const bool small = SMALL_PATH && __all_sync(__activemask(), fabs(x) <= 0.78539816339744828);
  double n = 0.0, r = x;
  int q = 0;
  if (!small) {
    n = rint(x * 6.36619772367581382433e-01);
    r = fma(-n, 1.57079632679489655800e+00, x);
    r = fma(-n, 6.12323399573676603587e-17, r);
    q = (int)n;
  }
  const double z = r * r;
  // CONST_COEF: coefficients as constant-bank operands instead of literals.  A literal costs two uniform moves, which
  // the compiler hoists out of a step loop when it has uniform registers to spare: the banded gravity kernel (config 1)
  // is 9 % FASTER with literals, the shared-operator kernel 4 % and the dense per-member kernel 1.5 % faster with the
  // constant bank (229 -> 149 UMOV in the shared-operator kernel's SASS); config 4 does not care.  Measured, so opt-in.
#define CRB_SC(i, lit) (CONST_COEF ? kSinCos[i] : (lit))
  double ps = fma(z, CRB_SC(0, 1.58969099521155010221e-10), CRB_SC(1, -2.50507602534068634195e-08));
  ps = fma(z, ps, CRB_SC(2, 2.75573137070700676789e-06));
  ps = fma(z, ps, CRB_SC(3, -1.98412698298579493134e-04));
  ps = fma(z, ps, CRB_SC(4, 8.33333333332248946124e-03));
  ps = fma(z, ps, CRB_SC(5, -1.66666666666666324348e-01));
  const double s0 = fma(r * z, ps, r);
  double pc = fma(z, CRB_SC(6, -1.13596475577881948265e-11), CRB_SC(7, 2.08757232129817482790e-09));
  pc = fma(z, pc, CRB_SC(8, -2.75573143513906633035e-07));
  pc = fma(z, pc, CRB_SC(9, 2.48015872894767294178e-05));
  pc = fma(z, pc, CRB_SC(10, -1.38888888888741095749e-03));
  pc = fma(z, pc, CRB_SC(11, 4.16666666666666019037e-02));
  const double c0 = fma(z * z, pc, fma(CRB_SC(12, -0.5), z, 1.0));
#undef CRB_SC
  if (small) {
    sn = s0;
    cs = c0;
    return;
  }
  const double a = (q & 1) ? c0 : s0, b = (q & 1) ? s0 : c0;
  sn = (q & 2) ? -a : a;
  cs = ((q + 1) & 2) ? -b : b;
}

// Gravity contribution of one pseudo-segment (gravity_forces.py:117-125).
template <bool CONST_COEF = false>
__device__ __forceinline__ void grav_pair(double phi, double hm, double gx, double gy, double& fa, double& ft) {
  double sn, cs;
  crb_sincos<false, CONST_COEF>(phi, sn, cs);
  fa = fma(cs, gx, sn * gy) * hm;
  ft = fma(-sn, gx, cs * gy) * hm;
}

// a = M^-1 ( -k(q) + f(x) + u(t) ) for the lane's slots.  q, v: stage state; out: acceleration.
// ONLY_FORCES: skip stiffness, inputs and the mass solve -> acc = built-in force vector f(x)
// LVU >= 0: uniform-mass solve with 1 << LVU lanes per member (L.fm); LVU = -1: stored-spike mass_solve (L.mc)
template <int M, unsigned FEAT = CRB_F_ALL, bool ONLY_FORCES = false, int LVU = -1>
__device__ __forceinline__ void beam_accel(const LaneCtx<M>& L, const crb_system_t& S, const RhsFlags F0,
                                           const double (&q)[M][3], const double (&v)[M][3],
                                           double t, double (&acc)[M][3]) {
  const int G = L.G;
  RhsFlags F;
  F.drag = (FEAT & CRB_F_DRAG) && F0.drag;
  F.grav_slot = (FEAT & CRB_F_GRAVS) && F0.grav_slot;
  F.grav_generic = (FEAT & CRB_F_GRAVG) && F0.grav_generic;
  F.mask = (FEAT & CRB_F_MASK) && F0.mask;
  F.uconst = (FEAT & CRB_F_INPUT) && F0.uconst;
  F.impulse = (FEAT & CRB_F_INPUT) && F0.impulse;
  F.fext = (FEAT & CRB_F_INPUT) && F0.fext;
  F.utime = (FEAT & CRB_F_INPUT) && F0.utime;
  F.gain = (FEAT & (CRB_F_GAINM | CRB_F_GAINS)) && F0.gain;
  constexpr bool kLin = (FEAT & CRB_F_LINEAR) != 0, kNl = (FEAT & CRB_F_NONLIN) != 0;
  // halo: q of the slot left of this lane's first slot (zero at the root / outside the beam)
  double qh[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    qh[d] = shfl_up_d(q[M - 1][d], 1, G);
    if (L.g == 0) qh[d] = 0.0;
  }
  double send[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[j][d] = 0.0;

#pragma unroll
  for (int j = 0; j < M; ++j) {
    if (ONLY_FORCES) {
    } else if (j == 0) {
      // single-type feature sets skip the type test: absent elements carry zero coefficients
      if (kLin && !kNl) elem_linear(L.kc[0], qh, q[0], send, acc[0]);
      else if (kNl && !kLin) elem_nonlinear(L.kc[0], qh, q[0], send, acc[0]);
      else if (L.et[0] == CRB_ELEM_LINEAR) elem_linear(L.kc[0], qh, q[0], send, acc[0]);
      else if (L.et[0] == CRB_ELEM_NONLINEAR) elem_nonlinear(L.kc[0], qh, q[0], send, acc[0]);
    } else {
      if (kLin && !kNl) elem_linear(L.kc[j], q[j - 1], q[j], acc[j - 1], acc[j]);
      else if (kNl && !kLin) elem_nonlinear(L.kc[j], q[j - 1], q[j], acc[j - 1], acc[j]);
      else if (L.et[j] == CRB_ELEM_LINEAR) elem_linear(L.kc[j], q[j - 1], q[j], acc[j - 1], acc[j]);
      else if (L.et[j] == CRB_ELEM_NONLINEAR) elem_nonlinear(L.kc[j], q[j - 1], q[j], acc[j - 1], acc[j]);
    }
    if (F.grav_slot) {
      // pseudo-segment between slot s-1 and s in REDUCED numbering
      if (L.gl[j] != 0.0) {
        const double pl = (j == 0) ? qh[2] : q[j == 0 ? 0 : j - 1][2];
        double fa, ft;
        grav_pair(0.5 * (pl + q[j][2]), L.gl[j], S.gx, S.gy, fa, ft);
        acc[j][0] += fa;
        acc[j][1] += ft;
        if (j == 0) { send[0] += fa; send[1] += ft; }
        else { acc[j == 0 ? 0 : j - 1][0] += fa; acc[j == 0 ? 0 : j - 1][1] += ft; }
      }
      if (L.gt[j] != 0.0) {  // last pseudo-segment: end node index falls off the reduced vector
        double fa, ft;
        grav_pair(q[j][2], L.gt[j], S.gx, S.gy, fa, ft);
        acc[j][0] += fa;
        acc[j][1] += ft;
      }
    }
    if (F.drag) {
      const double w = v[j][1];
      acc[j][1] += (-L.drag[j] * w) * fabs(w);
    }
  }
  // forces on the last slot of the left neighbour travel one lane down
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double r = shfl_down_d(send[d], 1, G);
    if (L.g == G - 1) r = 0.0;
    acc[M - 1][d] += r;
  }

  // ---- state feedback u_c = gain (ref - x) as an FP64 tensor-core contraction ----
  // The gain is ONE operator shared by the whole ensemble, so the feedback of 8 members is a
  // dense [8 x 2n] x [2n x n] product: mma.sync.m8n8k4.f64 with rows = members (4 lanes each,
  // plan.g == 4), k-tiles = the lanes' own state values (no shuffle needed to build A) and n-tiles
  // = the lanes' own position DOFs (the C fragment lands where the force is consumed).
  const bool gain_mma = (FEAT & CRB_F_GAINM) && F.gain && S.gain_frag != nullptr && G == 4 && S.gain_stride == 0;
  if (gain_mma) {
    constexpr int KT = 6 * M, NT = (3 * M + 1) / 2;
    double ev[KT];
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = L.ri[j][d];
        const double rq = (S.ref && r >= 0) ? __ldg(S.ref + r) : 0.0;
        const double rv = (S.ref && r >= 0) ? __ldg(S.ref + L.n + r) : 0.0;
        ev[3 * j + d] = r >= 0 ? rq - q[j][d] : 0.0;
        ev[3 * M + 3 * j + d] = r >= 0 ? rv - v[j][d] : 0.0;
      }
    const double* fr = S.gain_frag + L.lane;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const double bfrag = __ldg(fr + (kt * NT + nt) * 32);
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0), "+d"(c1)
                     : "d"(ev[kt]), "d"(bfrag));
      }
      acc[(2 * nt) / 3][(2 * nt) % 3] += c0;
      if (2 * nt + 1 < 3 * M) acc[(2 * nt + 1) / 3][(2 * nt + 1) % 3] += c1;
    }
  }

  // ---- reduced-vector scratch paths: feedback gain (other lane layouts) and generic-BC gravity ----
  if (((FEAT & CRB_F_GAINS) && F.gain && !gain_mma) || F.grav_generic) {
    double* e = L.scratch;  // [2n]: e = x (positions then velocities) in reduced order
    __syncwarp();
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (L.ri[j][d] >= 0) {
          e[L.ri[j][d]] = q[j][d];
          e[L.n + L.ri[j][d]] = v[j][d];
        }
    __syncwarp();
    const int n = L.n;
    if ((FEAT & CRB_F_GAINS) && F.gain && !gain_mma) {
      // the lane's 3M gain rows advance together along the state (3M independent FMA chains, one read of
      // e[c] for all of them); gain_stride != 0: one gain per member (crb_lqr_gains)
      const double* gbase = S.gain + (long long)L.member * S.gain_stride;
      int roff[M][3];
      double sacc[M][3];
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          roff[j][d] = (L.ri[j][d] >= 0 ? L.ri[j][d] : 0) * 2 * n;
          sacc[j][d] = 0.0;
        }
      if (L.gain_sm) {  // staged per-member gains: conflict-free LDS, 3M independent FMA chains
        const double* gs = L.gain_sm + L.lane;
#pragma unroll 4
        for (int c = 0; c < 2 * n; ++c) {
          const double ec = (S.ref ? __ldg(S.ref + c) : 0.0) - e[c];
#pragma unroll
          for (int j = 0; j < M; ++j)
#pragma unroll
            for (int d = 0; d < 3; ++d) sacc[j][d] = fma(gs[(c * 3 * M + 3 * j + d) * 32], ec, sacc[j][d]);
        }
      } else {
        for (int c = 0; c < 2 * n; ++c) {
          const double ec = (S.ref ? __ldg(S.ref + c) : 0.0) - e[c];
#pragma unroll
          for (int j = 0; j < M; ++j)
#pragma unroll
            for (int d = 0; d < 3; ++d) sacc[j][d] = fma(__ldg(gbase + roff[j][d] + c), ec, sacc[j][d]);
        }
      }
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          if (L.ri[j][d] >= 0) acc[j][d] += sacc[j][d];
    }
    if (F.grav_generic) {
      // gravity_forces.py:97-146 evaluated in REDUCED indices.  Pass 1: the lanes of the
      // member share the pseudo-segments (one sincos each; rotations read at reduced 3i+2, 3i+5) and
      // leave (axial, transverse) half-forces in shared memory.  Pass 2: DOF r = 3k + c (c = 0 axial,
      // 1 transverse) collects segment k (as its start node) and segment k-1 (as its end node).
      const double* hm = S.seg_half_mass + (S.force_shared ? 0ll : (long long)L.member * L.nseg);
      double* fseg = e + 2 * n;  // [nseg][2]
      for (int i = L.g; i < L.nseg; i += G) {
        const int ia = 3 * i + 2, ib = 3 * i + 5;
        double phi = 0.0;
        if (ia < n && ib < n) phi = 0.5 * (e[ia] + e[ib]);
        else if (ia < n) phi = e[ia];
        else if (ib < n) phi = e[ib];
        double fa, ft;
        grav_pair(phi, hm[i], S.gx, S.gy, fa, ft);
        fseg[2 * i] = fa;
        fseg[2 * i + 1] = ft;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int r = L.ri[j][d];
          if (r < 0) continue;
          const int c = r % 3, k = r / 3;
          if (c == 2) continue;
          double tot = 0.0;
          if (k - 1 >= 0 && k - 1 < L.nseg) tot += fseg[2 * (k - 1) + c];
          if (k < L.nseg) tot += fseg[2 * k + c];
          acc[j][d] += tot;
        }
    }
  }

  // ---- inputs u(t): constant part, impulse, external force (dynamic_beam_model.py:357-362) ----
  if (!ONLY_FORCES && F.impulse && L.imp_local >= 0 && t < S.imp_duration) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (L.imp_local == 3 * j + d) acc[j][d] += L.imp_amp;
  }
  if (!ONLY_FORCES && (F.uconst || F.fext)) {
    const long long mo = (long long)L.member * L.n;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = L.ri[j][d];
        if (r >= 0) {
          double u = 0.0;
          if (F.uconst) u += S.u_const[mo + r];
          if (F.fext) u += S.f_ext[mo + r];
          acc[j][d] += u;
        }
      }
  }
  // time-varying inputs u(t) at the stage time: sinusoid and piecewise-linear table (crb_system_t.u_sin_* / u_tab_*)
  if (!ONLY_FORCES && F.utime) {
    double sv = 0.0, wt = 0.0;
    int k0 = 0;
    if (S.u_sin_amp) {
      double cs;
      crb_sincos(fma(S.u_sin_omega, t, S.u_sin_phase), sv, cs);
    }
    if (S.u_tab_v) {
      const int K = S.u_tab_k;
      if (t >= S.u_tab_t[K - 1]) {
        k0 = K - 2;
        wt = 1.0;
      } else if (t > S.u_tab_t[0]) {
        int lo = 0, hi = K - 1;  // u_tab_t[lo] <= t < u_tab_t[hi]
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (S.u_tab_t[mid] <= t) lo = mid; else hi = mid;
        }
        k0 = lo;
        wt = (t - S.u_tab_t[lo]) / (S.u_tab_t[lo + 1] - S.u_tab_t[lo]);
      }
    }
    const long long mo = S.u_time_shared ? 0ll : (long long)L.member * L.n;
    const long long ks = S.u_time_shared ? (long long)L.n : (long long)S.n_members * L.n;  // doubles between knots
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = L.ri[j][d];
        if (r >= 0) {
          double u = 0.0;
          if (S.u_sin_amp) u = S.u_sin_amp[mo + r] * sv;
          if (S.u_tab_v) {
            const double va = S.u_tab_v[k0 * ks + mo + r], vb = S.u_tab_v[(k0 + 1) * ks + mo + r];
            u += fma(wt, vb - va, va);
          }
          acc[j][d] += u;
        }
      }
  }
  if (F.mask) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (L.ri[j][d] < 0) acc[j][d] = 0.0;
  }
  if (!ONLY_FORCES) {
    if constexpr (LVU >= 0) {
      double b[1][M][3];
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) b[0][j][d] = acc[j][d];
      fast_solve_r<M, (LVU >= 0 ? LVU : 0), 1>(b, L.fm);
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) acc[j][d] = b[0][j][d];
    } else {
      mass_solve<M>(acc, L.mc, L.levels);
    }
  }
}

// ------------------------------------------------------------------------------------------
// launch geometry, shared-memory staging and state I/O shared by all kernels
// ------------------------------------------------------------------------------------------
#define CRB_WARPS_PER_BLOCK 4
#define CRB_THREADS (32 * CRB_WARPS_PER_BLOCK)

// Shared memory: [mfac set (if mass_shared)] [per-member scratch of 2n doubles (if needed)]
struct SmemLayout {
  int mfac_doubles;     // 0 if not staged
  int scratch_doubles;  // per member, 0 if unused
};

__device__ __forceinline__ const double* stage_mfac(const crb_system_t& S, const KPlan& P, double* smem) {
  if (!S.mass_shared) return nullptr;
  // the stored-spike part of the factor set (the compact copy behind it is staged by stage_compact)
  const int count = 2 * CRB_SLOT_PAIRS * P.p + 2 * CRB_SCAN_PAIRS * (P.levels > 0 ? P.levels : 1) * P.g;
  for (int k = threadIdx.x; k < count; k += blockDim.x) smem[k] = S.mfac[k];
  __syncthreads();
  return smem;
}

// Shape-specialised kernels: stage only the compact copy of the shared factor set.
template <int M, int LV>
__device__ __forceinline__ void stage_compact(const crb_system_t& S, double* smem, FastMass& fm, int g) {
  constexpr int G = 1 << LV, LVE = LV > 0 ? LV : 1;
  constexpr int FAST_DOUBLES = crb_compact_doubles(M, G, LVE);
  const double* src = S.mfac + 2 * CRB_SLOT_PAIRS * (M * G) + 2 * CRB_SCAN_PAIRS * LVE * G;
  for (int k = threadIdx.x; k < FAST_DOUBLES; k += blockDim.x) smem[k] = src[k];
  __syncthreads();
  fm.g = g;
  fm.fslot = smem;
  fm.fscan = smem + crb_compact_slot_doubles(M, G);
}

// Per-member factor sets: every lane group stages its member's compact copy in its own shared-memory region.
template <int M, int LV>
__device__ __forceinline__ void stage_compact_pm(const crb_system_t& S, const KPlan& P, double* smem, FastMass& fm, int g,
                                                 int mloc, int member) {
  constexpr int G = 1 << LV, LVE = LV > 0 ? LV : 1;
  constexpr int FAST_DOUBLES = crb_compact_doubles(M, G, LVE);
  const double* src = S.mfac + (long long)member * P.mfac_doubles + 2 * CRB_SLOT_PAIRS * (M * G) + 2 * CRB_SCAN_PAIRS * LVE * G;
  double* dst = smem + mloc * FAST_DOUBLES;
  for (int k = g; k < FAST_DOUBLES; k += G) dst[k] = src[k];
  __syncthreads();
  fm.g = g;
  fm.fslot = dst;
  fm.fscan = dst + crb_compact_slot_doubles(M, G);
}

// One (q, v) pair of a recorded frame.  Full frames are state rows [q ; v]; lean frames (crb_system_t.out_sel_inv)
// hold only the selected entries, at the columns the inverse selection table names.
__device__ __forceinline__ int frame_width(const crb_system_t& S, int n) { return S.out_sel_inv ? S.out_n_sel : 2 * n; }
__device__ __forceinline__ void frame_put(const int* __restrict__ sel_inv, double* __restrict__ ym, int n, int r, double qv,
                                          double vv) {
  if (sel_inv) {
    const int cq = sel_inv[r], cv = sel_inv[n + r];
    if (cq >= 0) ym[cq] = qv;
    if (cv >= 0) ym[cv] = vv;
  } else {
    ym[r] = qv;
    ym[n + r] = vv;
  }
}

template <int M>
__device__ __forceinline__ void load_state(const LaneCtx<M>& L, const double* __restrict__ X,
                                           double (&q)[M][3], double (&v)[M][3]) {
  const double* x = X + (long long)L.member * 2 * L.n;
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int r = L.ri[j][d];
      q[j][d] = r >= 0 ? x[r] : 0.0;
      v[j][d] = r >= 0 ? x[L.n + r] : 0.0;
    }
}
// frame `frame` of a recording Y[T, B, width] (full or lean)
template <int M>
__device__ __forceinline__ void store_frame(const LaneCtx<M>& L, const crb_system_t& S, double* __restrict__ Y, long long frame,
                                            const double (&q)[M][3], const double (&v)[M][3]) {
  if (!L.active) return;
  double* ym = Y + (frame * S.n_members + L.member) * frame_width(S, L.n);
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int r = L.ri[j][d];
      if (r >= 0) frame_put(S.out_sel_inv, ym, L.n, r, q[j][d], v[j][d]);
    }
}
template <int M>
__device__ __forceinline__ void store_state(const LaneCtx<M>& L, double* __restrict__ X,
                                            const double (&q)[M][3], const double (&v)[M][3]) {
  if (!L.active) return;
  double* x = X + (long long)L.member * 2 * L.n;
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int r = L.ri[j][d];
      if (r >= 0) {
        x[r] = q[j][d];
        x[L.n + r] = v[j][d];
      }
    }
}

