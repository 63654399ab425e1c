// crb_rk4_fast.cuh -- fused fixed-step RK4 for all-linear beams with uniform element mass
// (BASELINE config 3 shape).  Same lane-group decomposition as crb_device.cuh, specialised so
// that the FP64 pipe, not shared-memory bandwidth, is the limiter:
//
//  * Nystrom form of the classical RK4 tableau (the RHS of a linear undamped beam needs only the
//    stage positions): 5 live vectors per DOF (Q0 = q + h/2 v, v, S = a1+a2+a3, A = a1+2a2+2a3+a4,
//    work) instead of 7 -- algebraically the same update as k1..k4 of north_star row R1.
//  * mass solve without stored spike vectors: the element coupling block O of the consistent mass
//    matrix (models/segments.py:64-78) is identical for every element of a uniform-mass beam, so
//    it travels as immediate kernel arguments and only Sinv (4 doubles / slot) plus the scan
//    products are read from shared memory; boundary corrections are recomputed by a second local
//    sweep (more DFMA, 2.5x fewer LDS wavefronts than crb_device.cuh::mass_solve).
//
//   forward  A: y~_last      (zero incoming)         -> Kogge-Stone scan -> y_last, xhat_in
//   forward  B: y_s = b_s - O xhat_{s-1},  xhat_s = Sinv_s y_s   (true incoming)
//   backward A: x~_first     (zero incoming)         -> scan -> x_first, x_in
//   backward B: x_s = xhat_s - Sinv_s (O^T x_{s+1})  (true incoming)
//
// Reference behaviour: models/segments.py:32-78, euler_bernoulli_beam.py:163-298,
// dynamic_beam_model.py:256-272, 343-362 (forces disabled, u = tip impulse or none).
#pragma once
#include "crb_device.cuh"

struct UniformMass {
  double o11, o12, o22, ou;  // 54 mu, 13 L mu, 3 L^2 mu, 70 mu   (mu = rho A L / 420)
};

template <int M>
struct FastCtx {
  int g;
  int member;
  bool active;
  int n;
  double4 kc[M];
  MassConsts<M> mc;
  UniformMass um;
  // impulse: amplitude (0 if none) and the (slot, dof) it acts on, as a flat local index or -1
  double imp_amp, imp_dur;
  int imp_local;
};

template <int M, int LV>
__device__ __forceinline__ void fast_solve(double (&b)[M][3], const FastCtx<M>& C) {
  constexpr int G = 1 << LV;
  const double o11 = C.um.o11, o12 = C.um.o12, o22 = C.um.o22, ou = C.um.ou;
  // Sinv of the lane's slots stays in registers for the four sweeps
  double s00[M], s01[M], s11[M], su[M];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const double2 a = C.mc.ld(2, j), c = C.mc.ld(3, j);
    s00[j] = a.x;
    s01[j] = a.y;
    s11[j] = c.y;
    su[j] = C.mc.ld(10, j).y;
  }
  // ---- forward A: y~ at the chunk's last slot, zero incoming ----
  double y0 = b[0][0], y1 = b[0][1], y2 = b[0][2];
#pragma unroll
  for (int j = 1; j < M; ++j) {
    const double xu = su[j - 1] * y0;
    const double xw = fma(s00[j - 1], y1, s01[j - 1] * y2);
    const double xp = fma(s01[j - 1], y1, s11[j - 1] * y2);
    y0 = fma(-ou, xu, b[j][0]);
    y1 = fma(-o11, xw, fma(o12, xp, b[j][1]));
    y2 = fma(-o12, xw, fma(o22, xp, b[j][2]));
  }
  double xi0 = 0.0, xi1 = 0.0, xi2 = 0.0;  // xhat of the left neighbour's last slot
  if (G > 1) {
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      const int d = 1 << l;
      const double t0 = shfl_up_d(y0, d, G), t1 = shfl_up_d(y1, d, G), t2 = shfl_up_d(y2, d, G);
      const double2 c0 = C.mc.lds(l, 0), c1 = C.mc.lds(l, 1), cu = C.mc.lds(l, 2);
      y0 = fma(cu.x, t0, y0);
      mv2(c0, c1, t1, t2, y1, y2);
    }
    const double xu = su[M - 1] * y0;
    const double xw = fma(s00[M - 1], y1, s01[M - 1] * y2);
    const double xp = fma(s01[M - 1], y1, s11[M - 1] * y2);
    xi0 = shfl_up_d(xu, 1, G);
    xi1 = shfl_up_d(xw, 1, G);
    xi2 = shfl_up_d(xp, 1, G);
    if (C.g == 0) { xi0 = 0.0; xi1 = 0.0; xi2 = 0.0; }
  }
  // ---- forward B: true incoming; b[j] <- xhat_j ----
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const double t0 = fma(-ou, xi0, b[j][0]);
    const double t1 = fma(-o11, xi1, fma(o12, xi2, b[j][1]));
    const double t2 = fma(-o12, xi1, fma(o22, xi2, b[j][2]));
    xi0 = su[j] * t0;
    xi1 = fma(s00[j], t1, s01[j] * t2);
    xi2 = fma(s01[j], t1, s11[j] * t2);
    b[j][0] = xi0;
    b[j][1] = xi1;
    b[j][2] = xi2;
  }
  // ---- backward A: x~ at the chunk's first slot, zero incoming ----
  double r0 = b[M - 1][0], r1 = b[M - 1][1], r2 = b[M - 1][2];
#pragma unroll
  for (int j = M - 2; j >= 0; --j) {
    const double t0 = ou * r0;
    const double t1 = fma(o11, r1, o12 * r2);
    const double t2 = -fma(o12, r1, o22 * r2);
    r0 = fma(-su[j], t0, b[j][0]);
    r1 = b[j][1] - fma(s00[j], t1, s01[j] * t2);
    r2 = b[j][2] - fma(s01[j], t1, s11[j] * t2);
  }
  double n0 = 0.0, n1 = 0.0, n2 = 0.0;  // x of the right neighbour's first slot
  if (G > 1) {
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      const int d = 1 << l;
      const double t0 = shfl_down_d(r0, d, G), t1 = shfl_down_d(r1, d, G), t2 = shfl_down_d(r2, d, G);
      const double2 c0 = C.mc.lds(l, 3), c1 = C.mc.lds(l, 4), cu = C.mc.lds(l, 5);
      r0 = fma(cu.x, t0, r0);
      mv2(c0, c1, t1, t2, r1, r2);
    }
    n0 = shfl_down_d(r0, 1, G);
    n1 = shfl_down_d(r1, 1, G);
    n2 = shfl_down_d(r2, 1, G);
    if (C.g == G - 1) { n0 = 0.0; n1 = 0.0; n2 = 0.0; }
  }
  // ---- backward B: true incoming; b[j] <- x_j ----
#pragma unroll
  for (int j = M - 1; j >= 0; --j) {
    const double t0 = ou * n0;
    const double t1 = fma(o11, n1, o12 * n2);
    const double t2 = -fma(o12, n1, o22 * n2);
    n0 = fma(-su[j], t0, b[j][0]);
    n1 = b[j][1] - fma(s00[j], t1, s01[j] * t2);
    n2 = b[j][2] - fma(s01[j], t1, s11[j] * t2);
    b[j][0] = n0;
    b[j][1] = n1;
    b[j][2] = n2;
  }
}

// a <- M^-1 (-K w + impulse(t));  `w` holds the stage positions on entry, accelerations on exit.
template <int M, int LV, bool IMP>
__device__ __forceinline__ void fast_accel(const FastCtx<M>& C, double (&w)[M][3], double t) {
  constexpr int G = 1 << LV;
  double qh[3], send[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    qh[d] = shfl_up_d(w[M - 1][d], 1, G);
    if (C.g == 0) qh[d] = 0.0;
  }
  double b[M][3];
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) b[j][d] = 0.0;
  elem_linear(C.kc[0], qh, w[0], send, b[0]);
#pragma unroll
  for (int j = 1; j < M; ++j) elem_linear(C.kc[j], w[j - 1], w[j], b[j - 1], b[j]);
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double r = shfl_down_d(send[d], 1, G);
    if (C.g == G - 1) r = 0.0;
    b[M - 1][d] += r;
  }
  if (IMP && C.imp_local >= 0 && t < C.imp_dur) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (C.imp_local == 3 * j + d) b[j][d] += C.imp_amp;
  }
  fast_solve<M, LV>(b, C);
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) w[j][d] = b[j][d];
}

template <int M, int LV, bool IMP>
__global__ void __launch_bounds__(CRB_THREADS)
crb_rk4_fast_kernel(KPlan P, crb_system_t S, UniformMass um, double* __restrict__ X, double t0, double h,
                    int nsteps, double* __restrict__ Y, int save_every) {
  extern __shared__ __align__(16) double smem[];
  const double* mf = stage_mfac(S, P, smem);
  constexpr int G = 1 << LV, mpw = 32 / G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int member = blockIdx.x * (CRB_WARPS_PER_BLOCK * mpw) + warp * mpw + lane / G;
  FastCtx<M> C;
  C.g = lane % G;
  C.n = P.n_free;
  C.active = member < S.n_members;
  C.member = C.active ? member : S.n_members - 1;
  C.um = um;
  const int s0 = C.g * M;
  {
    const double* kc = S.kcoef + (S.stiff_shared ? 0ll : (long long)C.member * (M * G) * 4);
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double2 k0 = *reinterpret_cast<const double2*>(kc + 4 * (s0 + j));
      const double2 k1 = *reinterpret_cast<const double2*>(kc + 4 * (s0 + j) + 2);
      C.kc[j] = make_double4(k0.x, k0.y, k1.x, k1.y);
    }
  }
  C.mc.slot = mf;
  C.mc.scan = mf + 2 * CRB_SLOT_PAIRS * (M * G);
  C.mc.G = G;
  C.mc.g = C.g;
  C.imp_amp = (IMP && S.imp_amp) ? S.imp_amp[C.member] : 0.0;
  C.imp_dur = S.imp_duration;
  C.imp_local = -1;
  if (IMP && S.imp_amp) {
    const int rel = S.imp_dof - 3 * s0;  // contiguous plan: reduced index = 3 slot + dof
    if (rel >= 0 && rel < 3 * M) C.imp_local = rel;
  }

  // contiguous plan: the lane's 3M position DOFs are consecutive in the reduced vector
  const int n = C.n;
  double* xq = X + (long long)C.member * 2 * n + 3 * s0;
  double Q0[M][3], v[M][3], Sa[M][3], Aa[M][3], w[M][3];
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      Q0[j][d] = xq[3 * j + d];  // holds q until the first stage is set up
      v[j][d] = xq[n + 3 * j + d];
    }
  const double hh = 0.5 * h, h6 = h / 6.0, hq = 0.25 * h * h, hs = 0.5 * h * h, hx = h * h / 6.0;
  for (int k = 0; k < nsteps; ++k) {
    const double t = t0 + k * h;
    // Nystrom form of the classical tableau; the stage loop is kept rolled so the step body
    // (one copy of the RHS) stays inside the instruction cache.
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        w[j][d] = Q0[j][d];                     // stage 1 input: q
        Q0[j][d] = fma(hh, v[j][d], Q0[j][d]);  // Q0 = q + h/2 v
        Sa[j][d] = 0.0;
        Aa[j][d] = 0.0;
      }
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
      const double ts = t + (st == 0 ? 0.0 : (st == 3 ? h : hh));
      fast_accel<M, LV, IMP>(C, w, ts);
      // weights: S = a1+a2+a3, A = a1+2a2+2a3+a4; next input:
      //   st0 -> Q0 ; st1 -> Q0 + h^2/4 a1 ; st2 -> Q0 + h/2 v + h^2/2 a2
      const double wa = (st == 1 || st == 2) ? 2.0 : 1.0;
      const double ws = st == 3 ? 0.0 : 1.0;
      const double ca = st == 1 ? hq : (st == 2 ? hs : 0.0);  // coefficient of the recovered stage accel
      const double cv = st == 2 ? hh : 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double a = w[j][d];
          // acceleration entering the next stage input: a1 (= S before this update) after
          // stage 2, a2 (= A - S before this update) after stage 3
          const double prev = st == 1 ? Sa[j][d] : (Aa[j][d] - Sa[j][d]);
          Sa[j][d] = fma(ws, a, Sa[j][d]);
          Aa[j][d] = fma(wa, a, Aa[j][d]);
          w[j][d] = fma(ca, prev, fma(cv, v[j][d], Q0[j][d]));
        }
    }
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        Q0[j][d] = fma(hx, Sa[j][d], fma(hh, v[j][d], Q0[j][d]));  // q+
        v[j][d] = fma(h6, Aa[j][d], v[j][d]);                      // v+
      }
    if (Y && save_every > 0 && (k + 1) % save_every == 0 && C.active) {
      double* yq = Y + ((long long)((k + 1) / save_every - 1) * S.n_members + C.member) * 2 * n + 3 * s0;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          yq[3 * j + d] = Q0[j][d];
          yq[n + 3 * j + d] = v[j][d];
        }
    }
  }
  if (C.active) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        xq[3 * j + d] = Q0[j][d];
        xq[n + 3 * j + d] = v[j][d];
      }
  }
}
