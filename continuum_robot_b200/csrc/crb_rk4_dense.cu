// crb_rk4_dense.cu -- fused RK4 for an ensemble of LINEAR DESIGNS under state feedback with ONE GAIN PER MEMBER
// (the rollout that follows crb_lqr_gains; examples/lqr_control.py:95-111 with a design-specific K).
//
// With a gain per member there is no operand shared across the ensemble (no tensor-core contraction), and the
// banded kernels have to keep each member's n x 2n gain on chip next to the banded solve.  Here the whole closed
// loop of a member is ONE dense operator (crb_member_operators):
//     a = [ -M^-1 (K + G_q) | -M^-1 G_v ] [q; v] + M^-1 f_gravity(q) + impulse(t) M^-1 e_k + M^-1 G ref
// One member per warp, lane l < n owns free DOF l: its operator row (3n+1 doubles) stays in REGISTERS for the
// launch, the state is exchanged through a small shared-memory vector (broadcast reads), gravity is evaluated by
// the first nseg lanes in the reference's reduced-index form (gravity_forces.py:97-146).
// Reference behaviour: dynamic_beam_model.py:256-272, 343-362; control/full_state_linear.py:58.
#include "crb_internal.h"

#define CRB_DENSE_WARPS 4

template <int NP, bool GRAV>
__global__ void __launch_bounds__(32 * CRB_DENSE_WARPS, (NP <= 18 ? 3 : 2))
crb_rk4_dense_kernel(KPlan P, crb_system_t S, double* __restrict__ X, double t0, double h, int nsteps,
                     double* __restrict__ Y, int save_every) {
  __shared__ __align__(16) double ysm[CRB_DENSE_WARPS][3 * NP + 2 * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = P.n_free, nseg = P.N, RL = 3 * n + 1;
  const int member_raw = blockIdx.x * CRB_DENSE_WARPS + warp;
  const bool active = member_raw < S.n_members;
  const int member = active ? member_raw : S.n_members - 1;
  const bool own = lane < n;
  double* y = ysm[warp];          // [q (NP) ; v (NP) ; f_gravity (NP)]
  double* fseg = y + 3 * NP;      // [nseg][2], nseg <= 16

  // operator row of this lane's DOF
  double Wr[2 * NP], Mr[GRAV ? NP : 1], c0 = 0.0, mimp = 0.0;
  const double* row = S.member_op + ((long long)member * n + (own ? lane : 0)) * RL;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    Wr[j] = (own && j < n) ? row[j] : 0.0;
    Wr[NP + j] = (own && j < n) ? row[n + j] : 0.0;
    if (GRAV) Mr[j] = (own && j < n) ? row[2 * n + j] : 0.0;
  }
  if (own) c0 = row[3 * n];
  const double amp = S.imp_amp ? S.imp_amp[member] : 0.0;
  if (own && S.imp_amp) mimp = amp * row[2 * n + S.imp_dof];
  const double dur = S.imp_duration;
  double hm = 0.0;
  if (GRAV && lane < nseg) hm = S.seg_half_mass[(S.force_shared ? 0ll : (long long)member * nseg) + lane];
  for (int j = lane; j < 3 * NP; j += 32) y[j] = 0.0;
  __syncwarp();

  // acceleration of the own DOF for the stage state (qs, vs) at time ts
  auto accel = [&](double qs, double vs, double ts) -> double {
    if (own) {
      y[lane] = qs;
      y[NP + lane] = vs;
    }
    __syncwarp();
    double a = c0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
    if (ts < dur) a += mimp;
    if (GRAV) {
      if (lane < nseg) {  // segment-average rotation read at reduced indices 3 i + 2 and 3 i + 5
        const int ia = 3 * lane + 2, ib = 3 * lane + 5;
        double phi = 0.0;
        if (ia < n && ib < n) phi = 0.5 * (y[ia] + y[ib]);
        else if (ia < n) phi = y[ia];
        else if (ib < n) phi = y[ib];
        double fa, ft;
        grav_pair<true>(phi, hm, S.gx, S.gy, fa, ft);  // sincos coefficients from the constant bank: +1.5 % here
        fseg[2 * lane] = fa;
        fseg[2 * lane + 1] = ft;
      }
      __syncwarp();
      double fg = 0.0;
      if (own) {
        const int c = lane % 3, k = lane / 3;
        if (c < 2) {
          if (k - 1 >= 0 && k - 1 < nseg) fg += fseg[2 * (k - 1) + c];
          if (k < nseg) fg += fseg[2 * k + c];
        }
      }
      if (lane < NP) y[2 * NP + lane] = fg;
      __syncwarp();
      const double2* g2 = reinterpret_cast<const double2*>(y + 2 * NP);  // 128-bit broadcast reads
#pragma unroll
      for (int j = 0; j < NP / 2; ++j) {
        const double2 f = g2[j];
        a = fma(Mr[2 * j], f.x, a);
        a2 = fma(Mr[2 * j + 1], f.y, a2);
      }
    }
    // four accumulation chains: a dependent DFMA chain of 2 NP links would leave the FP64 pipe waiting on its own
    // latency.  (Evaluating the sincos on every lane, branch-free, to overlap it with this product measured 4 % slower.)
    const double2* y2 = reinterpret_cast<const double2*>(y);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const double2 f = y2[j];
      if (j & 1) {
        a3 = fma(Wr[2 * j], f.x, a3);
        a4 = fma(Wr[2 * j + 1], f.y, a4);
      } else {
        a = fma(Wr[2 * j], f.x, a);
        a2 = fma(Wr[2 * j + 1], f.y, a2);
      }
    }
    a = (a + a2) + (a3 + a4);
    __syncwarp();  // every lane has read the stage vector before the next stage overwrites it
    return a;
  };

  double* xm = X + (long long)member * 2 * n;
  double q = own ? xm[lane] : 0.0, v = own ? xm[n + lane] : 0.0;
  const double hh = 0.5 * h, h6 = h / 6.0;
  for (int k = 0; k < nsteps; ++k) {
    const double t = t0 + k * h;
    const double a1 = accel(q, v, t);
    const double q2 = fma(hh, v, q), v2 = fma(hh, a1, v);
    const double a2 = accel(q2, v2, t + hh);
    const double q3 = fma(hh, v2, q), v3 = fma(hh, a2, v);
    const double a3 = accel(q3, v3, t + hh);
    const double q4 = fma(h, v3, q), v4 = fma(h, a3, v);
    const double a4 = accel(q4, v4, t + h);
    q = fma(h6, v + 2.0 * v2 + 2.0 * v3 + v4, q);
    v = fma(h6, a1 + 2.0 * a2 + 2.0 * a3 + a4, v);
    if (Y && save_every > 0 && (k + 1) % save_every == 0 && active && own) {
      double* ym = Y + ((long long)((k + 1) / save_every - 1) * S.n_members + member) * frame_width(S, n);
      frame_put(S.out_sel_inv, ym, n, lane, q, v);
    }
  }
  if (active && own) {
    xm[lane] = q;
    xm[n + lane] = v;
  }
}

bool crb_dense_eligible(const crb_plan_t* plan, const crb_system_t* sys) {
  return sys->member_op && sys->gain && sys->gain_stride != 0 && sys->all_linear && !sys->drag && !sys->u_const &&
         !sys->f_ext && !crb_time_varying_input(sys) && !sys->force_general && !sys->force_staged && plan->n_free <= 32 && plan->n_elements <= 16 &&
         (sys->grav_mode == 0 || sys->seg_half_mass);
}

int crb_launch_rk4_dense(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                         double* Y_out, int save_every, cudaStream_t stream) {
  const KPlan P = kplan_of(plan);
  const int grid = (sys->n_members + CRB_DENSE_WARPS - 1) / CRB_DENSE_WARPS;
  const bool grav = sys->grav_mode != 0;
  const int n = plan->n_free;
#define CRB_DENSE_CASE(NPV)                                                                                         \
  if (n <= NPV) {                                                                                                   \
    if (grav) crb_rk4_dense_kernel<NPV, true><<<grid, 32 * CRB_DENSE_WARPS, 0, stream>>>(P, *sys, X, t0, h, nsteps, Y_out, save_every); \
    else crb_rk4_dense_kernel<NPV, false><<<grid, 32 * CRB_DENSE_WARPS, 0, stream>>>(P, *sys, X, t0, h, nsteps, Y_out, save_every);    \
    return 0;                                                                                                       \
  }
  CRB_DENSE_CASE(6) CRB_DENSE_CASE(12) CRB_DENSE_CASE(18) CRB_DENSE_CASE(24) CRB_DENSE_CASE(32)
#undef CRB_DENSE_CASE
  return crb_fail(CRB_E_LIMIT, "crb_rk4: dense per-member operator path needs n <= 32, got %d", n);
}
