// crb_rk4.cu -- general fused RK4 kernel (any element mix, forces, inputs, boundary conditions).
#include "crb_internal.h"

// Classical RK4, nsteps fused: state, stage state and the running combination stay in registers.
// LV >= 0: lanes per member = 1 << LV known at compile time (constant shuffle widths and
// shared-memory offsets); LV = -1: generic.
// PM: per-member mass factors (shape-specialised kernels only): compact solve on per-member shared-memory regions.
// GST: one gain PER MEMBER (crb_system_t.gain_stride != 0) staged in shared memory, one warp per block (a member's
// n x 2n gain is 5 KB for the 6-element LQR example: 16 members per warp fill 83 KB); without it the gains are
// re-read from L2 at every RHS evaluation.
#ifndef CRB_RK4_UMS_ALL
#define CRB_RK4_UMS_ALL 0  // 1: the nonlinear profile uses the compact mass solve too
#endif
template <int M, unsigned FEAT, int LV, bool PM = false, bool GST = false>
__global__ void __launch_bounds__(GST ? 32 : CRB_THREADS)
crb_rk4_kernel(KPlan P, crb_system_t S, SmemLayout SL, double* __restrict__ X, double t0, double h,
               int nsteps, double* __restrict__ Y, int save_every) {
  extern __shared__ __align__(16) double smem[];
  // Shape-specialised LINEAR kernels (profile A) use the compact mass solve (fast_solve_r) on the compact
  // factor copy: measured +30 % on config 3's shape (shared-memory bound with the stored-spike solve).
  // The nonlinear profile keeps the stored spikes: it is latency-bound and the compact solve's four
  // dependent sweeps measured 3-5 % slower there than two sweeps plus independent corrections.
  constexpr bool UMS = LV >= 0 && (FEAT == CRB_F_PROFILE_A || FEAT == CRB_F_PROFILE_C || PM || CRB_RK4_UMS_ALL);
  const double* mf = UMS ? smem : stage_mfac(S, P, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Gk = LV >= 0 ? (1 << (LV >= 0 ? LV : 0)) : P.g;
  const int mpw = 32 / Gk;
  const int mloc = warp * mpw + lane / Gk;
  const int member = blockIdx.x * ((GST ? 1 : CRB_WARPS_PER_BLOCK) * mpw) + mloc;
  LaneCtx<M> L;
  load_lane_ctx<M>(L, P, S, member, lane % Gk, mf,
                   SL.scratch_doubles ? smem + SL.mfac_doubles + mloc * SL.scratch_doubles : nullptr);
  if (LV >= 0) {  // let the compiler see the constants
    L.G = Gk;
    L.levels = LV;
    if (UMS && PM) {
      stage_compact_pm<M, (LV >= 0 ? LV : 0)>(S, P, smem, L.fm, lane % Gk, mloc, L.member);
    } else if (UMS) {
      stage_compact<M, (LV >= 0 ? LV : 0)>(S, smem, L.fm, lane % Gk);
    } else {
      L.mc.G = Gk;
      L.mc.slot = smem;  // specialised kernels require a shared mass set: plain LDS instead of generic loads
      L.mc.scan = smem + 2 * CRB_SLOT_PAIRS * (M * Gk);
    }
  }
  if (GST) {  // stage the own rows of this member's gain: [c][3 j + d][lane]
    double* gsm = smem + SL.mfac_doubles + mpw * SL.scratch_doubles;
    const double* gm = S.gain + (long long)L.member * S.gain_stride;
    const int n2 = 2 * L.n;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = L.ri[j][d];
        for (int c = 0; c < n2; ++c) gsm[(c * 3 * M + 3 * j + d) * 32 + lane] = r >= 0 ? gm[(long long)r * n2 + c] : 0.0;
      }
    L.gain_sm = gsm;
    __syncwarp();
  }
  const RhsFlags F = make_flags(S, P);
  double q[M][3], v[M][3];
  load_state<M>(L, X, q, v);
  const double hh = 0.5 * h, h6 = h / 6.0, h3 = h / 3.0;
  for (int k = 0; k < nsteps; ++k) {
    const double t = t0 + k * h;
    double qs[M][3], vs[M][3], aq[M][3], av[M][3], a[M][3];
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        qs[j][d] = q[j][d];
        vs[j][d] = v[j][d];
        aq[j][d] = q[j][d];
        av[j][d] = v[j][d];
      }
    // the stage loop stays rolled: ONE copy of the (large) RHS in the instruction stream
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
      const double ts = t + (st == 0 ? 0.0 : (st == 3 ? h : hh));
      beam_accel<M, FEAT, false, (UMS ? LV : -1)>(L, S, F, qs, vs, ts, a);
      const double wgt = (st == 0 || st == 3) ? h6 : h3;  // b = (1/6, 1/3, 1/3, 1/6)
      const double cn = st == 2 ? h : hh;                 // next stage: x + c k
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          aq[j][d] = fma(wgt, vs[j][d], aq[j][d]);
          av[j][d] = fma(wgt, a[j][d], av[j][d]);
          qs[j][d] = fma(cn, vs[j][d], q[j][d]);
          vs[j][d] = fma(cn, a[j][d], v[j][d]);
        }
    }
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        q[j][d] = aq[j][d];
        v[j][d] = av[j][d];
      }
    if (Y && save_every > 0 && (k + 1) % save_every == 0) {
      store_frame<M>(L, S, Y, (k + 1) / save_every - 1, q, v);
    }
  }
  store_state<M>(L, X, q, v);
}


int crb_launch_rk4_general(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                           int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
  size_t bytes;
  const SmemLayout SL = smem_layout(plan, sys, &bytes);
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  const unsigned need = crb_needed_features(plan, sys);
  const unsigned prof = crb_pick_profile(need);
#define CRB_RK4_LAUNCH(MM, PROF, LL, PMV)                                                                          \
  {                                                                                                                 \
    if (int rc = set_smem(crb_rk4_kernel<MM, PROF, LL, PMV>, bytes, "crb_rk4")) return rc;                          \
    crb_rk4_kernel<MM, PROF, LL, PMV><<<grid, CRB_THREADS, bytes, stream>>>(P, *sys, SL, X, t0, h, nsteps, Y_out, save_every); \
    return 0;                                                                                                       \
  }
  // one gain per member: one-warp blocks with the gains of the warp's members staged in shared memory
#define CRB_RK4_CASE_GST(MM, LL)                                                                                   \
  if (plan->m == MM && plan->levels == LL && prof == CRB_F_PROFILE_C && sys->gain_stride != 0) {                    \
    const int mpw = 32 >> LL;                                                                                       \
    SmemLayout SL = smem_layout(plan, sys, &bytes);                                                                 \
    SL.mfac_doubles = crb_compact_doubles(plan->m, plan->g, plan->levels) * (sys->mass_shared ? 1 : mpw);           \
    const size_t gst_bytes = sizeof(double) * ((size_t)SL.mfac_doubles + (size_t)SL.scratch_doubles * mpw +         \
                                               (size_t)2 * plan->n_free * 3 * MM * 32);                             \
    if (gst_bytes <= 113 * 1024) { /* two blocks per SM at least */                                                 \
      const int gst_grid = (sys->n_members + mpw - 1) / mpw;                                                        \
      if (sys->mass_shared) {                                                                                       \
        if (int rc = set_smem(crb_rk4_kernel<MM, CRB_F_PROFILE_C, LL, false, true>, gst_bytes, "crb_rk4")) return rc; \
        crb_rk4_kernel<MM, CRB_F_PROFILE_C, LL, false, true><<<gst_grid, 32, gst_bytes, stream>>>(P, *sys, SL, X, t0, h, nsteps, Y_out, save_every); \
      } else {                                                                                                      \
        if (int rc = set_smem(crb_rk4_kernel<MM, CRB_F_PROFILE_C, LL, true, true>, gst_bytes, "crb_rk4")) return rc; \
        crb_rk4_kernel<MM, CRB_F_PROFILE_C, LL, true, true><<<gst_grid, 32, gst_bytes, stream>>>(P, *sys, SL, X, t0, h, nsteps, Y_out, save_every); \
      }                                                                                                             \
      return 0;                                                                                                     \
    }                                                                                                               \
  }
  if (!sys->force_staged) { CRB_RK4_CASE_GST(3, 1) CRB_RK4_CASE_GST(2, 2) CRB_RK4_CASE_GST(3, 2) }
#undef CRB_RK4_CASE_GST
#define CRB_RK4_CASE(MM, LL)                                                                                       \
  if (plan->m == MM && plan->levels == LL && prof != CRB_F_ALL) {                                                   \
    if (sys->mass_shared) {                                                                                         \
      const SmemLayout SL = (prof != CRB_F_PROFILE_B || CRB_RK4_UMS_ALL) ? smem_layout_compact(plan, sys, &bytes) : smem_layout(plan, sys, &bytes); \
      if (prof == CRB_F_PROFILE_A) CRB_RK4_LAUNCH(MM, CRB_F_PROFILE_A, LL, false)                                   \
      else if (prof == CRB_F_PROFILE_C) CRB_RK4_LAUNCH(MM, CRB_F_PROFILE_C, LL, false)                              \
      else CRB_RK4_LAUNCH(MM, CRB_F_PROFILE_B, LL, false)                                                           \
    } else {                                                                                                        \
      const SmemLayout SL = smem_layout_compact_pm(plan, sys, &bytes);                                              \
      if (bytes <= 200 * 1024) {                                                                                    \
        if (prof == CRB_F_PROFILE_A) CRB_RK4_LAUNCH(MM, CRB_F_PROFILE_A, LL, true)                                  \
        else if (prof == CRB_F_PROFILE_C) CRB_RK4_LAUNCH(MM, CRB_F_PROFILE_C, LL, true)                             \
        else CRB_RK4_LAUNCH(MM, CRB_F_PROFILE_B, LL, true)                                                          \
      }                                                                                                             \
      smem_layout(plan, sys, &bytes);                                                                               \
    }                                                                                                               \
  }
  CRB_SPECIALISED_SHAPES(CRB_RK4_CASE)
#undef CRB_RK4_CASE
#undef CRB_RK4_LAUNCH
  CRB_DISPATCH_M(plan->m, {
    if (int rc = set_smem(crb_rk4_kernel<M, CRB_F_ALL, -1>, bytes, "crb_rk4")) return rc;
    crb_rk4_kernel<M, CRB_F_ALL, -1><<<grid, CRB_THREADS, bytes, stream>>>(P, *sys, SL, X, t0, h, nsteps, Y_out, save_every);
  });
  return 0;
}
