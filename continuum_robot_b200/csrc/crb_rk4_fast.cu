// crb_rk4_fast.cu -- launcher of the linear / uniform-mass fused RK4 kernel.
#include "crb_internal.h"
#include "crb_rk4_fast.cuh"

template <int M, int LV, bool PM>
static int launch(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                  double* Y_out, int save_every, cudaStream_t stream) {
  const int mpb = CRB_FAST_WARPS * (32 >> LV);
  // compact factor copy: ONE shared set, or (PM) one region per member of the block
  const size_t bytes = sizeof(double) * (size_t)crb_compact_doubles(plan->m, plan->g, plan->levels) * (PM ? mpb : 1);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  const bool uc = sys->u_const || sys->f_ext, imp = sys->imp_amp != nullptr;
#define CRB_LIN2N(UCV, IMPV, NCV)                                                                             \
  {                                                                                                            \
    if (int rc = set_smem(crb_rk4_lin2_kernel<M, LV, UCV, IMPV, PM, NCV>, bytes, "crb_rk4")) return rc;        \
    crb_rk4_lin2_kernel<M, LV, UCV, IMPV, PM, NCV><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, X, t0, h, nsteps, \
                                                                                             Y_out, save_every); \
  }
#define CRB_LIN2(UCV, IMPV) CRB_LIN2N(UCV, IMPV, false)
  const bool nc = !(plan->contiguous && plan->p_act == plan->p);
  if (sys->grav_mode == 1) {  // slot-space gravity: stage-by-stage kernel (the force is nonlinear in the rotations)
    if (PM || uc) return 1;
    if constexpr (!PM) {
#define CRB_FASTG(IMPV, NCV)                                                                                  \
  {                                                                                                           \
    if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, IMPV, true, NCV>, bytes, "crb_rk4")) return rc;          \
    crb_rk4_fast_kernel<M, LV, IMPV, true, NCV><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, X, t0, h, nsteps, \
                                                                                          Y_out, save_every); \
  }
      if (imp && nc) CRB_FASTG(true, true)
      else if (imp) CRB_FASTG(true, false)
      else if (nc) CRB_FASTG(false, true)
      else CRB_FASTG(false, false)
#undef CRB_FASTG
    }
    return 0;
  }
  // constrained DOFs inside active slots (PINNED root, interior supports ...) or phantom slots: the NC variants
  // (reduced-index table for state I/O, masked right-hand sides); shared mass factors only
  if (nc) {
    if (PM || sys->force_staged) return 1;
    if constexpr (!PM) {
      if (uc && imp) CRB_LIN2N(true, true, true)
      else if (uc) CRB_LIN2N(true, false, true)
      else if (imp) CRB_LIN2N(false, true, true)
      else CRB_LIN2N(false, false, true)
    }
    return 0;
  }
  if (!sys->force_staged || PM) {  // paired operator applications (forcing piecewise constant in time)
    if (uc && imp) CRB_LIN2(true, true)
    else if (uc) CRB_LIN2(true, false)
    else if (imp) CRB_LIN2(false, true)
    else CRB_LIN2(false, false)
  } else if (uc) {
    return 1;  // the stage-by-stage fast kernel has no constant-force path: use the general kernel
  } else if (imp) {
    if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, true>, bytes, "crb_rk4")) return rc;
    crb_rk4_fast_kernel<M, LV, true><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, X, t0, h, nsteps, Y_out, save_every);
  } else {
    if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, false>, bytes, "crb_rk4")) return rc;
    crb_rk4_fast_kernel<M, LV, false><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, X, t0, h, nsteps, Y_out, save_every);
  }
#undef CRB_LIN2
#undef CRB_LIN2N
  return 0;
}

int crb_fast_members_per_sm(int members_per_warp) { return CRB_FAST_MINBLOCKS * CRB_FAST_WARPS * members_per_warp; }

int crb_launch_rk4_fast(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                        int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
  if (!sys->mass_shared) {  // per-member mass factors: paired kernel with one factor region per member
    if (sys->force_staged) return 1;
#define CRB_CASE_PM(MM, LL) \
  if (plan->m == MM && plan->levels == LL) return launch<MM, LL, true>(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
    CRB_CASE_PM(4, 3) CRB_CASE_PM(4, 4) CRB_CASE_PM(4, 2) CRB_CASE_PM(3, 2) CRB_CASE_PM(3, 3) CRB_CASE_PM(3, 1) CRB_CASE_PM(4, 5)
#undef CRB_CASE_PM
    return 1;
  }
#define CRB_CASE(MM, LL) \
  if (plan->m == MM && plan->levels == LL) return launch<MM, LL, false>(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
  CRB_CASE(4, 3) CRB_CASE(4, 4) CRB_CASE(4, 5) CRB_CASE(4, 2)
  CRB_CASE(3, 1) CRB_CASE(3, 2) CRB_CASE(3, 3) CRB_CASE(2, 0) CRB_CASE(2, 1) CRB_CASE(4, 0) CRB_CASE(3, 0) CRB_CASE(1, 0) CRB_CASE(4, 1)
  CRB_CASE(3, 4) CRB_CASE(3, 5)
#undef CRB_CASE
  return 1;  // shape not instantiated: the caller falls back to the general kernel
}

// ------------------------------------------------------------------------------------------
// implicit midpoint
// ------------------------------------------------------------------------------------------
template <int M, int LV, bool PM>
static int launch_midpoint(const crb_plan_t* plan, const crb_system_t* sys, const double* afac, double* X, double t0,
                           double h, int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
  const int mpb = CRB_FAST_WARPS * (32 >> LV);
  const size_t bytes = sizeof(double) * (size_t)crb_compact_doubles(plan->m, plan->g, plan->levels) * (PM ? mpb : 1);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  const bool uc = sys->u_const || sys->f_ext, imp = sys->imp_amp != nullptr;
#define CRB_MID(UCV, IMPV, NCV)                                                                                \
  {                                                                                                           \
    if (int rc = set_smem(crb_midpoint_kernel<M, LV, UCV, IMPV, PM, NCV>, bytes, "crb_midpoint")) return rc;  \
    crb_midpoint_kernel<M, LV, UCV, IMPV, PM, NCV><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, afac, X, t0, h, nsteps, \
                                                                                             Y_out, save_every); \
  }
  if (!(plan->contiguous && plan->p_act == plan->p)) {  // any boundary conditions / phantom slots: NC variants
    if (uc || imp) CRB_MID(true, true, true)
    else CRB_MID(false, false, true)
  } else if (uc) CRB_MID(true, true, false)   // forcing variant handles both (imp_amp may be NULL)
  else if (imp) CRB_MID(false, true, false)
  else CRB_MID(false, false, false)
#undef CRB_MID
  return 0;
}

int crb_launch_midpoint(const crb_plan_t* plan, const crb_system_t* sys, const double* afac, int afac_shared, double* X,
                        double t0, double h, int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
#define CRB_CASE_MID(MM, LL)                                                                                        \
  if (plan->m == MM && plan->levels == LL)                                                                          \
    return afac_shared ? launch_midpoint<MM, LL, false>(plan, sys, afac, X, t0, h, nsteps, Y_out, save_every, stream) \
                       : launch_midpoint<MM, LL, true>(plan, sys, afac, X, t0, h, nsteps, Y_out, save_every, stream);
  CRB_CASE_MID(4, 3) CRB_CASE_MID(4, 4) CRB_CASE_MID(4, 5) CRB_CASE_MID(4, 2) CRB_CASE_MID(4, 1) CRB_CASE_MID(4, 0)
  CRB_CASE_MID(3, 1) CRB_CASE_MID(3, 2) CRB_CASE_MID(3, 3) CRB_CASE_MID(3, 0) CRB_CASE_MID(2, 0) CRB_CASE_MID(2, 1)
  CRB_CASE_MID(1, 0) CRB_CASE_MID(3, 4) CRB_CASE_MID(3, 5)
#undef CRB_CASE_MID
  return crb_fail(CRB_E_LIMIT, "crb_midpoint: lane layout m=%d, levels=%d is not instantiated", plan->m, plan->levels);
}
