// crb_midpoint.cu -- launcher of the implicit-midpoint kernels (own translation unit: parallel nvcc builds).
#include <algorithm>

#include "crb_internal.h"
#include "crb_rk4_fast.cuh"

// ------------------------------------------------------------------------------------------
// implicit midpoint
// ------------------------------------------------------------------------------------------
template <int M, int LV, bool PM>
static int launch_midpoint(const crb_plan_t* plan, const crb_system_t* sys, const double* afac, double* X, double t0,
                           double h, int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
  const int mpb = CRB_FAST_WARPS * (32 >> LV);
  // compact factor copy (one shared set, or one region per member of the block), then the staging rows of recorded
  // full-state frames (FrameWriter)
  const int fac_doubles = ((crb_compact_doubles(plan->m, plan->g, plan->levels) * (PM ? mpb : 1)) + 1) & ~1;
  // per warp: staging rows of full frames [mpw][2n], or the column table of a lean recording [6M][32] int32
  const bool stage = Y_out && (sys->out_sel_inv || ((uintptr_t)Y_out & 15) == 0);
  const int stage_off = stage ? fac_doubles : 0;
  const int stage_stride = (std::max((32 >> LV) * 2 * plan->n_free, 96 * M) + 1) & ~1;
  const size_t bytes = sizeof(double) * ((size_t)fac_doubles + (stage ? (size_t)CRB_FAST_WARPS * stage_stride : 0));
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  const bool uc = sys->u_const || sys->f_ext, imp = sys->imp_amp != nullptr;
#define CRB_MID(UCV, IMPV, NCV)                                                                                \
  {                                                                                                           \
    if (Y_out) {                                                                                               \
      if (int rc = set_smem(crb_midpoint_kernel<M, LV, UCV, IMPV, PM, NCV, true>, bytes, "crb_midpoint")) return rc; \
      crb_midpoint_kernel<M, LV, UCV, IMPV, PM, NCV, true><<<grid, CRB_FAST_THREADS, bytes, stream>>>(         \
          P, *sys, afac, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);                        \
    } else {                                                                                                   \
      if (int rc = set_smem(crb_midpoint_kernel<M, LV, UCV, IMPV, PM, NCV, false>, bytes, "crb_midpoint")) return rc; \
      crb_midpoint_kernel<M, LV, UCV, IMPV, PM, NCV, false><<<grid, CRB_FAST_THREADS, bytes, stream>>>(        \
          P, *sys, afac, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);                        \
    }                                                                                                          \
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
