// crb_rk4_fast.cu -- launcher of the linear / uniform-mass fused RK4 kernel (shared mass factor set).
#include "crb_rk4_fast_launch.cuh"

int crb_launch_rk4_fast_pm(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                           double* Y_out, int save_every, cudaStream_t stream);  // crb_rk4_fast_pm.cu

int crb_fast_members_per_sm(int members_per_warp) { return CRB_FAST_MINBLOCKS * CRB_FAST_WARPS * members_per_warp; }

int crb_launch_rk4_fast(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                        int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
  if (!sys->mass_shared)  // per-member mass factors: paired kernel with one factor region per member (crb_rk4_fast_pm.cu)
    return crb_launch_rk4_fast_pm(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
#define CRB_CASE(MM, LL) \
  if (plan->m == MM && plan->levels == LL) return launch<MM, LL, false>(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
  CRB_CASE(4, 3) CRB_CASE(4, 4) CRB_CASE(4, 5) CRB_CASE(4, 2)
  CRB_CASE(3, 1) CRB_CASE(3, 2) CRB_CASE(3, 3) CRB_CASE(2, 0) CRB_CASE(2, 1) CRB_CASE(4, 0) CRB_CASE(3, 0) CRB_CASE(1, 0) CRB_CASE(4, 1)
  CRB_CASE(3, 4) CRB_CASE(3, 5)
#undef CRB_CASE
  return 1;  // shape not instantiated: the caller falls back to the general kernel
}

