// crb_rk4_fast_pm.cu -- fast RK4 family with PER-MEMBER mass factors (own translation unit: parallel nvcc builds).
#include "crb_rk4_fast_launch.cuh"

int crb_launch_rk4_fast_pm(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                           double* Y_out, int save_every, cudaStream_t stream) {
  if (sys->force_staged) return 1;
#define CRB_CASE_PM(MM, LL) \
  if (plan->m == MM && plan->levels == LL) return launch<MM, LL, true>(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
  CRB_CASE_PM(4, 3) CRB_CASE_PM(4, 4) CRB_CASE_PM(4, 2) CRB_CASE_PM(3, 2) CRB_CASE_PM(3, 3) CRB_CASE_PM(3, 1) CRB_CASE_PM(4, 5)
#undef CRB_CASE_PM
  return 1;  // shape not instantiated: the caller falls back to the general kernel
}
