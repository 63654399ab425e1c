// crb_rk45.cu -- launcher of the adaptive Dormand-Prince kernel.
#include "crb_internal.h"
#include "crb_rk45.cuh"

int crb_launch_rk45(const crb_plan_t* plan, const crb_system_t* sys, double* X, double* t, double* h_abs,
                    double t_bound, double rtol, double atol, const double* t_eval, int n_eval, double* Y_eval,
                    int* status, long long* counters, int max_attempts, cudaStream_t stream) {
  size_t bytes;
  const SmemLayout SL = smem_layout(plan, sys, &bytes);
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  Rk45Args A;
  A.X = X; A.t = t; A.h_abs = h_abs; A.t_bound = t_bound; A.rtol = rtol; A.atol = atol;
  A.t_eval = t_eval; A.n_eval = n_eval; A.Y_eval = Y_eval; A.status = status; A.counters = counters;
  A.max_attempts = max_attempts;
  const DpTab T = make_dp_tab();
  const unsigned need = crb_needed_features(plan, sys);
  const unsigned prof = crb_pick_profile(need);
#define CRB_RK45_CASE(MM, LL)                                                                             \
  if (plan->m == MM && plan->levels == LL && prof != CRB_F_ALL && sys->mass_shared) {                      \
    const SmemLayout SL = prof == CRB_F_PROFILE_A ? smem_layout_compact(plan, sys, &bytes) : smem_layout(plan, sys, &bytes); \
    const size_t total = bytes + sizeof(double) * 21 * MM * CRB_THREADS;                                   \
    if (prof == CRB_F_PROFILE_A) {                                                                         \
      if (int rc = set_smem(crb_rk45_kernel<MM, CRB_F_PROFILE_A, LL>, total, "crb_rk45")) return rc;       \
      crb_rk45_kernel<MM, CRB_F_PROFILE_A, LL><<<grid, CRB_THREADS, total, stream>>>(P, *sys, SL, A, T);   \
    } else {                                                                                               \
      if (int rc = set_smem(crb_rk45_kernel<MM, CRB_F_PROFILE_B, LL>, total, "crb_rk45")) return rc;       \
      crb_rk45_kernel<MM, CRB_F_PROFILE_B, LL><<<grid, CRB_THREADS, total, stream>>>(P, *sys, SL, A, T);   \
    }                                                                                                      \
    return 0;                                                                                              \
  }
  CRB_SPECIALISED_SHAPES(CRB_RK45_CASE)
#undef CRB_RK45_CASE
  CRB_DISPATCH_M(plan->m, {
    const size_t total = bytes + sizeof(double) * 21 * M * CRB_THREADS;  // kv stage storage
    if (int rc = set_smem(crb_rk45_kernel<M, CRB_F_ALL, -1>, total, "crb_rk45")) return rc;
    crb_rk45_kernel<M, CRB_F_ALL, -1><<<grid, CRB_THREADS, total, stream>>>(P, *sys, SL, A, T);
  });
  return 0;
}
