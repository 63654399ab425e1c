// crb_rk45.cu -- launcher of the adaptive Dormand-Prince kernel.
#include <cstdlib>

#include "crb_internal.h"
#include "crb_rk45.cuh"

int crb_launch_rk45(const crb_plan_t* plan, const crb_system_t* sys, double* X, double* t, double* h_abs,
                    double t_bound, double rtol, double atol, const double* t_eval, int n_eval, double* Y_eval,
                    int* status, long long* counters, int max_attempts, cudaStream_t stream) {
  constexpr int WPB = CRB_RK45_WPB;  // warps per block of the adaptive kernels (see crb_rk45.cuh)
  static const size_t pad_bytes = getenv("CRB_RK45_PAD_SMEM") ? (size_t)strtoul(getenv("CRB_RK45_PAD_SMEM"), nullptr, 10) : 0;  // occupancy experiments only
  size_t bytes;
  SmemLayout SL = smem_layout(plan, sys, &bytes);
  const int mpb = WPB * (32 / plan->g);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  Rk45Args A;
  A.X = X; A.t = t; A.h_abs = h_abs; A.t_bound = t_bound; A.rtol = rtol; A.atol = atol;
  A.t_eval = t_eval; A.n_eval = n_eval; A.Y_eval = Y_eval; A.status = status; A.counters = counters;
  A.max_attempts = max_attempts;
  const DpTab T = make_dp_tab();
  const unsigned need = crb_needed_features(plan, sys);
  unsigned prof = crb_pick_profile(need);
  if (prof == CRB_F_PROFILE_C) prof = CRB_F_ALL;  // the reduced-vector feedback path is compiled into the generic RK45 kernel only
  // shared memory: [mass factors][per-member scratch][stage accelerations (+ committed state with CRB_RK45_QVS) 3 m doubles per vector and thread]
  auto total_bytes = [&](const SmemLayout& L, int m) {
    return sizeof(double) * ((size_t)L.mfac_doubles + (size_t)L.scratch_doubles * mpb + (size_t)3 * CRB_RK45_STAGE_VECTORS * m * 32 * WPB) + pad_bytes;
  };
#define CRB_RK45_LAUNCH(MM, PROF, LL, PMV)                                                                 \
  {                                                                                                        \
    const size_t total = total_bytes(SL, MM);                                                              \
    if (int rc = set_smem(crb_rk45_kernel<MM, PROF, LL, WPB, PMV>, total, "crb_rk45")) return rc;          \
    /* the driver's default carve-out (168 KB here) holds 4 blocks of the 64-element shape; the registers allow 5 */ \
    cudaFuncSetAttribute(crb_rk45_kernel<MM, PROF, LL, WPB, PMV>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
    crb_rk45_kernel<MM, PROF, LL, WPB, PMV><<<grid, 32 * WPB, total, stream>>>(P, *sys, SL, A, T);         \
    return 0;                                                                                              \
  }
#define CRB_RK45_CASE(MM, LL)                                                                             \
  if (plan->m == MM && plan->levels == LL && prof != CRB_F_ALL) {                                          \
    if (sys->mass_shared) {                                                                                \
      if (prof == CRB_F_PROFILE_A || CRB_RK45_UMS_ALL) SL = smem_layout_compact(plan, sys, &bytes);        \
      if (prof == CRB_F_PROFILE_A) CRB_RK45_LAUNCH(MM, CRB_F_PROFILE_A, LL, false)                         \
      else CRB_RK45_LAUNCH(MM, CRB_F_PROFILE_B, LL, false)                                                 \
    } else {                                                                                               \
      const SmemLayout keep = SL;                                                                          \
      SL = smem_layout(plan, sys, &bytes);                                                                 \
      SL.mfac_doubles = crb_compact_doubles(plan->m, plan->g, plan->levels) * mpb;                         \
      if (total_bytes(SL, MM) <= 200 * 1024) {                                                             \
        if (prof == CRB_F_PROFILE_A) CRB_RK45_LAUNCH(MM, CRB_F_PROFILE_A, LL, true)                        \
        else CRB_RK45_LAUNCH(MM, CRB_F_PROFILE_B, LL, true)                                                \
      }                                                                                                    \
      SL = keep;                                                                                           \
    }                                                                                                      \
  }
  CRB_SPECIALISED_SHAPES(CRB_RK45_CASE)
#undef CRB_RK45_CASE
#undef CRB_RK45_LAUNCH
  CRB_DISPATCH_M(plan->m, {
    const size_t total = total_bytes(SL, M);
    if (int rc = set_smem(crb_rk45_kernel<M, CRB_F_ALL, -1, WPB>, total, "crb_rk45")) return rc;
    crb_rk45_kernel<M, CRB_F_ALL, -1, WPB><<<grid, 32 * WPB, total, stream>>>(P, *sys, SL, A, T);
  });
  return 0;
}
