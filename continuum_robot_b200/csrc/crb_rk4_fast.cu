// crb_rk4_fast.cu -- launcher of the linear / uniform-mass fused RK4 kernel.
#include "crb_internal.h"
#include "crb_rk4_fast.cuh"

int crb_launch_rk4_fast(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                        int nsteps, double* Y_out, int save_every, cudaStream_t stream) {
  size_t bytes;
  const SmemLayout SL = smem_layout(plan, sys, &bytes);
  (void)SL;
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  UniformMass um = {sys->um[0], sys->um[1], sys->um[2], sys->um[3]};
  CRB_DISPATCH_M(plan->m, {
    if (int rc = set_smem(crb_rk4_fast_kernel<M>, bytes, "crb_rk4")) return rc;
    crb_rk4_fast_kernel<M><<<grid, CRB_THREADS, bytes, stream>>>(P, *sys, um, X, t0, h, nsteps, Y_out, save_every);
  });
  return 0;
}
