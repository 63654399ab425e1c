// crb_rk4_fast_launch.cuh -- launcher template of the fast RK4 family, shared by crb_rk4_fast.cu (one shared mass
// factor set) and crb_rk4_fast_pm.cu (per-member factor sets): the two sets of instantiations compile in parallel.
#pragma once
#include "crb_internal.h"
#include "crb_rk4_fast.cuh"

#include <algorithm>
#include <cstdlib>

template <int M, int LV, bool PM>
static int launch(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                  double* Y_out, int save_every, cudaStream_t stream) {
  const int mpb = CRB_FAST_WARPS * (32 >> LV);
  // compact factor copy: ONE shared set, or (PM) one region per member of the block; behind it the staging rows of
  // recorded full-state frames (FrameWriter), when Y_out rows are 16-byte aligned
  const int fac_doubles = ((crb_compact_doubles(plan->m, plan->g, plan->levels) * (PM ? mpb : 1)) + 1) & ~1;
  // per warp: staging rows of full frames [mpw][2n], or the column table of a lean recording [6M][32] int32
  const bool stage = Y_out && (sys->out_sel_inv || ((uintptr_t)Y_out & 15) == 0);
  const int stage_off = stage ? fac_doubles : 0;
  const int stage_stride = (std::max((32 >> LV) * 2 * plan->n_free, 96 * M) + 1) & ~1;
  const size_t bytes = sizeof(double) * ((size_t)fac_doubles + (stage ? (size_t)CRB_FAST_WARPS * stage_stride : 0));
  const int grid = (sys->n_members + mpb - 1) / mpb;
  KPlan P = kplan_of(plan);
  P.n_sm = crb_sm_count();
  const bool uc = sys->u_const || sys->f_ext, imp = sys->imp_amp != nullptr;
  // shared mass factors: persistent paired kernel (bulk-copy tiles; needs 16-byte aligned rows); per-member factors
  // (PM) and lean-recording requests it cannot serve keep the one-tile-per-block kernel
  const bool bulk_ok = !PM && (((uintptr_t)X | (uintptr_t)(sys->out_sel_inv ? nullptr : Y_out) | (uintptr_t)sys->kcoef) & 15) == 0;
#define CRB_LIN2N(UCV, IMPV, NCV)                                                                             \
  {                                                                                                            \
    if constexpr (!PM) {                                                                                       \
      if (bulk_ok) {                                                                                           \
        typedef FastTileGeom<M, LV> TG;                                                                        \
        const size_t pbytes = TG::smem_bytes(CRB_PERSIST_WARPS);                                               \
        const int n_tiles = (sys->n_members + TG::mpw - 1) / TG::mpw;                                          \
        const int pgrid = std::min((n_tiles + CRB_PERSIST_WARPS - 1) / CRB_PERSIST_WARPS,                      \
                                   P.n_sm * CRB_FAST_MINBLOCKS * CRB_FAST_WARPS / CRB_PERSIST_WARPS);          \
        if (int rc = set_smem(crb_rk4_lin2p_kernel<M, LV, UCV, IMPV, NCV>, pbytes, "crb_rk4")) return rc;      \
        if (sys->tile_counter && n_tiles > pgrid)                                                              \
          if (cudaMemsetAsync(sys->tile_counter, 0, sizeof(int32_t), stream) != cudaSuccess)                   \
            return crb_fail(CRB_E_CUDA, "crb_rk4: cannot reset the tile counter");                             \
        crb_rk4_lin2p_kernel<M, LV, UCV, IMPV, NCV><<<pgrid, 32 * CRB_PERSIST_WARPS, pbytes, stream>>>(        \
            P, *sys, X, t0, h, nsteps, Y_out, save_every, n_tiles > pgrid ? sys->tile_counter : nullptr);      \
        return 0;                                                                                              \
      }                                                                                                        \
    }                                                                                                          \
    if constexpr (PM) {                                                                                        \
      if (Y_out) {                                                                                             \
        if (int rc = set_smem(crb_rk4_lin2_kernel<M, LV, UCV, IMPV, PM, NCV, true>, bytes, "crb_rk4")) return rc; \
        crb_rk4_lin2_kernel<M, LV, UCV, IMPV, PM, NCV, true><<<grid, CRB_FAST_THREADS, bytes, stream>>>(       \
            P, *sys, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);                            \
      } else {                                                                                                 \
        if (int rc = set_smem(crb_rk4_lin2_kernel<M, LV, UCV, IMPV, PM, NCV, false>, bytes, "crb_rk4")) return rc; \
        crb_rk4_lin2_kernel<M, LV, UCV, IMPV, PM, NCV, false><<<grid, CRB_FAST_THREADS, bytes, stream>>>(      \
            P, *sys, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);                            \
      }                                                                                                        \
    } else {                                                                                                   \
      return 1; /* rows not 16-byte aligned: general kernel */                                                 \
    }                                                                                                          \
  }
#define CRB_LIN2(UCV, IMPV) CRB_LIN2N(UCV, IMPV, false)
  const bool nc = !(plan->contiguous && plan->p_act == plan->p);
  if (sys->grav_mode == 1) {  // slot-space gravity: stage-by-stage kernel (the force is nonlinear in the rotations)
    if (PM || uc) return 1;
    if constexpr (!PM) {
#define CRB_FASTG(IMPV, NCV)                                                                                  \
  {                                                                                                           \
    if (Y_out) {                                                                                              \
      if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, IMPV, true, NCV, true>, bytes, "crb_rk4")) return rc;  \
      crb_rk4_fast_kernel<M, LV, IMPV, true, NCV, true><<<grid, CRB_FAST_THREADS, bytes, stream>>>(           \
          P, *sys, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);                             \
    } else {                                                                                                  \
      if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, IMPV, true, NCV, false>, bytes, "crb_rk4")) return rc; \
      crb_rk4_fast_kernel<M, LV, IMPV, true, NCV, false><<<grid, CRB_FAST_THREADS, bytes, stream>>>(          \
          P, *sys, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);                             \
    }                                                                                                         \
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
    if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, true, false, false, true>, bytes, "crb_rk4")) return rc;
    crb_rk4_fast_kernel<M, LV, true, false, false, true><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);
  } else {
    if (int rc = set_smem(crb_rk4_fast_kernel<M, LV, false, false, false, true>, bytes, "crb_rk4")) return rc;
    crb_rk4_fast_kernel<M, LV, false, false, false, true><<<grid, CRB_FAST_THREADS, bytes, stream>>>(P, *sys, X, t0, h, nsteps, Y_out, save_every, stage_off, stage_stride);
  }
#undef CRB_LIN2
#undef CRB_LIN2N
  return 0;
}

