// crb_internal.h -- host-side helpers shared by the translation units of libcrb.so.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "crb.h"
#include "crb_device.cuh"

#define CRB_E_ARG (-1)
#define CRB_E_CUDA (-2)
#define CRB_E_LIMIT (-3)

int crb_fail(int code, const char* fmt, ...);

// inputs that change within a step (sinusoid / table): only the general RHS evaluates them
inline bool crb_time_varying_input(const crb_system_t* s) { return s->u_sin_amp || s->u_tab_v; }  // sets the thread-local error text, returns code

inline KPlan kplan_of(const crb_plan_t* p) {
  KPlan k;
  k.N = p->n_elements;
  k.n_free = p->n_free;
  k.n0 = p->n0;
  k.p_act = p->p_act;
  k.m = p->m;
  k.g = p->g;
  k.p = p->p;
  k.levels = p->levels;
  k.contiguous = p->contiguous;
  k.has_mask = p->has_mask;
  k.mfac_doubles = p->mfac_doubles;
  k.n_sm = 0;
  return k;
}

inline SmemLayout smem_layout(const crb_plan_t* plan, const crb_system_t* sys, size_t* bytes) {
  SmemLayout SL;
  SL.mfac_doubles = sys->mass_shared ? 2 * CRB_SLOT_PAIRS * plan->p + 2 * CRB_SCAN_PAIRS * (plan->levels > 0 ? plan->levels : 1) * plan->g : 0;
  SL.scratch_doubles = ((sys->gain && !(sys->gain_frag && plan->g == 4 && sys->gain_stride == 0)) || sys->grav_mode == 2)
                           ? 2 * plan->n_free + 2 * plan->n_elements : 0;  // reduced state + per-segment gravity
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  *bytes = sizeof(double) * ((size_t)SL.mfac_doubles + (size_t)SL.scratch_doubles * mpb);
  return SL;
}

// shape-specialised kernels stage only the compact copy (Sinv + scan products) of the shared factor set
inline SmemLayout smem_layout_compact(const crb_plan_t* plan, const crb_system_t* sys, size_t* bytes) {
  SmemLayout SL = smem_layout(plan, sys, bytes);
  SL.mfac_doubles = crb_compact_doubles(plan->m, plan->g, plan->levels);
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  *bytes = sizeof(double) * ((size_t)SL.mfac_doubles + (size_t)SL.scratch_doubles * mpb);
  return SL;
}
// ... or, with per-member mass, one compact copy per member of the block
inline SmemLayout smem_layout_compact_pm(const crb_plan_t* plan, const crb_system_t* sys, size_t* bytes) {
  SmemLayout SL = smem_layout(plan, sys, bytes);
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  SL.mfac_doubles = crb_compact_doubles(plan->m, plan->g, plan->levels) * mpb;
  *bytes = sizeof(double) * ((size_t)SL.mfac_doubles + (size_t)SL.scratch_doubles * mpb);
  return SL;
}

// SM count of the current device (queried per call: a cheap attribute read, no global state)
inline int crb_sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
    return 148;
  return sms;
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* who) {
  if (bytes > 48 * 1024) {
    if (bytes > 227 * 1024) return crb_fail(CRB_E_LIMIT, "%s: needs %zu bytes of shared memory (> 227 KB)", who, bytes);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
  }
  return 0;
}

#define CRB_DISPATCH_M(mval, ...)                                    \
  switch (mval) {                                                    \
    case 1: { constexpr int M = 1; __VA_ARGS__; } break;             \
    case 2: { constexpr int M = 2; __VA_ARGS__; } break;             \
    case 3: { constexpr int M = 3; __VA_ARGS__; } break;             \
    case 4: { constexpr int M = 4; __VA_ARGS__; } break;             \
    default: return crb_fail(CRB_E_LIMIT, "unsupported slots per lane %d", mval); \
  }

// feature bits a system needs (see CRB_F_* in crb_device.cuh) and the smallest compiled profile
inline unsigned crb_needed_features(const crb_plan_t* plan, const crb_system_t* s) {
  unsigned f = 0;
  f |= s->all_linear ? CRB_F_LINEAR : (CRB_F_LINEAR | CRB_F_NONLIN);
  if (s->all_nonlinear) f = (f & ~CRB_F_LINEAR) | CRB_F_NONLIN;
  if (s->drag) f |= CRB_F_DRAG;
  if (s->grav_mode == 1) f |= CRB_F_GRAVS;
  if (s->grav_mode == 2) f |= CRB_F_GRAVG;
  if (plan->has_mask) f |= CRB_F_MASK;
  if (s->u_const || s->imp_amp || s->f_ext || s->u_sin_amp || s->u_tab_v) f |= CRB_F_INPUT;
  if (s->gain) f |= (s->gain_frag && plan->g == 4 && s->gain_stride == 0) ? CRB_F_GAINM : CRB_F_GAINS;
  return f;
}
inline unsigned crb_pick_profile(unsigned need) {
  if ((need & ~CRB_F_PROFILE_A) == 0) return CRB_F_PROFILE_A;
  if ((need & ~CRB_F_PROFILE_C) == 0) return CRB_F_PROFILE_C;
  if ((need & ~CRB_F_PROFILE_B) == 0) return CRB_F_PROFILE_B;
  return CRB_F_ALL;
}
// (slots per lane, log2 lanes per member) shapes that get fully specialised profile kernels:
// N = 20 (3,3); N = 64 (4,4) and (2,5); N = 6 (2,2) and (3,1); N = 10 (3,2); N = 32 (4,3) and (2,4)
#define CRB_SPECIALISED_SHAPES(X) X(3, 3) X(4, 4) X(2, 5) X(2, 2) X(3, 1) X(3, 2) X(4, 3) X(2, 4)

#define CRB_DISPATCH_PROFILE(need, ...)                                          \
  switch (crb_pick_profile(need)) {                                              \
    case CRB_F_PROFILE_A: { constexpr unsigned FEAT = CRB_F_PROFILE_A; __VA_ARGS__; } break; \
    case CRB_F_PROFILE_B: { constexpr unsigned FEAT = CRB_F_PROFILE_B; __VA_ARGS__; } break; \
    default: { constexpr unsigned FEAT = CRB_F_ALL; __VA_ARGS__; } break;        \
  }

int crb_fast_members_per_sm(int members_per_warp);  // resident members per SM of the fast RK4 family (crb_rk4_fast.cu)
bool crb_shared_eligible(const crb_plan_t* plan, const crb_system_t* sys);  // crb_shared.cu
bool crb_dense_eligible(const crb_plan_t* plan, const crb_system_t* sys);   // crb_rk4_dense.cu
int crb_launch_rk4_dense(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                         double* Y_out, int save_every, cudaStream_t stream);
int crb_launch_rk4_shared(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                          double* Y_out, int save_every, cudaStream_t stream);
// launchers implemented in their own translation units (parallel nvcc builds)
int crb_launch_rk4_general(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                           int nsteps, double* Y_out, int save_every, cudaStream_t stream);
int crb_launch_rk4_fast(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                        int nsteps, double* Y_out, int save_every, cudaStream_t stream);
int crb_launch_midpoint(const crb_plan_t* plan, const crb_system_t* sys, const double* afac, int afac_shared, double* X,
                        double t0, double h, int nsteps, double* Y_out, int save_every, cudaStream_t stream);
int crb_launch_rk45(const crb_plan_t* plan, const crb_system_t* sys, double* X, double* t, double* h_abs,
                    double t_bound, double rtol, double atol, const double* t_eval, int n_eval, double* Y_eval,
                    int* status, long long* counters, int max_attempts, cudaStream_t stream);
