// crb_api.cu -- kernels + C ABI of libcrb.so (see include/crb.h).  sm_100a only.
#include <algorithm>
#include <vector>

#include "crb_internal.h"
#include "crb_assemble.cuh"

// ------------------------------------------------------------------------------------------
// error handling (thread-local string; no other global mutable state)
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int crb_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define fail crb_fail

extern "C" int crb_version(void) { return CRB_VERSION; }
extern "C" int crb_abi_sizes(int32_t* plan_bytes, int32_t* system_bytes) {
  if (!plan_bytes || !system_bytes) return fail(CRB_E_ARG, "crb_abi_sizes: null argument");
  *plan_bytes = (int32_t)sizeof(crb_plan_t);
  *system_bytes = (int32_t)sizeof(crb_system_t);
  return 0;
}
extern "C" const char* crb_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// crb_plan (host)
// ------------------------------------------------------------------------------------------
extern "C" int crb_plan(int32_t n_elements, const uint8_t* bc, int32_t max_slots_per_lane, crb_plan_t* out) {
  if (!bc || !out) return fail(CRB_E_ARG, "crb_plan: null argument");
  if (n_elements < 1) return fail(CRB_E_ARG, "crb_plan: n_elements must be >= 1, got %d", n_elements);
  const int N = n_elements;
  for (int i = 0; i <= N; ++i)
    if (bc[i] > CRB_BC_PINNED) return fail(CRB_E_ARG, "crb_plan: bad boundary condition code %d at node %d", bc[i], i);
  memset(out, 0, sizeof(*out));
  const int mmax = max_slots_per_lane > 0 ? (max_slots_per_lane > 4 ? 4 : max_slots_per_lane) : 4;
  const int n0 = bc[0] == CRB_BC_FIXED ? 1 : 0;
  const int p_act = N + 1 - n0;
  int g = 1;
  while ((p_act + g - 1) / g > mmax && g < 32) g *= 2;
  const int m = (p_act + g - 1) / g;
  if (m > 4 || m * g > CRB_MAX_SLOTS)
    return fail(CRB_E_LIMIT, "crb_plan: %d active nodes exceed the %d-slot limit of one lane group", p_act, 32 * 4);
  int levels = 0;
  while ((1 << levels) < g) ++levels;
  out->n_elements = N;
  out->n0 = n0;
  out->p_act = p_act;
  out->m = m;
  out->g = g;
  out->p = m * g;
  out->levels = levels;
  int r = 0, contiguous = 1;
  for (int s = 0; s < out->p; ++s) {
    for (int d = 0; d < 3; ++d) {
      int idx = -1;
      if (s < p_act) {
        const int node = s + n0;
        const bool constrained = bc[node] == CRB_BC_FIXED || (bc[node] == CRB_BC_PINNED && d < 2);
        if (!constrained) idx = r++;
        else contiguous = 0;
      }
      out->red_index[3 * s + d] = idx;
    }
  }
  if (r == 0) return fail(CRB_E_ARG, "crb_plan: cannot constrain all degrees of freedom");
  out->n_free = r;
  out->contiguous = contiguous;
  out->has_mask = !contiguous;
  out->mfac_doubles = 2ll * CRB_SLOT_PAIRS * out->p + 2ll * CRB_SCAN_PAIRS * (levels > 0 ? levels : 1) * g +
                      crb_compact_doubles(out->m, g, levels);  // + compact copy (crb_device.cuh::fast_solve_r)
  out->kcoef_doubles = 4ll * out->p;
  return 0;
}

static int fill_topo(const crb_plan_t* plan, const uint8_t* elem_type_host, AsmTopo* T) {
  memset(T, 0, sizeof(*T));
  for (int s = 0; s < plan->p; ++s) {
    uint8_t fb = 0;
    for (int d = 0; d < 3; ++d)
      if (plan->red_index[3 * s + d] >= 0) fb |= (1u << d);
    T->free_bits[s] = fb;
    const int e = plan->n0 + s - 1;
    uint8_t et = CRB_ELEM_ABSENT;
    if (s < plan->p_act && e >= 0 && e < plan->n_elements) {
      et = elem_type_host[e];
      if (et > CRB_ELEM_NONLINEAR) return fail(CRB_E_ARG, "crb_assemble: bad element type %d at element %d", et, e);
    }
    T->etype[s] = et;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// crb_assemble
// ------------------------------------------------------------------------------------------
extern "C" int crb_assemble(const crb_plan_t* plan, const double* params, int32_t n_param_sets,
                            const uint8_t* elem_type_host, const uint8_t* bc_host, int32_t n_mass,
                            int32_t n_stiff, int32_t n_force, double fluid_density, double* mfac,
                            double* kcoef, uint8_t* elem_type_slots, double* drag, double* grav,
                            double* seg_half_mass, void* stream) {
  (void)bc_host;
  if (!plan || !params || !elem_type_host || !mfac || !kcoef || !elem_type_slots)
    return fail(CRB_E_ARG, "crb_assemble: null argument");
  if (n_param_sets < 1) return fail(CRB_E_ARG, "crb_assemble: n_param_sets must be >= 1");
  for (int v : {n_mass, n_stiff, n_force})
    if (v != 1 && v != n_param_sets) return fail(CRB_E_ARG, "crb_assemble: set counts must be 1 or n_param_sets");
  AsmTopo T;
  if (int rc = fill_topo(plan, elem_type_host, &T)) return rc;
  const int nthreads = std::max(std::max(n_mass, n_stiff), n_force);
  const int block = 128;
  const int grid = (nthreads + block - 1) / block;
  crb_assemble_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(kplan_of(plan), T, params, n_param_sets, n_mass,
                                                               n_stiff, n_force, fluid_density, 0.0, mfac, kcoef,
                                                               elem_type_slots, drag, grav, seg_half_mass);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_assemble: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int crb_assemble_shifted(const crb_plan_t* plan, const double* params, int32_t n_param_sets,
                                    const uint8_t* elem_type_host, const uint8_t* bc_host, int32_t n_sets,
                                    double shift, double* afac, void* stream) {
  (void)bc_host;
  if (!plan || !params || !elem_type_host || !afac) return fail(CRB_E_ARG, "crb_assemble_shifted: null argument");
  if (n_param_sets < 1) return fail(CRB_E_ARG, "crb_assemble_shifted: n_param_sets must be >= 1");
  if (n_sets != 1 && n_sets != n_param_sets) return fail(CRB_E_ARG, "crb_assemble_shifted: n_sets must be 1 or n_param_sets");
  if (!(shift >= 0.0) || !std::isfinite(shift)) return fail(CRB_E_ARG, "crb_assemble_shifted: shift must be finite and >= 0");
  for (int e = 0; e < plan->n_elements; ++e)
    if (elem_type_host[e] != CRB_ELEM_LINEAR)
      return fail(CRB_E_ARG, "crb_assemble_shifted: element %d is not linear (M + shift K needs a stiffness matrix)", e);
  AsmTopo T;
  if (int rc = fill_topo(plan, elem_type_host, &T)) return rc;
  const int block = 128;
  const int grid = (n_sets + block - 1) / block;
  crb_assemble_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(kplan_of(plan), T, params, n_param_sets, n_sets, 0, 0, 0.0,
                                                               shift, afac, nullptr, nullptr, nullptr, nullptr, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_assemble_shifted: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------
// kernels: RHS and fused RK4
// ------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(CRB_THREADS)
crb_rhs_kernel(KPlan P, crb_system_t S, SmemLayout SL, const double* __restrict__ X, double t,
               double* __restrict__ dX) {
  extern __shared__ __align__(16) double smem[];
  const double* mf = stage_mfac(S, P, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mpw = 32 / P.g;
  const int mloc = warp * mpw + lane / P.g;
  const int member = blockIdx.x * (CRB_WARPS_PER_BLOCK * mpw) + mloc;
  LaneCtx<M> L;
  load_lane_ctx<M>(L, P, S, member, lane % P.g, mf,
                   SL.scratch_doubles ? smem + SL.mfac_doubles + mloc * SL.scratch_doubles : nullptr);
  const RhsFlags F = make_flags(S, P);
  double q[M][3], v[M][3], a[M][3];
  load_state<M>(L, X, q, v);
  beam_accel<M>(L, S, F, q, v, t, a);
  store_state<M>(L, dX, v, a);
}

// built-in force vector f(x)[B, n] (drag + gravity), no stiffness / inputs / solve
template <int M>
__global__ void __launch_bounds__(CRB_THREADS)
crb_forces_kernel(KPlan P, crb_system_t S, SmemLayout SL, const double* __restrict__ X, double* __restrict__ Fout) {
  extern __shared__ __align__(16) double smem[];
  const double* mf = stage_mfac(S, P, smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mpw = 32 / P.g;
  const int mloc = warp * mpw + lane / P.g;
  const int member = blockIdx.x * (CRB_WARPS_PER_BLOCK * mpw) + mloc;
  LaneCtx<M> L;
  load_lane_ctx<M>(L, P, S, member, lane % P.g, mf,
                   SL.scratch_doubles ? smem + SL.mfac_doubles + mloc * SL.scratch_doubles : nullptr);
  RhsFlags F = make_flags(S, P);
  F.gain = false;
  double q[M][3], v[M][3], a[M][3];
  load_state<M>(L, X, q, v);
  beam_accel<M, CRB_F_ALL, true>(L, S, F, q, v, 0.0, a);
  if (!L.active) return;
  double* f = Fout + (long long)L.member * L.n;
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (L.ri[j][d] >= 0) f[L.ri[j][d]] = a[j][d];
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
static int check_system(const char* who, const crb_plan_t* plan, const crb_system_t* sys) {
  if (!plan || !sys) return fail(CRB_E_ARG, "%s: null plan/system", who);
  if (sys->n_members < 1) return fail(CRB_E_ARG, "%s: n_members must be >= 1, got %d", who, sys->n_members);
  if (!sys->mfac || !sys->kcoef || !sys->elem_type || !sys->red_index)
    return fail(CRB_E_ARG, "%s: system not assembled (mfac/kcoef/elem_type/red_index missing)", who);
  if (sys->grav_mode < 0 || sys->grav_mode > 2) return fail(CRB_E_ARG, "%s: bad grav_mode %d", who, sys->grav_mode);
  if (sys->grav_mode == 1 && (!sys->grav || !plan->contiguous))
    return fail(CRB_E_ARG, "%s: slot-space gravity needs grav[] and a contiguous plan", who);
  if (sys->grav_mode == 2 && !sys->seg_half_mass)
    return fail(CRB_E_ARG, "%s: generic gravity needs seg_half_mass[]", who);
  if (sys->gain_stride != 0 && sys->gain_stride != 2ll * plan->n_free * plan->n_free)
    return fail(CRB_E_ARG, "%s: gain_stride must be 0 (shared gain) or n*2n = %lld, got %lld", who,
                2ll * plan->n_free * plan->n_free, (long long)sys->gain_stride);
  if (sys->member_op && (!sys->gain || sys->gain_stride == 0))
    return fail(CRB_E_ARG, "%s: member_op needs the per-member gains it was built from (gain, gain_stride)", who);
  if (sys->gain_stride != 0 && (sys->gain_frag || sys->shared_op))
    return fail(CRB_E_ARG, "%s: per-member gains (gain_stride != 0) exclude gain_frag / shared_op", who);
  if (sys->imp_amp && (sys->imp_dof < 0 || sys->imp_dof >= plan->n_free))
    return fail(CRB_E_ARG, "%s: imp_dof %d outside [0,%d)", who, sys->imp_dof, plan->n_free);
  if (sys->out_sel_inv && (sys->out_n_sel < 1 || sys->out_n_sel > 2 * plan->n_free))
    return fail(CRB_E_ARG, "%s: out_n_sel must be in [1, 2n = %d] when out_sel_inv is set, got %d", who, 2 * plan->n_free,
                sys->out_n_sel);
  if (sys->u_tab_v && (!sys->u_tab_t || sys->u_tab_k < 2))
    return fail(CRB_E_ARG, "%s: the input table needs its knots u_tab_t and u_tab_k >= 2 (got %d)", who, sys->u_tab_k);
  return 0;
}

extern "C" int crb_rhs(const crb_plan_t* plan, const crb_system_t* sys, const double* X, double t, double* dX,
                       void* stream) {
  if (int rc = check_system("crb_rhs", plan, sys)) return rc;
  if (!X || !dX) return fail(CRB_E_ARG, "crb_rhs: null state pointer");
  size_t bytes;
  const SmemLayout SL = smem_layout(plan, sys, &bytes);
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  CRB_DISPATCH_M(plan->m, {
    if (int rc = set_smem(crb_rhs_kernel<M>, bytes, "crb_rhs")) return rc;
    crb_rhs_kernel<M><<<grid, CRB_THREADS, bytes, (cudaStream_t)stream>>>(P, *sys, SL, X, t, dX);
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_rhs: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int crb_forces(const crb_plan_t* plan, const crb_system_t* sys, const double* X, double* F_out,
                          void* stream) {
  if (int rc = check_system("crb_forces", plan, sys)) return rc;
  if (!X || !F_out) return fail(CRB_E_ARG, "crb_forces: null pointer");
  size_t bytes;
  crb_system_t s2 = *sys;
  s2.gain = nullptr;
  const SmemLayout SL = smem_layout(plan, &s2, &bytes);
  const int mpb = CRB_WARPS_PER_BLOCK * (32 / plan->g);
  const int grid = (sys->n_members + mpb - 1) / mpb;
  const KPlan P = kplan_of(plan);
  CRB_DISPATCH_M(plan->m, {
    if (int rc = set_smem(crb_forces_kernel<M>, bytes, "crb_forces")) return rc;
    crb_forces_kernel<M><<<grid, CRB_THREADS, bytes, (cudaStream_t)stream>>>(P, s2, SL, X, F_out);
  });
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_forces: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// fast path: all-linear, no drag, no feedback and no input other than a constant force / tip impulse (BASELINE
// config 3 shape); any mass distribution, shared or per member; any boundary conditions (plans with constrained
// DOFs in active slots or phantom slots take the NC variants of the paired kernel); slot-space gravity (config 1
// as an ensemble) on the stage-by-stage kernel
static bool rk4_fast_eligible(const crb_plan_t* plan, const crb_system_t* sys) {
  return sys->all_linear && !sys->drag && sys->grav_mode != 2 && !sys->gain && !sys->force_general &&
         !crb_time_varying_input(sys);
}

extern "C" int crb_rk4_wave_members(const crb_plan_t* plan, const crb_system_t* sys, int32_t* out) {
  if (!plan || !sys || !out) return fail(CRB_E_ARG, "crb_rk4_wave_members: null argument");
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return fail(CRB_E_CUDA, "crb_rk4_wave_members: cannot query the device");
  const int mpw = 32 / plan->g;
  // resident blocks per SM follow the register allocation of the kernel families (launch bounds)
  const int per_sm = rk4_fast_eligible(plan, sys) ? crb_fast_members_per_sm(mpw) : 2 * CRB_WARPS_PER_BLOCK * mpw;
  *out = sms * per_sm;
  return 0;
}

static int rk4_dispatch(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int32_t nsteps,
                        double* Y_out, int32_t save_every, cudaStream_t stream) {
  int rc;
  if (crb_shared_eligible(plan, sys)) {  // one design + gain shared by all members: dense tensor-core contraction
    rc = crb_launch_rk4_shared(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
  } else if (crb_dense_eligible(plan, sys)) {  // linear designs with one gain per member: dense operator per member
    rc = crb_launch_rk4_dense(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
  } else {
    rc = rk4_fast_eligible(plan, sys) ? crb_launch_rk4_fast(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream) : 1;
    if (rc == 1)  // not eligible, or shape not instantiated in the fast family
      rc = crb_launch_rk4_general(plan, sys, X, t0, h, nsteps, Y_out, save_every, stream);
  }
  if (rc) return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_rk4: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int crb_rk4(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
                       int32_t nsteps, double* Y_out, int32_t save_every, void* stream) {
  if (int rc = check_system("crb_rk4", plan, sys)) return rc;
  if (!X) return fail(CRB_E_ARG, "crb_rk4: null state pointer");
  if (nsteps < 0) return fail(CRB_E_ARG, "crb_rk4: nsteps must be >= 0");
  if (!(h > 0.0) || !std::isfinite(h)) return fail(CRB_E_ARG, "crb_rk4: step h must be positive and finite");
  if (Y_out && save_every < 1) return fail(CRB_E_ARG, "crb_rk4: save_every must be >= 1 when Y_out is given");
  if (nsteps == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (sys->imp_amp) {
    // The impulse (examples/example_utilities.py:144-148: a force while t < duration) is the only
    // time-dependent term.  Steps that start at or after the end of the window see no input at any
    // stage, so they run as the input-free system (cheaper kernel variants, e.g. the force-free paired
    // kernel).  k_off = first such step, plus one step of margin so that the gate is never decided
    // differently by host and device rounding of t0 + k h.
    long long k_off = 0;
    if (t0 < sys->imp_duration) {
      k_off = (long long)std::ceil((sys->imp_duration - t0) / h);
      while (k_off > 0 && t0 + (double)(k_off - 1) * h >= sys->imp_duration) --k_off;
      while (t0 + (double)k_off * h < sys->imp_duration) ++k_off;
      ++k_off;
    }
    if (Y_out) k_off = (k_off + save_every - 1) / save_every * save_every;  // split on a frame boundary (later is always valid)
    if (k_off < nsteps) {
      crb_system_t quiet = *sys;
      quiet.imp_amp = nullptr;
      if (k_off > 0)
        if (int rc = rk4_dispatch(plan, sys, X, t0, h, (int32_t)k_off, Y_out, save_every, st)) return rc;
      const long long width = sys->out_sel_inv ? sys->out_n_sel : 2 * plan->n_free;
      double* Y2 = Y_out ? Y_out + (k_off / save_every) * (long long)sys->n_members * width : nullptr;
      return rk4_dispatch(plan, &quiet, X, t0 + (double)k_off * h, h, nsteps - (int32_t)k_off, Y2, save_every, st);
    }
  }
  return rk4_dispatch(plan, sys, X, t0, h, nsteps, Y_out, save_every, st);
}

extern "C" int crb_midpoint(const crb_plan_t* plan, const crb_system_t* sys, const double* afac, int32_t afac_shared,
                            double* X, double t0, double h, int32_t nsteps, double* Y_out, int32_t save_every,
                            void* stream) {
  if (int rc = check_system("crb_midpoint", plan, sys)) return rc;
  if (!X || !afac) return fail(CRB_E_ARG, "crb_midpoint: null state / factor pointer");
  if (nsteps < 0) return fail(CRB_E_ARG, "crb_midpoint: nsteps must be >= 0");
  if (!(h > 0.0) || !std::isfinite(h)) return fail(CRB_E_ARG, "crb_midpoint: step h must be positive and finite");
  if (Y_out && save_every < 1) return fail(CRB_E_ARG, "crb_midpoint: save_every must be >= 1 when Y_out is given");
  if (!(sys->all_linear && !sys->drag && sys->grav_mode == 0 && !sys->gain && !crb_time_varying_input(sys)))
    return fail(CRB_E_ARG, "crb_midpoint: needs an all-linear beam without drag / gravity / feedback "
                           "(the implicit step solves with M + h^2/4 K)");
  if (nsteps == 0) return 0;
  if (int rc = crb_launch_midpoint(plan, sys, afac, afac_shared, X, t0, h, nsteps, Y_out, save_every, (cudaStream_t)stream))
    return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_midpoint: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------
// crb_rk45
// ------------------------------------------------------------------------------------------
extern "C" int crb_rk45(const crb_plan_t* plan, const crb_system_t* sys, double* X, double* t, double* h_abs,
                        double t_bound, double rtol, double atol, const double* t_eval, int32_t n_eval,
                        double* Y_eval, int32_t* status, int64_t* counters, int32_t max_attempts, void* stream) {
  if (int rc = check_system("crb_rk45", plan, sys)) return rc;
  if (!X || !t || !h_abs || !status || !counters) return fail(CRB_E_ARG, "crb_rk45: null argument");
  if (n_eval < 0 || (n_eval > 0 && (!t_eval || !Y_eval))) return fail(CRB_E_ARG, "crb_rk45: bad t_eval / Y_eval");
  if (!(rtol > 0.0) || !(atol >= 0.0)) return fail(CRB_E_ARG, "crb_rk45: rtol must be > 0 and atol >= 0");
  if (max_attempts < 1) return fail(CRB_E_ARG, "crb_rk45: max_attempts must be >= 1");
  if (int rc = crb_launch_rk45(plan, sys, X, t, h_abs, t_bound, rtol, atol, t_eval, n_eval, Y_eval, status,
                               reinterpret_cast<long long*>(counters), max_attempts, (cudaStream_t)stream))
    return rc;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(CRB_E_CUDA, "crb_rk45: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

// ------------------------------------------------------------------------------------------
// crb_gain_fragments (host)
// ------------------------------------------------------------------------------------------
extern "C" int64_t crb_gain_fragments(const crb_plan_t* plan, const double* gain, double* out) {
  if (!plan) return fail(CRB_E_ARG, "crb_gain_fragments: null plan");
  if (plan->g != 4) return fail(CRB_E_ARG, "crb_gain_fragments: needs a plan with 4 lanes per member, got %d", plan->g);
  const int m = plan->m, n = plan->n_free, KT = 6 * m, NT = (3 * m + 1) / 2;
  const int64_t count = (int64_t)KT * NT * 32;
  if (!out) return count;
  if (!gain) return fail(CRB_E_ARG, "crb_gain_fragments: null gain");
  for (int kt = 0; kt < KT; ++kt)
    for (int nt = 0; nt < NT; ++nt)
      for (int lane = 0; lane < 32; ++lane) {
        const int src = lane % 4, col8 = lane / 4, dst = col8 / 2, el = col8 % 2;
        const int oidx = 2 * nt + el;  // own position DOF of the destination lane
        double val = 0.0;
        if (oidx < 3 * m) {
          const int r_out = plan->red_index[3 * (dst * m) + oidx];
          const int kk = kt < 3 * m ? kt : kt - 3 * m;
          const int r_in = plan->red_index[3 * (src * m) + kk];
          if (r_out >= 0 && r_in >= 0) val = gain[(int64_t)r_out * 2 * n + (kt < 3 * m ? r_in : n + r_in)];
        }
        out[((int64_t)kt * NT + nt) * 32 + lane] = val;
      }
  return count;
}

// ------------------------------------------------------------------------------------------
// crb_dense_matrices (host): BC-reduced M and K of one design, for LQR synthesis on the host.
// models/segments.py:32-78, euler_bernoulli_beam.py:139-161, 265, 422-511.
// ------------------------------------------------------------------------------------------
extern "C" int crb_dense_matrices(const crb_plan_t* plan, const double* par, const uint8_t* elem_type_host,
                                  const uint8_t* bc_host, double* M_out, double* K_out) {
  if (!plan || !par || !elem_type_host || !bc_host || !M_out) return fail(CRB_E_ARG, "crb_dense_matrices: null argument");
  const int N = plan->n_elements, nf = 3 * (N + 1), n = plan->n_free;
  if (K_out)
    for (int e = 0; e < N; ++e)
      if (elem_type_host[e] != CRB_ELEM_LINEAR)
        return fail(CRB_E_ARG, "Cannot extract stiffness matrix from beam with nonlinear segments. Segment %d is nonlinear.", e);
  std::vector<int> red(nf, -1);
  int r = 0;
  for (int node = 0; node <= N; ++node)
    for (int d = 0; d < 3; ++d) {
      const bool c = bc_host[node] == CRB_BC_FIXED || (bc_host[node] == CRB_BC_PINNED && d < 2);
      if (!c) red[3 * node + d] = r++;
    }
  if (r != n) return fail(CRB_E_ARG, "crb_dense_matrices: plan / bc mismatch (%d vs %d free DOFs)", r, n);
  for (long long k = 0; k < (long long)n * n; ++k) M_out[k] = 0.0;
  if (K_out)
    for (long long k = 0; k < (long long)n * n; ++k) K_out[k] = 0.0;
  static const double mi[6][6] = {{140, 0, 0, 70, 0, 0},   {0, 156, -22, 0, 54, 13}, {0, -22, 4, 0, -13, -3},
                                  {70, 0, 0, 140, 0, 0},   {0, 54, -13, 0, 156, 22}, {0, 13, -3, 0, 22, 4}};
  static const int mp[6][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 1}, {0, 1, 2, 0, 1, 2},
                               {0, 0, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 1}, {0, 1, 2, 0, 1, 2}};
  for (int e = 0; e < N; ++e) {
    const double* q = par + e * CRB_NPARAM;
    const double L = q[CRB_P_LENGTH], mu = q[CRB_P_RHO] * q[CRB_P_AREA] * L / 420;
    const double EI = q[CRB_P_E] * q[CRB_P_I], EA = q[CRB_P_E] * q[CRB_P_AREA];
    const double a = EA / L, c1 = EI / L, c2 = EI / (L * L), c3 = EI / (L * L * L);
    const double ke[6][6] = {{a, 0, 0, -a, 0, 0},
                             {0, 12 * c3, -6 * c2, 0, -12 * c3, -6 * c2},
                             {0, -6 * c2, 4 * c1, 0, 6 * c2, 2 * c1},
                             {-a, 0, 0, a, 0, 0},
                             {0, -12 * c3, 6 * c2, 0, 12 * c3, 6 * c2},
                             {0, -6 * c2, 2 * c1, 0, 6 * c2, 4 * c1}};
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) {
        const int ri = red[3 * e + i], rj = red[3 * e + j];
        if (ri < 0 || rj < 0) continue;
        M_out[(long long)ri * n + rj] += mi[i][j] * std::pow(L, mp[i][j]) * mu;
        if (K_out) K_out[(long long)ri * n + rj] += ke[i][j];
      }
  }
  return 0;
}
