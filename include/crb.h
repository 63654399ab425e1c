/* crb.h -- C ABI of libcrb.so, the B200 (sm_100a) batched beam integrator.
 *
 * Drop-in boundary for the ONE hot path of cram9030/continuum-robot: evaluating the
 * Euler-Bernoulli finite-element state-space right-hand side and stepping it (RK4 / RK45)
 * over an ensemble of beams.  The reference has no FFI (it is pure Python); each entry point
 * below names the reference interface it replaces (paths relative to
 * /root/reference/src/continuum_robot/).  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *  - every pointer marked "device" is CUDA device memory owned by the caller (PyTorch in the
 *    shipped host layer); the library never allocates, frees or synchronises device memory;
 *  - all work is enqueued on the caller's cudaStream_t (passed as void*);
 *  - return value: 0 = success, negative = error, text in crb_last_error() (thread-local);
 *  - per-member numerical outcomes (RK45 status, nfev ...) are data, not errors;
 *  - the library keeps no mutable global state besides the thread-local error string.
 *
 * Memory layout ("member" = one beam of the ensemble, B members, N elements, n free position
 * DOFs after boundary conditions, state x = [q(n) ; v(n)] in the reference's order
 * (models/dynamic_beam_model.py:120-149, DOF order [u,w,phi] per free node)):
 *     X        [B, 2n]  double, row-major, member-major (same vector the reference integrates)
 *     params   [Bp, N, 7] double: length, elastic_modulus, moment_inertia, density, cross_area,
 *              wetted_area, drag_coef  (the CSV columns of dynamic_beam_model.py:78-93);
 *              Bp = 1 (one design shared by all members) or B
 */
#ifndef CRB_H_
#define CRB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRB_VERSION 107
#define CRB_MAX_SLOTS 256      /* node slots per member handled by one lane group */
#define CRB_MAX_LEVELS 5       /* log2(32) scan levels */
#define CRB_LQR_MAX_ELEMENTS 128 /* crb_dense_matrices_batched: elements per beam */
#define CRB_LQR_MAX_N 96         /* crb_lqr_gains: free position DOFs per design (32-element cantilever) */

/* element types: models/abstractions.py:9-13 (ElementType) */
#define CRB_ELEM_LINEAR 0
#define CRB_ELEM_NONLINEAR 1
#define CRB_ELEM_ABSENT 255
/* boundary conditions: models/abstractions.py:16-20, euler_bernoulli_beam.py:240-254 */
#define CRB_BC_NONE 0
#define CRB_BC_FIXED 1
#define CRB_BC_PINNED 2

/* column order of params[..., 7] */
enum { CRB_P_LENGTH = 0, CRB_P_E = 1, CRB_P_I = 2, CRB_P_RHO = 3, CRB_P_AREA = 4,
       CRB_P_WETTED = 5, CRB_P_CD = 6, CRB_NPARAM = 7 };

/* Topology of one ensemble: how the N+1 nodes map to "slots" of a lane group.
 * A member is owned by G lanes of a warp, each lane owning m consecutive node slots
 * (P = m*G >= active nodes).  A fully FIXED node 0 is trimmed (slot s = node s + n0).
 * Filled by crb_plan() (host only). */
typedef struct crb_plan_t {
  int32_t n_elements;   /* N */
  int32_t n_free;       /* n: free position DOFs (reduced vector length) */
  int32_t n0;           /* 1 if node 0 is FIXED and trimmed, else 0 */
  int32_t p_act;        /* active slots = N + 1 - n0 */
  int32_t m;            /* slots per lane */
  int32_t g;            /* lanes per member (power of two <= 32) */
  int32_t p;            /* m * g */
  int32_t levels;       /* log2(g) */
  int32_t contiguous;   /* 1 if reduced index of (slot s, dof d) == 3 s + d for all s < p_act */
  int32_t has_mask;     /* 1 if some DOF of an active slot is constrained */
  int64_t mfac_doubles; /* doubles per mass-factor set */
  int64_t kcoef_doubles;/* doubles per stiffness-coefficient set (4 per slot) */
  int32_t red_index[3 * CRB_MAX_SLOTS]; /* (slot,dof) -> reduced index, or -1 */
} crb_plan_t;

/* Everything a RHS evaluation needs; all pointers are device pointers (NULL = feature off).
 * Replaces the closure state of DynamicEulerBernoulliBeam (dynamic_beam_model.py:16-364):
 * M_inv (:60) -> mfac, the stiffness closure (euler_bernoulli_beam.py:163-298) -> kcoef /
 * elem_type / red_index, the auto-registered forces (:220-241) -> drag / grav. */
typedef struct crb_system_t {
  int32_t n_members;        /* B */
  int32_t mass_shared;      /* 1: one mfac set for all members, 0: one per member */
  int32_t stiff_shared;     /* same for kcoef */
  int32_t force_shared;     /* same for drag / grav */
  const double* mfac;       /* [Bm, plan.mfac_doubles]   (crb_assemble) */
  const double* kcoef;      /* [Bk, P, 4]                (crb_assemble) */
  const uint8_t* elem_type; /* [P] type of the element LEFT of slot s (CRB_ELEM_*) */
  const int32_t* red_index; /* [3P] device copy of plan.red_index */
  /* FluidDragForce (models/fluid_forces.py:24-142): F[w_k] = -drag[k] * wdot*|wdot| */
  const double* drag;       /* [Bf, P] 0.5*rho_f*C_d*A_w per slot, or NULL */
  /* GravityForce (models/gravity_forces.py:6-173), reduced-index semantics */
  const double* grav;       /* [Bf, P, 2] half masses (left two-ended, tail), or NULL */
  const double* seg_half_mass; /* [Bf, N] 0.5*rho*A*L per CSV row: generic-BC gravity path */
  double gx, gy;            /* gravity_vector[0], [1] (gz ignored, gravity_forces.py:20) */
  int32_t grav_mode;        /* 0 off, 1 slot-space (contiguous plans), 2 generic reduced-index */
  /* inputs u (dynamic_beam_model.py:294-362): u = U + impulse(t) + gain (ref - x) */
  const double* u_const;    /* [B, n] generalized force held constant over the call, or NULL */
  const double* imp_amp;    /* [B] amplitude of a force on reduced DOF imp_dof for t < imp_duration */
  int32_t imp_dof;          /* reduced position index (examples use n-2: tip w) */
  double imp_duration;
  /* FullStateLinear (control/full_state_linear.py:58): u_c = gain @ (ref - x) */
  const double* gain;       /* [n, 2n] shared gain (or [B, n, 2n], see gain_stride), or NULL */
  const double* ref;        /* [2n] shared reference, NULL = 0 */
  const double* gain_frag;  /* [6m][(3m+1)/2][32] gain re-tiled as FP64 mma B-fragments (crb_gain_fragments),
                               used when plan.g == 4 (8 members per warp = the 8 rows of mma.m8n8k4); else NULL */
  /* additive external generalized force (user torch callables evaluated between launches) */
  const double* f_ext;      /* [B, n] or NULL */
  /* dispatch hints filled by the host layer from the parameter table */
  int32_t all_linear;       /* 1: every element is CRB_ELEM_LINEAR */
  int32_t all_nonlinear;    /* 1: every element is CRB_ELEM_NONLINEAR */
  int32_t force_general;    /* 1: always use the general kernels (testing / comparison) */
  int32_t force_staged;     /* 1: fast path keeps the stage-by-stage kernel (no paired operator form); per-member
                               gains stay in global memory (no shared-memory staging) -- testing / comparison */
  /* shared-operator form (crb_shared_operator): device copy of the operator blob, or NULL.  When set,
   * crb_rk4 evaluates the whole closed-loop RHS as a dense FP64 tensor-core contraction; gain / ref /
   * gravity vector / imp_dof are the ones the blob was built with. */
  const double* shared_op;
  int64_t shared_op_doubles;
  /* doubles between the gains of consecutive members: 0 = `gain` is ONE [n,2n] matrix shared by the ensemble,
   * n*2n = `gain` is [B,n,2n], one design-specific gain per member (crb_lqr_gains); gain_frag / shared_op
   * must be NULL then */
  int64_t gain_stride;
  /* per-member closed-loop operators (crb_member_operators) for ensembles of linear designs with one gain per
   * member: [B, n, 3n+1] rows [ -M^-1 (K + G_q) | -M^-1 G_v | M^-1 | M^-1 G ref ], or NULL.  When set (with
   * gain / gain_stride / ref describing the same feedback), crb_rk4 evaluates the closed-loop RHS as ONE dense
   * product per member (one lane per free DOF, the operator row in registers). */
  const double* member_op;
  /* Time-varying generalized force, evaluated at every stage time: `force = u(t)` of
   * dynamic_beam_model.py:357-360 for the input families that can be fused (the reference's tests drive
   * u(t) = sin(t) * ones(n), tests/test_dynamic_beam.py:214-215, 234-235):
   *   u(t)[r] += u_sin_amp[m, r] * sin(u_sin_omega * t + u_sin_phase)
   *   u(t)[r] += linear interpolation of u_tab_v[k, m, r] over the ascending knots u_tab_t[k], held constant
   *              outside [u_tab_t[0], u_tab_t[u_tab_k - 1]]   (numpy.interp semantics)
   * u_time_shared = 1: the tables have no member axis (u_sin_amp [n], u_tab_v [u_tab_k, n]). */
  const double* u_sin_amp;  /* [B, n] (or [n]), or NULL */
  double u_sin_omega, u_sin_phase;
  const double* u_tab_t;    /* [u_tab_k], or NULL */
  const double* u_tab_v;    /* [u_tab_k, B, n] (or [u_tab_k, n]) */
  int32_t u_tab_k;          /* >= 2 when u_tab_v is set */
  int32_t u_time_shared;
  /* Optional scratch for the persistent kernels: ONE device int32, caller-owned, used as the ticket counter of
   * their dynamic tile scheduling (zeroed on the stream by every launch that uses it, so launches that share a
   * counter must be ordered: same stream, or distinct counters).  NULL: tiles are dealt round-robin (7 % slower on
   * config 3: the warp schedulers are not fair, see crb_rk4_fast.cuh). */
  int32_t* tile_counter;
  /* Lean recording.  Callers of the reference read a few rows of sol.y after the step loop -- the tip trace
   * y[n-2] (examples/lqr_control.py:166-183) or one entry per node for the beam shape
   * (examples/example_utilities.py:173-205) -- not the whole state.  out_sel_inv (device int32 [2n], or NULL):
   * column of state entry r in a recorded frame, or -1 for entries that are not recorded; frames written by
   * crb_rk4 / crb_midpoint (Y_out) and crb_rk45 (Y_eval) are then [T, B, out_n_sel] instead of [T, B, 2n]. */
  const int32_t* out_sel_inv;
  int32_t out_n_sel;
  /* crb_rk45 only.  Launch order of the members: device int32 [B], a permutation of 0 .. B-1, or NULL for the natural
   * order.  Members of an adaptive run take different numbers of attempts and a member cannot be split, so the launch
   * ends with its longest stragglers; the hardware hands blocks out in launch order, and "longest first" fills the
   * tail with short members (solve_ensemble orders by the time each member covered in a short pilot launch). */
  const int32_t* member_order;
} crb_system_t;

/* library version (CRB_VERSION of the build) */
int crb_version(void);
/* sizeof(crb_plan_t), sizeof(crb_system_t) of the build: a binding that mirrors the structs (ctypes, cgo ...) checks
 * its own layout against these before the first call */
int crb_abi_sizes(int32_t* plan_bytes, int32_t* system_bytes);
/* Machine probe for the roofline denominator of FP64-bound kernels: enqueues `iters` rounds of 8 independent FMA chains
 * per thread on every SM and reports the flops that launch executes in *flops_out (the caller times it with events).
 * scratch: device, at least *flops_out / (16 iters) doubles (one per thread); scratch = NULL only queries flops_out. */
int crb_probe_dfma(int32_t iters, double* scratch, int64_t scratch_doubles, int64_t* flops_out, void* stream);
/* thread-local text of the last error returned on this thread ("" if none) */
const char* crb_last_error(void);

/* Host only.  Derive the slot layout for an N-element beam with boundary conditions bc[N+1]
 * (CRB_BC_*; replaces EulerBernoulliBeam.apply_boundary_conditions bookkeeping,
 * models/euler_bernoulli_beam.py:221-298).  max_slots_per_lane: 0 = library default. */
int crb_plan(int32_t n_elements, const uint8_t* bc, int32_t max_slots_per_lane, crb_plan_t* out);

/* One thread per member: element mass/stiffness assembly (segments.py:32-78,
 * euler_bernoulli_beam.py:139-161), BC reduction (:221-298), and the factorisation that
 * replaces scipy.sparse.linalg.inv(M) (dynamic_beam_model.py:60).
 *   params device [Bp,N,7]; elem_type_host [N] (CRB_ELEM_LINEAR/NONLINEAR per CSV row)
 *   n_mass / n_stiff / n_force: number of sets to emit (1 or Bp)
 *   outputs (device): mfac [n_mass, mfac_doubles], kcoef [n_stiff, P, 4],
 *     elem_type_slots [P], drag [n_force, P] (NULL to skip), grav [n_force, P, 2] (NULL to skip),
 *     seg_half_mass [n_force, N] (NULL to skip) */
int crb_assemble(const crb_plan_t* plan, const double* params, int32_t n_param_sets,
                 const uint8_t* elem_type_host, const uint8_t* bc_host,
                 int32_t n_mass, int32_t n_stiff, int32_t n_force, double fluid_density,
                 double* mfac, double* kcoef, uint8_t* elem_type_slots,
                 double* drag, double* grav, double* seg_half_mass, void* stream);

/* dX[B,2n] = f(t, X)  -- get_dynamic_system()(t, x, u) of dynamic_beam_model.py:343-362 for
 * every member: [v ; M^-1(-k(q) + f(x) + u(t))]. */
int crb_rhs(const crb_plan_t* plan, const crb_system_t* sys, const double* X, double t,
            double* dX, void* stream);

/* F_out[B,n] = sum of the enabled built-in force components at state X (drag, gravity), i.e.
 * ForceRegistry.create_aggregated_function()(x, t) restricted to the fused components
 * (models/force_registry.py:51-81, fluid_forces.py:103-142, gravity_forces.py:66-148). */
int crb_forces(const crb_plan_t* plan, const crb_system_t* sys, const double* X, double* F_out, void* stream);

/* Classical fixed-step RK4 (north_star R1), nsteps steps fused in one launch, in place on X.
 * t_k = t0 + k h; inputs evaluated at stage times.  If Y_out != NULL the state after every
 * save_every-th step is stored at Y_out[(k/save_every - 1), B, 2n]. */
int crb_rk4(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h,
            int32_t nsteps, double* Y_out, int32_t save_every, void* stream);

/* Members one full wave of the RK4 kernel chosen for (plan, sys) integrates at once on the current
 * device (SMs x resident members per SM): the natural chunk size for pipelined calls. */
int crb_rk4_wave_members(const crb_plan_t* plan, const crb_system_t* sys, int32_t* out);

/* Sub-ensemble [lo, lo+count) of a system: every per-member device pointer is offset. */
int crb_system_slice(const crb_plan_t* plan, const crb_system_t* sys, int32_t lo, int32_t count,
                     crb_system_t* out);

/* Host-resident ensembles.  The reference keeps each beam's state in host memory and fans whole
 * simulations out over processes (examples/example_utilities.py:116-170,
 * examples/beam_comparison_gravity.py:72-73); crb_rk4_host is that contract for an ensemble:
 * X_host[B,2n] (pinned host memory) is advanced in place by nsteps RK4 steps.  Members are cut into
 * chunks (chunk_members, 0 = two kernel waves); chunk c+1 is copied in and chunk c-1 copied out while
 * chunk c integrates, on three streams owned by the pipeline handle (the handle owns streams and
 * events only -- X_dev[B,2n] is caller-owned device workspace).  The call returns immediately; its
 * work is ordered after everything enqueued on `stream` so far.  Consecutive calls on the same
 * buffers and chunking overlap chunk by chunk.  crb_pipeline_wait makes `stream` wait for all
 * outstanding copies; crb_pipeline_synchronize blocks the host.  A handle belongs to one device and is
 * used by one host thread at a time (create one per thread, like a stream). */
typedef struct crb_pipeline crb_pipeline_t;
int crb_pipeline_create(crb_pipeline_t** out);
int crb_pipeline_destroy(crb_pipeline_t* p);
int crb_pipeline_wait(crb_pipeline_t* p, void* stream);
int crb_pipeline_synchronize(crb_pipeline_t* p);
int crb_rk4_host(crb_pipeline_t* p, const crb_plan_t* plan, const crb_system_t* sys, double* X_host,
                 double* X_dev, int32_t chunk_members, double t0, double h, int32_t nsteps, void* stream);

/* Implicit midpoint rule (= Newmark average acceleration) for all-linear beams, the stiff-capable
 * companion of crb_rk4: the reference's examples integrate with LSODA because explicit steps are
 * stability-limited (h < 2.8 / omega_max, examples/example_utilities.py:153-159); this rule is
 * unconditionally stable, second order and conserves the beam's quadratic energy exactly.
 *     dv = h (M + h^2/4 K)^-1 ( u(t + h/2) - K (q + h/2 v) ),   v+ = v + dv,   q+ = q + h v + h/2 dv
 * crb_assemble_shifted factors M + shift K (shift = h^2/4) into afac[n_sets, plan.mfac_doubles]
 * (n_sets = 1 for one shared design, else n_param_sets); crb_midpoint advances X[B,2n] in place by
 * nsteps fused steps (inputs: u_const / f_ext / impulse of crb_system_t, evaluated at the midpoint). */
int crb_assemble_shifted(const crb_plan_t* plan, const double* params, int32_t n_param_sets,
                         const uint8_t* elem_type_host, const uint8_t* bc_host, int32_t n_sets, double shift,
                         double* afac, void* stream);
int crb_midpoint(const crb_plan_t* plan, const crb_system_t* sys, const double* afac, int32_t afac_shared,
                 double* X, double t0, double h, int32_t nsteps, double* Y_out, int32_t save_every, void* stream);

/* Adaptive Dormand-Prince 5(4) with SciPy's controller (scipy/integrate/_ivp/rk.py:86-180),
 * one independent (t, h) per member.  Replaces solve_ivp(method="RK45") as called at
 * examples/pyodide_example/pyodide_example.py:69-75 and throughout the reference's tests.
 *   t [B] in/out current time, h_abs [B] in/out next step size (<= 0: select_initial_step),
 *   t_eval [n_eval] ascending output times (dense output), Y_eval [n_eval, B, 2n],
 *   status [B] in/out: 0 running/finished at t_bound, -1 step size too small, 1 attempt budget hit
 *   counters [B,3]: nfev, accepted, rejected.  max_attempts bounds the work of one launch.
 * Resuming: a member whose status is 1 ON ENTRY continues exactly where the budget stopped it -- same step sequence,
 * same counters as an uninterrupted run: f(t, y) is re-evaluated without being counted, and h_abs < 0 on entry means
 * "|h_abs| is the next step size and the last attempt was rejected" (rk.py:160-176: the growth factor of the next
 * accepted step is capped at 1), which is how a budget stop writes it. */
int crb_rk45(const crb_plan_t* plan, const crb_system_t* sys, double* X, double* t,
             double* h_abs, double t_bound, double rtol, double atol, const double* t_eval,
             int32_t n_eval, double* Y_eval, int32_t* status, int64_t* counters,
             int32_t max_attempts, void* stream);

/* Building blocks of the UNFUSED adaptive driver, for right-hand sides that contain user code (torch force
 * plug-ins, models/abstractions.py:153-173; free-form input callables u(t) as in the reference's own tests,
 * tests/test_dynamic_beam.py:201-244).  The host evaluates the stage derivatives K[s] = f(ts, Ys) (crb_rhs plus
 * the user's code); stage combination, error norm, SciPy's step controller (scipy/integrate/_ivp/rk.py:111-176),
 * dense output at t_eval and the set-up of the next attempt run on the device, one (t, h) per member, without a
 * host synchronisation inside an attempt.  All arrays are device memory: Y, Ys, Ynew [B, n2]; K [7, B, n2];
 * t, t_next, h_abs, h_step, ts [B]; ie, status, flags int32 [B]; counters int64 [B, 3] (nfev, accepted, rejected).
 *   crb_rk45_stage   Ys = Y + h_step * sum_{l < stage} a[stage][l] K[l], ts = t + c[stage] h_step;
 *                    stage = 1..5, or 6 for y_new (the 5th-order weights)
 *   crb_rk45_control begin_only = 1: only set up the first attempt (h_abs must hold the initial step; flags = 5:
 *                    running | first attempt of a step); begin_only = 0: close the attempt whose K[0..6], Ynew are
 *                    complete (accept: Y <- Ynew, K[0] <- K[6], t <- t_next, outputs for t_eval in (t, t_next])
 *                    and set up the next one (h_step = 0 for members that are finished or failed). */
int crb_rk45_stage(int32_t n2, int32_t n_members, int32_t stage, const double* Y, const double* K, const double* t,
                   const double* h_step, double* Ys, double* ts, void* stream);
int crb_rk45_control(int32_t n2, int32_t n_members, double* Y, double* K, const double* Ynew, double* t,
                     double* t_next, double* h_abs, double* h_step, double t_bound, double rtol, double atol,
                     const double* t_eval, int32_t n_eval, int32_t* ie, double* Y_eval, int32_t* status,
                     int64_t* counters, int32_t* flags, int32_t begin_only, void* stream);

/* Host only.  Re-tile a shared feedback gain[n,2n] (control/full_state_linear.py:58) into the
 * B-fragment order of mma.sync.m8n8k4.f64 for the lane layout of `plan` (requires plan->g == 4):
 * out[kt][nt][lane], kt < 6m (k-tile = the kt-th own state value of each of the 4 lanes of a member),
 * nt < (3m+1)/2 (n-tile = own position DOFs 2nt, 2nt+1 of each lane).  Returns the number of doubles
 * written (out may be NULL to query), or a negative error code. */
int64_t crb_gain_fragments(const crb_plan_t* plan, const double* gain_host, double* out_host);

/* Host only.  Shared-operator form of the RHS for ensembles whose members share ONE linear design,
 * gravity setting and feedback gain (BASELINE config 5, examples/lqr_control.py:87-130):
 *   a = W [q;v] + (M^-1 Gc) cos(P q) + (M^-1 Gs) sin(P q) + c0 + imp(t) M^-1 e_k,
 *   W = [-M^-1 (K + G_q) | -M^-1 G_v],  c0 = M^-1 G r   (dynamic_beam_model.py:256-272, 343-362,
 *   control/full_state_linear.py:58, gravity_forces.py:97-146 in reduced indices),
 * re-tiled as mma.sync.m8n8k4.f64 B fragments.  gain_host [n,2n] / ref_host [2n] may be NULL,
 * imp_dof = -1 for none.  Requires n_free <= 24 (and <= 8 segments with gravity).  Returns the
 * number of doubles (out_host NULL = query) or a negative error code; the caller copies the blob to
 * the device and sets crb_system_t.shared_op / shared_op_doubles. */
int64_t crb_shared_operator(const crb_plan_t* plan, const double* params_host, const uint8_t* elem_type_host,
                            const uint8_t* bc_host, const double* gain_host, const double* ref_host,
                            double gx, double gy, int32_t gravity_on, int32_t imp_dof, double* out_host);

/* Host only.  The operator blob ends with the internal DOF order crb_shared_operator chose (int32[4 KQ]) and the masks
 * of the operator tiles that hold a nonzero (uint32[5]: Wq, Wv, P, Gc, Gs).  The rollout kernel carries, per (KQ, GKP),
 * one code path compiled for the tile masks of the reference's LQR example (straight FIXED-root beam, gain without
 * axial / bending coupling, gravity along y: examples/lqr_control.py:26-84) next to the all-tiles path; a blob whose
 * masks are a subset of out5 runs the former.  Returns 0 or a negative error code. */
int crb_shared_sparse_masks(int32_t KQ, int32_t GKP, uint32_t* out5);

/* Dense BC-reduced matrices for one parameter set (host convenience for LQR synthesis;
 * replaces get_mass_matrix / get_stiffness_matrix, euler_bernoulli_beam.py:357-361,422-511).
 * params_host [N,7]; M_out, K_out host [n,n] row-major (K_out may be NULL). */
int crb_dense_matrices(const crb_plan_t* plan, const double* params_host,
                       const uint8_t* elem_type_host, const uint8_t* bc_host,
                       double* M_out, double* K_out);

/* Device.  crb_dense_matrices for every parameter set at once: the per-member A / B build of an LQR design
 * ensemble (euler_bernoulli_beam.py:139-161, 265, 422-511; segments.py:32-78).  params device
 * [n_param_sets, N, 7]; elem_type_host [N], bc_host [N+1] host; M_out, K_out device [n_param_sets, n, n]
 * row-major (K_out may be NULL).  N <= CRB_LQR_MAX_ELEMENTS. */
int crb_dense_matrices_batched(const crb_plan_t* plan, const double* params, int32_t n_param_sets,
                               const uint8_t* elem_type_host, const uint8_t* bc_host, double* M_out,
                               double* K_out, void* stream);

/* Device.  Batched LQR synthesis: replaces LinearQuadraticRegulator.compute_gain_matrix
 * (control/linear_quadratic_regulator.py:84-191; `ct.lqr(A, B, Q, R)` at :180) for an ENSEMBLE of designs.
 * Per member: A = [[0, I], [-M^-1 K, 0]], B = [[0], [M^-1]] (:84-146), the stabilising solution S of
 * A^T S + S A - S B R^-1 B^T S + Q = 0 (matrix sign function of the Hamiltonian + `refine_passes` Newton-Kleinman
 * correction steps, 1 recommended), gain = R^-1 B^T S.
 *   M_beam, K_beam  device [B or 1, n, n] (m_shared / k_shared = 1: one matrix for all members)
 *   Q device [2n,2n], R device [n,n] (shared);  gain_out device [B, n, 2n];  S_out device [B, 2n, 2n] or NULL
 *   residual_out device [B] or NULL: ||A^T S + S A - S G S + Q||_F / ||Q||_F of the returned S
 *   status_out device int32 [B]: 0 ok, 1 M or R singular, 2 no stabilising solution found (sign iteration
 *       failed), 3 closed loop A - B gain not stable (:185-189); gain/S are NaN unless status is 0 or 3
 *   workspace device, crb_lqr_workspace_bytes(n, B) bytes.  n <= CRB_LQR_MAX_N; up to n = 42 the 4n x 4n
 *   Hamiltonian lives in shared memory, beyond it in the workspace (slower: one pass over it per pivot). */
int crb_lqr_workspace_bytes(int32_t n, int32_t n_members, size_t* out);
/* Device.  Dense closed-loop operator of every member of a linear design ensemble under state feedback
 * u_c = gain (ref - x) (control/full_state_linear.py:58 inside examples/lqr_control.py:95-111):
 *   a = -M^-1 K q + M^-1 (f_gravity(q) + impulse(t) + gain (ref - x))
 *     = [ -M^-1 (K + G_q) | -M^-1 G_v ] [q; v] + M^-1 f_gravity(q) + impulse(t) M^-1 e_k + M^-1 G ref.
 * M_beam, K_beam device [B or 1, n, n]; gain device [B, n, 2n]; ref device [2n] or NULL;
 * op_out device [B, n, 3n+1]; n <= 32.  status_out device int32 [B]: 1 where M is singular. */
int crb_member_operators(int32_t n, int32_t n_members, const double* M_beam, int32_t m_shared, const double* K_beam,
                         int32_t k_shared, const double* gain, const double* ref, double* op_out,
                         int32_t* status_out, void* stream);
int crb_lqr_gains(int32_t n, int32_t n_members, const double* M_beam, int32_t m_shared, const double* K_beam,
                  int32_t k_shared, const double* Q, const double* R, int32_t refine_passes, double* gain_out,
                  double* S_out, double* residual_out, int32_t* status_out, void* workspace,
                  size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRB_H_ */
