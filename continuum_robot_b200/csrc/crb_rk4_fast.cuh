// crb_rk4_fast.cuh -- fused fixed-step RK4 for all-linear beams without built-in forces
// (BASELINE config 3 shape).  Same lane-group decomposition as crb_device.cuh, specialised so
// that the FP64 pipe, not shared-memory bandwidth, is the limiter:
//
//  * mass solve = crb_device.cuh::fast_solve_r (compact factor copy, recomputed partition
//    corrections, pure-DFMA sweeps), applied to TWO right-hand sides at once in the paired kernel;
//  * crb_rk4_fast_kernel: Nystrom form of the classical RK4 tableau (the RHS of a linear undamped
//    beam needs only the stage positions): 5 live vectors per DOF (Q0 = q + h/2 v, v,
//    S = a1+a2+a3, A = a1+2a2+2a3+a4, work) instead of 7 -- algebraically the same update as
//    k1..k4 of north_star row R1;
//  * crb_rk4_lin2_kernel: paired operator applications (see below), the default.
//
// More resident warps do not help: builds capped at 216 / 200 registers (9 / 10 warps per SM) measured
// 18-45 % slower than the 8-warp build, with the stored-spike-free T-form just as with the earlier forms.
//
// Reference behaviour: models/segments.py:32-78, euler_bernoulli_beam.py:163-298,
// dynamic_beam_model.py:256-272, 343-362 (forces disabled, u = constant force / tip impulse / none).
#pragma once
#include "crb_device.cuh"
#include "crb_tile.cuh"

#ifndef CRB_FAST_WARPS
#define CRB_FAST_WARPS 2      // warps per block of the fast kernel
#endif
#ifndef CRB_TICKET_AHEAD
#define CRB_TICKET_AHEAD 1  // 1: tile tickets are drawn one tile ahead, so the atomic's round trip overlaps a tile's integration (0.7585 -> 0.7515 ms per 20-step launch of config 3, A/B on one box)
#endif
#ifndef CRB_FAST_MINBLOCKS
#define CRB_FAST_MINBLOCKS 4  // resident blocks per SM the register allocation is sized for
#endif
#define CRB_FAST_THREADS (32 * CRB_FAST_WARPS)

// PT / PU / PS: constants pinned in registers for the whole launch (the rest is re-read from shared memory
// by every solve): 2x2 T blocks of the first PT slots, the tu values, Sinv of the first PS slots.
// Measured on B200, cfg 3 (ms per 50-step launch): none 1.886, T 1.832, T + tu 1.841, T + tu + 2 Sinv 1.805,
// everything 1.838.  Kernels that carry forcing vectors pin less (they would spill).
// NC: plan with constrained DOFs inside active slots, an untrimmed root or phantom slots (any boundary
// condition pattern): state I/O goes through the reduced-index table and constrained / phantom DOFs are
// masked before the solve (their factor rows are identity rows).
template <int M, int PT = 0, int PU = 0, int PS = 0, bool NC = false>
struct FastCtx {
  static constexpr bool kNC = NC;
  int ri[NC ? M : 1][3];  // reduced index of own DOFs (-1: constrained / phantom)
  int g;
  int member;
  bool active;
  int n;
  double4 kc[M];
  const double* fslot;  // compact factor copy, slot part (shared memory)
  const double* fscan;  // compact factor copy, scan part
  static constexpr int kPin = PT < M ? PT : M;
  static constexpr int kPinU = PU;
  static constexpr int kPinS = PS < M ? PS : M;
  double pin[kPin > 0 ? kPin : 1][4], pinu[M + 1], pins[kPinS > 0 ? kPinS : 1][4];
  // impulse: amplitude (0 if none) and the (slot, dof) it acts on, as a flat local index or -1
  double imp_amp, imp_dur;
  int imp_local;
};

// Load the pinned constants (opaque loads: the compiler would otherwise re-materialise them inside the
// step loop, i.e. not pin them at all).
template <int M, int G, typename CT>
__device__ __forceinline__ void fast_pin_load(CT& C) {
#pragma unroll
  for (int j = 0; j < CT::kPin; ++j) {
    const unsigned a2 = (unsigned)__cvta_generic_to_shared(C.fslot + (((2 * M + j) * G + C.g) << 1));
    const unsigned a3 = (unsigned)__cvta_generic_to_shared(C.fslot + (((3 * M + j) * G + C.g) << 1));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(C.pin[j][0]), "=d"(C.pin[j][1]) : "r"(a2));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(C.pin[j][2]), "=d"(C.pin[j][3]) : "r"(a3));
  }
  if (CT::kPinU) {
#pragma unroll
    for (int jj = 0; jj < (M + 1) / 2; ++jj) {
      const unsigned a4 = (unsigned)__cvta_generic_to_shared(C.fslot + (((4 * M + jj) * G + C.g) << 1));
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(C.pinu[2 * jj]), "=d"(C.pinu[2 * jj + 1]) : "r"(a4));
    }
  }
#pragma unroll
  for (int j = 0; j < CT::kPinS; ++j) {
    const unsigned a0 = (unsigned)__cvta_generic_to_shared(C.fslot + (((0 * M + j) * G + C.g) << 1));
    const unsigned a1 = (unsigned)__cvta_generic_to_shared(C.fslot + (((1 * M + j) * G + C.g) << 1));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(C.pins[j][0]), "=d"(C.pins[j][1]) : "r"(a0));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(C.pins[j][2]), "=d"(C.pins[j][3]) : "r"(a1));
  }
}

// Linear element force as VALUES: f on node 1 = (fu, V, m1), on node 2 = (-fu, -V, m2)
// (models/segments.py:32-62); c = (EA/L, 12EI/L^3, 6EI/L^2, 2EI/L).
__device__ __forceinline__ void elem_linear_vals(const double4 c, const double (&qa)[3], const double (&qb)[3],
                                                 double& fu, double& V, double& m1, double& m2) {
  const double d = qa[1] - qb[1];
  const double s = qa[2] + qb[2];
  fu = c.x * (qa[0] - qb[0]);
  V = fma(c.y, d, -c.z * s);
  const double R = fma(c.w, s, -c.z * d);
  m1 = fma(c.w, qa[2], R);
  m2 = fma(c.w, qb[2], R);
}

// Slot-space gravity constants of a lane (gravity_forces.py:97-146 in REDUCED indices; contiguous plans):
// half mass of the pseudo-segment between slot s - 1 and s, and of the tail pseudo-segment whose end node falls
// off the reduced vector.
template <int M>
struct FastGrav {
  double gl[M], gt[M];
  double gx, gy;
};
template <int M>
__device__ __forceinline__ void fast_grav_load(FastGrav<M>& Gv, const KPlan& P, const crb_system_t& S, int member, int s0) {
  const long long fo = S.force_shared ? 0ll : (long long)member * P.p;
#pragma unroll
  for (int j = 0; j < M; ++j) {
    Gv.gl[j] = S.grav[2 * (fo + s0 + j)];
    Gv.gt[j] = S.grav[2 * (fo + s0 + j) + 1];
  }
  Gv.gx = S.gx;
  Gv.gy = S.gy;
}

// a <- M^-1 (-K w + gravity(w) + impulse(t));  `w` holds the stage positions on entry, accelerations on exit.
template <int M, int LV, bool IMP, typename CT>
__device__ __forceinline__ void fast_accel(const CT& C, double (&w)[M][3], double t, const double (*uc)[3] = nullptr,
                                           const FastGrav<M>* gv = nullptr) {
  constexpr int G = 1 << LV;
  double qh[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    qh[d] = shfl_up_d(w[M - 1][d], 1, G);
    if (C.g == 0) qh[d] = 0.0;
  }
  // element j sits left of slot j; element M (left of the right neighbour's first slot) is
  // evaluated by that neighbour and its node-1 force comes back by shuffle
  double fu[M + 1], V[M + 1], m1[M + 1], m2[M];
  elem_linear_vals(C.kc[0], qh, w[0], fu[0], V[0], m1[0], m2[0]);
#pragma unroll
  for (int j = 1; j < M; ++j) elem_linear_vals(C.kc[j], w[j - 1], w[j], fu[j], V[j], m1[j], m2[j]);
  fu[M] = shfl_down_d(fu[0], 1, G);
  V[M] = shfl_down_d(V[0], 1, G);
  m1[M] = shfl_down_d(m1[0], 1, G);
  if (C.g == G - 1) { fu[M] = 0.0; V[M] = 0.0; m1[M] = 0.0; }
  double b[1][M][3];
#pragma unroll
  for (int j = 0; j < M; ++j) {
    // b = -(f_node2(element j) + f_node1(element j+1))
    b[0][j][0] = fu[j] - fu[j + 1];
    b[0][j][1] = V[j] - V[j + 1];
    b[0][j][2] = -(m2[j] + m1[j + 1]);
  }
  if (gv) {  // pseudo-segment j joins slot j - 1 (left neighbour's last slot for j = 0) and slot j
    double fa[M + 1], ft[M + 1];
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double pl = (j == 0) ? qh[2] : w[j == 0 ? 0 : j - 1][2];
      grav_pair(0.5 * (pl + w[j][2]), gv->gl[j], gv->gx, gv->gy, fa[j], ft[j]);  // zero half mass: no segment
    }
    fa[M] = shfl_down_d(fa[0], 1, G);
    ft[M] = shfl_down_d(ft[0], 1, G);
    if (C.g == G - 1) { fa[M] = 0.0; ft[M] = 0.0; }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      b[0][j][0] += fa[j] + fa[j + 1];
      b[0][j][1] += ft[j] + ft[j + 1];
      if (gv->gt[j] != 0.0) {
        double ta, tt;
        grav_pair(w[j][2], gv->gt[j], gv->gx, gv->gy, ta, tt);
        b[0][j][0] += ta;
        b[0][j][1] += tt;
      }
    }
  }
  if (IMP && C.imp_local >= 0 && t < C.imp_dur) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (C.imp_local == 3 * j + d) b[0][j][d] += C.imp_amp;
  }
  if (uc) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) b[0][j][d] += uc[j][d];
  }
  if (CT::kNC) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (C.ri[CT::kNC ? j : 0][d] < 0) b[0][j][d] = 0.0;
  }
  fast_solve_r<M, LV, 1>(b, C);
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) w[j][d] = b[0][j][d];
}

// Common prologue of the fast kernels: which member / lane this thread is, the compact factor copy staged in
// shared memory (one shared set, or with PM one region per member of the block, read from `fac`), the pinned
// constants and the lane's stiffness coefficients.  Returns the first slot of the lane.
template <int M, int LV, bool PM, typename CT>
__device__ __forceinline__ int fast_ctx_init(CT& C, const KPlan& P, const crb_system_t& S, const double* fac, double* smem) {
  constexpr int G = 1 << LV, mpw = 32 / G, LVE = LV > 0 ? LV : 1;
  constexpr int FAST_DOUBLES = crb_compact_doubles(M, G, LVE);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mloc = warp * mpw + lane / G;
  const int member = blockIdx.x * (CRB_FAST_WARPS * mpw) + mloc;
  C.g = lane % G;
  C.n = P.n_free;
  C.active = member < S.n_members;
  C.member = C.active ? member : S.n_members - 1;
  const long long off = 2 * CRB_SLOT_PAIRS * (M * G) + 2 * CRB_SCAN_PAIRS * LVE * G;
  if (PM) {
    const double* src = fac + (long long)C.member * P.mfac_doubles + off;
    double* dst = smem + mloc * FAST_DOUBLES;
    for (int k = C.g; k < FAST_DOUBLES; k += G) dst[k] = src[k];
  } else {
    const double* src = fac + off;
    for (int k = threadIdx.x; k < FAST_DOUBLES; k += blockDim.x) smem[k] = src[k];
  }
  __syncthreads();
  C.fslot = smem + (PM ? mloc * FAST_DOUBLES : 0);
  C.fscan = C.fslot + crb_compact_slot_doubles(M, G);
  fast_pin_load<M, G, CT>(C);
  const int s0 = C.g * M;
  const double* kc = S.kcoef + (S.stiff_shared ? 0ll : (long long)C.member * (M * G) * 4);
#pragma unroll
  for (int j = 0; j < M; ++j) {
    const double2 k0 = *reinterpret_cast<const double2*>(kc + 4 * (s0 + j));
    const double2 k1 = *reinterpret_cast<const double2*>(kc + 4 * (s0 + j) + 2);
    C.kc[j] = make_double4(k0.x, k0.y, k1.x, k1.y);
  }
  if (CT::kNC) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) C.ri[CT::kNC ? j : 0][d] = S.red_index[3 * (s0 + j) + d];
  }
  C.imp_amp = 0.0;
  C.imp_dur = S.imp_duration;
  C.imp_local = -1;
  return s0;
}

// impulse target of this lane: contiguous plan, reduced index = 3 slot + dof
template <int M, typename CT>
__device__ __forceinline__ void fast_ctx_impulse(CT& C, const crb_system_t& S, int s0) {
  if (!S.imp_amp) return;
  C.imp_amp = S.imp_amp[C.member];
  if (CT::kNC) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (C.ri[CT::kNC ? j : 0][d] == S.imp_dof) C.imp_local = 3 * j + d;
    return;
  }
  const int rel = S.imp_dof - 3 * s0;
  if (rel >= 0 && rel < 3 * M) C.imp_local = rel;
}

// Recorded frames of the fused integrators of this file (one warp = mpw consecutive members = consecutive rows of Y).
//   full state   registers -> the warp's staging rows in shared memory -> ONE bulk store (cp.async.bulk shared ->
//                global) per frame: every byte leaves coalesced and the copy engine, not the LSU pipe the step loop
//                lives on, moves it (one 8-byte store per lane and entry touched 32 sectors per instruction and ran
//                the implicit-midpoint kernel at 2.0 TB/s of frames; 128-bit stores through the LSU reach 3.1 TB/s)
//   lean         (crb_system_t.out_sel_inv) a per-lane bit mask of the own entries that are selected plus their output
//                columns in a small shared-memory table, both built once per launch: most lanes own no selected entry
//                and a frame costs them one test; the others store straight from their registers
// The staging rows double as the column table (a launch records full or lean frames, not both).
template <int M, bool NC>
struct FrameWriter {
  unsigned selmask;
  double* stage;  // the warp's staging rows [mpw][2n] (16-byte aligned), or NULL: direct stores
  template <typename RIX>
  __device__ __forceinline__ void init(const crb_system_t& S, const double* Y, int n, double* stage_rows, RIX rix) {
    selmask = 0;
    stage = stage_rows;
    if (Y && S.out_sel_inv) {
      int* col = reinterpret_cast<int*>(stage_rows) + (threadIdx.x & 31);  // [entry][lane]
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int r = rix(j, d);
          if (!NC || r >= 0) {
            const int cq = S.out_sel_inv[r], cv = S.out_sel_inv[n + r];
            if (cq >= 0) selmask |= 1u << (3 * j + d);
            if (cv >= 0) selmask |= 1u << (3 * M + 3 * j + d);
            if (stage_rows) {
              col[(3 * j + d) * 32] = cq;
              col[(3 * M + 3 * j + d) * 32] = cv;
            }
          }
        }
    }
  }
  // m0: first member of the warp, ml: this lane's member within the warp, cnt: members of the warp inside the ensemble
  template <typename RIX>
  __device__ __forceinline__ void write(const crb_system_t& S, double* __restrict__ Y, long long frame, int m0, int ml, int cnt,
                                        int n, const double (&q)[M][3], const double (&v)[M][3], RIX rix) const {
    const int lane = threadIdx.x & 31;
    if (S.out_sel_inv) {
      if (selmask && ml < cnt) {
        double* ym = Y + (frame * S.n_members + m0 + ml) * S.out_n_sel;
        const int* col = reinterpret_cast<const int*>(stage) + lane;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            if (selmask & (1u << (3 * j + d))) ym[stage ? col[(3 * j + d) * 32] : S.out_sel_inv[rix(j, d)]] = q[j][d];
            if (selmask & (1u << (3 * M + 3 * j + d)))
              ym[stage ? col[(3 * M + 3 * j + d) * 32] : S.out_sel_inv[n + rix(j, d)]] = v[j][d];
          }
      }
      return;
    }
    if (!stage) {  // no staging rows (Y rows not 16-byte aligned): direct stores
      if (ml < cnt) {
        double* ym = Y + (frame * S.n_members + m0 + ml) * 2 * n;
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int r = rix(j, d);
            if (!NC || r >= 0) {
              ym[r] = q[j][d];
              ym[n + r] = v[j][d];
            }
          }
      }
      return;
    }
    put_rows(ml, n, q, v, rix);
    send(Y + (frame * S.n_members + m0) * 2 * n, cnt, n);
  }
  // registers -> staging rows (after the previous bulk store has drained them)
  template <typename RIX>
  __device__ __forceinline__ void put_rows(int ml, int n, const double (&q)[M][3], const double (&v)[M][3], RIX rix) const {
    constexpr bool VEC = !NC && (3 * M) % 2 == 0;
    if ((threadIdx.x & 31) == 0) crb_bulk_wait_read<0>();
    __syncwarp();
    double* xo = stage + ml * 2 * n;
    if (VEC) {
      const int r0 = rix(0, 0);
      double2* q2 = reinterpret_cast<double2*>(xo + r0);
      double2* v2 = reinterpret_cast<double2*>(xo + n + r0);
#pragma unroll
      for (int i = 0; i < (3 * M) / 2; ++i) {
        q2[i] = make_double2((&q[0][0])[2 * i], (&q[0][0])[2 * i + 1]);
        v2[i] = make_double2((&v[0][0])[2 * i], (&v[0][0])[2 * i + 1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int r = rix(j, d);
          if (!NC || r >= 0) {
            xo[r] = q[j][d];
            xo[n + r] = v[j][d];
          }
        }
    }
  }
  // staging rows -> cnt rows of 2n doubles at dst, as one bulk store
  __device__ __forceinline__ void send(double* dst, int cnt, int n) const {
    crb_fence_proxy_async();  // generic-proxy writes above -> visible to the bulk-copy (async) proxy
    __syncwarp();
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
      crb_bulk_store(dst, crb_smem_u32(stage), (unsigned)cnt * 16u * (unsigned)n);
      crb_bulk_commit();
    }
  }
  // before the block retires: the last bulk store has read the staging rows
  __device__ __forceinline__ void drain() const {
    if (stage && (threadIdx.x & 31) == 0) crb_bulk_wait_read<0>();
  }
};

// GRAV: slot-space gravity (config 1 as an ensemble: linear beams under gravity; the force depends on the stage
// positions only, so the Nystrom form still applies).  NC: see FastCtx.
// REC: frames are recorded (FrameWriter code compiled in; without it the step loop keeps the registers of the
// recording-free kernel: measured 7-9 % on the one-tile-per-block kernels)
template <int M, int LV, bool IMP, bool GRAV = false, bool NC = false, bool REC = false>
__global__ void __launch_bounds__(CRB_FAST_THREADS, CRB_FAST_MINBLOCKS)
crb_rk4_fast_kernel(KPlan P, crb_system_t S, double* __restrict__ X, double t0, double h,
                    int nsteps, double* __restrict__ Y, int save_every, int stage_off, int stage_stride) {
  extern __shared__ __align__(16) double smem[];
  double* const stage_rows = stage_off ? smem + stage_off : nullptr;  // staging rows of recorded frames (FrameWriter)
  typedef FastCtx<M, 0, 0, 0, NC> Ctx;
  Ctx C;
  const int s0 = fast_ctx_init<M, LV, false>(C, P, S, S.mfac, smem);
  if (IMP) fast_ctx_impulse<M>(C, S, s0);
  FastGrav<M> Gv;
  if (GRAV) fast_grav_load<M>(Gv, P, S, C.member, s0);

  // reduced index of own DOF (j, d): consecutive on contiguous plans without phantom slots, else the plan's table
  const int n = C.n;
  auto rix = [&](int j, int d) -> int { return NC ? C.ri[NC ? j : 0][d] : 3 * (s0 + j) + d; };
  // recorded frames: the warp's members are consecutive rows of Y (FrameWriter; only the selection mask lives across
  // the step loop, everything else is recomputed when a frame is written)
  FrameWriter<M, NC> FW;
  if (REC) FW.init(S, Y, n, stage_rows ? stage_rows + (threadIdx.x >> 5) * stage_stride : nullptr, rix);
  auto write_frame = [&](long long frame, const double (&fq)[M][3], const double (&fv)[M][3]) {
    constexpr int fw_mpw = 32 >> LV;
    const int fw_m0 = blockIdx.x * (CRB_FAST_WARPS * fw_mpw) + (threadIdx.x >> 5) * fw_mpw;
    FW.stage = stage_rows ? stage_rows + (threadIdx.x >> 5) * stage_stride : nullptr;
    FW.write(S, Y, frame, fw_m0, (threadIdx.x & 31) >> LV, max(0, min(fw_mpw, S.n_members - fw_m0)), n, fq, fv, rix);
  };
  double* xm = X + (long long)C.member * 2 * n;
  double Q0[M][3], v[M][3], Sa[M][3], Aa[M][3], w[M][3];
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int r = rix(j, d);
      w[j][d] = (!NC || r >= 0) ? xm[r] : 0.0;
      v[j][d] = (!NC || r >= 0) ? xm[n + r] : 0.0;
      Q0[j][d] = fma(0.5 * h, v[j][d], w[j][d]);
    }
  const double hh = 0.5 * h, h6 = h / 6.0, hq = 0.25 * h * h, hs = 0.5 * h * h, hx = h * h / 6.0;
  for (int k = 0; k < nsteps; ++k) {
    const double t = t0 + k * h;
    // Nystrom form of the classical tableau.  Invariant at loop entry: w = q, Q0 = q + h/2 v.
    // The stage loop stays rolled so that one copy of the RHS serves all four stages (I-cache).
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
      const double ts = t + (st == 0 ? 0.0 : (st == 3 ? h : hh));
      fast_accel<M, LV, IMP>(C, w, ts, nullptr, GRAV ? &Gv : nullptr);
      if (st == 0) {  // a1: next input q + h/2 v
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            Sa[j][d] = w[j][d];
            Aa[j][d] = w[j][d];
            w[j][d] = Q0[j][d];
          }
      } else if (st == 1) {  // a2: next input q + h/2 v + h^2/4 a1
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double a2 = w[j][d], a1 = Sa[j][d];
            Sa[j][d] = a1 + a2;
            Aa[j][d] = fma(2.0, a2, a1);
            w[j][d] = fma(hq, a1, Q0[j][d]);
          }
      } else if (st == 2) {  // a3: next input q + h v + h^2/2 a2  (a2 = A - S)
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double a3 = w[j][d];
            const double a2 = Aa[j][d] - Sa[j][d];
            Sa[j][d] += a3;
            Aa[j][d] = fma(2.0, a3, Aa[j][d]);
            Q0[j][d] = fma(hh, v[j][d], Q0[j][d]);  // now q + h v
            w[j][d] = fma(hs, a2, Q0[j][d]);
          }
      } else {  // a4: close the step, set up the next one
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double A4 = Aa[j][d] + w[j][d];
            const double qn = fma(hx, Sa[j][d], Q0[j][d]);  // q+ = q + h v + h^2/6 (a1+a2+a3)
            const double vn = fma(h6, A4, v[j][d]);         // v+ = v + h/6 (a1+2a2+2a3+a4)
            w[j][d] = qn;
            v[j][d] = vn;
            Q0[j][d] = fma(hh, vn, qn);
          }
      }
    }
    if (REC && Y && save_every > 0 && (k + 1) % save_every == 0) write_frame((k + 1) / save_every - 1, w, v);
  }
  if (C.active) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = rix(j, d);
        if (!NC || r >= 0) {
          xm[r] = w[j][d];
          xm[n + r] = v[j][d];
        }
      }
  }
  if (REC) {
    FW.stage = stage_rows;
    FW.drain();
  }
}

// ==========================================================================================
// Paired variant for the autonomous force-free linear case (config 3: u = 0).
//
// For x' = J x with J = [[0, I], [-A, 0]], A = M^-1 K, the classical RK4 update is the degree-4
// Taylor polynomial of hJ; with F(w) = -A w,
//     a1 = F(q), p = F(v), r = F(a1), s = F(p)
//     a2 = a1 + h/2 p,  a3 = a1 + h/2 p + h^2/4 r,  a4 = a1 + h p + h^2/2 r + h^3/4 s
//     q+ = q + h/6 (k1q + 2 k2q + 2 k3q + k4q) = q + h v + h^2/2 a1 + h^3/6 p + h^4/24 r
//     v+ = v + h/6 (a1 + 2 a2 + 2 a3 + a4)     = v + h a1 + h^2/2 p + h^3/6 r + h^4/24 s
// i.e. the same four stage derivatives k1..k4 of north_star row R1, evaluated as TWO rounds of TWO
// independent operator applications.  The pair shares every mass constant read from shared
// memory (half the LDS traffic per step) and gives the FP64 pipe two independent dependency
// chains per lane.
// ==========================================================================================
// w[r] <- -M^-1 K w[r] for r = 0, 1 (two independent operator applications).
template <int M, int LV, typename CT>
__device__ __forceinline__ void fast_apply2(const CT& C, double (&w)[2][M][3]) {
  constexpr int G = 1 << LV;
  double b[2][M][3];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    double qh[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      qh[d] = shfl_up_d(w[r][M - 1][d], 1, G);
      if (C.g == 0) qh[d] = 0.0;
    }
    double fu[M + 1], V[M + 1], m1[M + 1], m2[M];
    elem_linear_vals(C.kc[0], qh, w[r][0], fu[0], V[0], m1[0], m2[0]);
#pragma unroll
    for (int j = 1; j < M; ++j) elem_linear_vals(C.kc[j], w[r][j - 1], w[r][j], fu[j], V[j], m1[j], m2[j]);
    fu[M] = shfl_down_d(fu[0], 1, G);
    V[M] = shfl_down_d(V[0], 1, G);
    m1[M] = shfl_down_d(m1[0], 1, G);
    if (C.g == G - 1) { fu[M] = 0.0; V[M] = 0.0; m1[M] = 0.0; }
#pragma unroll
    for (int j = 0; j < M; ++j) {
      b[r][j][0] = fu[j] - fu[j + 1];
      b[r][j][1] = V[j] - V[j + 1];
      b[r][j][2] = -(m2[j] + m1[j + 1]);
    }
    if (CT::kNC) {
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          if (C.ri[CT::kNC ? j : 0][d] < 0) b[r][j][d] = 0.0;
    }
  }
  fast_solve_r<M, LV, 2>(b, C);
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) w[r][j][d] = b[r][j][d];
}

// With a forcing that is piecewise constant in time (constant generalized force UC and / or a
// tip impulse IMP), c_i = M^-1 u(t_i) at the stage times (c2 = c3):
//     L = F(q), p = F(v);  a1 = L + c1,  a2 = L + h/2 p + c2;  r1 = F(a1), r2 = F(a2)
//     a3 = a2 + h^2/4 r1,  a4 = a2 + h/2 p + (c4 - c2) + h^2/2 r2
//     q+ = q + h v + h^2/6 (a1 + 2 a2) + h^4/24 r1
//     v+ = v + h/6 (a1 + 5 a2 + h/2 p + c4 - c2) + h^3/12 (r1 + r2)
// (reduces to the force-free formulas for c = 0).  M^-1 u_const and M^-1 e_k are obtained once per
// launch by one paired solve.
// PM: every member has its OWN mass factors (density / area / lengths differ per member): each lane group stages its member's compact factor copy in its own shared-memory
// region.
template <int M, int LV, bool UC, bool IMP, bool PM, bool NC = false, bool REC = false>
__global__ void __launch_bounds__(CRB_FAST_THREADS, CRB_FAST_MINBLOCKS)
crb_rk4_lin2_kernel(KPlan P, crb_system_t S, double* __restrict__ X, double t0, double h, int nsteps,
                    double* __restrict__ Y, int save_every, int stage_off, int stage_stride) {
  extern __shared__ __align__(16) double smem[];
  double* const stage_rows = stage_off ? smem + stage_off : nullptr;  // staging rows of recorded frames (FrameWriter)
  typedef FastCtx<M, ((UC || IMP) ? 0 : 4), ((UC || IMP) ? 0 : 1), ((UC || IMP || NC) ? 0 : 2), NC> Ctx;  // forcing vectors / index tables need the registers
  Ctx C;
  const int s0 = fast_ctx_init<M, LV, PM>(C, P, S, S.mfac, smem);
  const int n = C.n;
  // reduced index of own DOF (j, d): 3 (s0 + j) + d on contiguous plans without phantom slots, else the
  // plan's table (-1: constrained or phantom, held at zero)
  auto rix = [&](int j, int d) -> int { return NC ? C.ri[NC ? j : 0][d] : 3 * (s0 + j) + d; };
  // recorded frames: the warp's members are consecutive rows of Y (FrameWriter; only the selection mask lives across
  // the step loop, everything else is recomputed when a frame is written)
  FrameWriter<M, NC> FW;
  if (REC) FW.init(S, Y, n, stage_rows ? stage_rows + (threadIdx.x >> 5) * stage_stride : nullptr, rix);
  auto write_frame = [&](long long frame, const double (&fq)[M][3], const double (&fv)[M][3]) {
    constexpr int fw_mpw = 32 >> LV;
    const int fw_m0 = blockIdx.x * (CRB_FAST_WARPS * fw_mpw) + (threadIdx.x >> 5) * fw_mpw;
    FW.stage = stage_rows ? stage_rows + (threadIdx.x >> 5) * stage_stride : nullptr;
    FW.write(S, Y, frame, fw_m0, (threadIdx.x & 31) >> LV, max(0, min(fw_mpw, S.n_members - fw_m0)), n, fq, fv, rix);
  };
  double* xm = X + (long long)C.member * 2 * n;
  double q[M][3], v[M][3], w[2][M][3];
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int r = rix(j, d);
      q[j][d] = (!NC || r >= 0) ? xm[r] : 0.0;
      v[j][d] = (!NC || r >= 0) ? xm[n + r] : 0.0;
    }
  // forcing in acceleration space: cu = M^-1 (u_const + f_ext), ci = amp * M^-1 e_k
  double cu[UC ? M : 1][3], ci[IMP ? M : 1][3];
  if (UC || IMP) {
    const long long mo = (long long)C.member * n;
    const double amp = IMP ? S.imp_amp[C.member] : 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = rix(j, d);
        double u = 0.0;
        if (!NC || r >= 0) {
          if (UC && S.u_const) u += S.u_const[mo + r];
          if (UC && S.f_ext) u += S.f_ext[mo + r];
        }
        w[0][j][d] = u;
        w[1][j][d] = (IMP && r == S.imp_dof) ? amp : 0.0;
      }
    fast_solve_r<M, LV, 2>(w, C);
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (UC) cu[j][d] = w[0][j][d];
        if (IMP) ci[j][d] = w[1][j][d];
      }
  }
  const double hh = 0.5 * h, h2 = 0.5 * h * h, h3 = h * h * h / 6.0, h4 = h * h * h * h / 24.0;
  const double h6 = h / 6.0, hx = h * h / 6.0, h12 = h * h * h / 12.0;
  for (int k = 0; k < nsteps; ++k) {
    const double t = t0 + k * h;
    const double g1 = (IMP && t < S.imp_duration) ? 1.0 : 0.0;
    const double g2 = (IMP && t + hh < S.imp_duration) ? 1.0 : 0.0;
    const double g4 = (IMP && t + h < S.imp_duration) ? 1.0 : 0.0;
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        w[0][j][d] = q[j][d];
        w[1][j][d] = v[j][d];
      }
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
      fast_apply2<M, LV>(C, w);
      if (round == 0) {  // w = (L, p)
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double Lq = w[0][j][d], p = w[1][j][d];
            if (!UC && !IMP) {
              q[j][d] = fma(h3, p, fma(h2, Lq, fma(h, v[j][d], q[j][d])));
              v[j][d] = fma(h2, p, fma(h, Lq, v[j][d]));
            } else {
              const double cuv = UC ? cu[j][d] : 0.0, civ = IMP ? ci[j][d] : 0.0;
              const double a1 = Lq + fma(g1, civ, cuv);
              const double a2 = fma(hh, p, Lq) + fma(g2, civ, cuv);
              q[j][d] = fma(hx, fma(2.0, a2, a1), fma(h, v[j][d], q[j][d]));
              v[j][d] = fma(h6, fma(5.0, a2, a1) + fma(hh, p, (g4 - g2) * civ), v[j][d]);
              w[0][j][d] = a1;
              w[1][j][d] = a2;
            }
          }
      } else {  // w = (r, s) or (r1, r2)
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double r = w[0][j][d], s = w[1][j][d];
            if (!UC && !IMP) {
              q[j][d] = fma(h4, r, q[j][d]);
              v[j][d] = fma(h4, s, fma(h3, r, v[j][d]));
            } else {
              q[j][d] = fma(h4, r, q[j][d]);
              v[j][d] = fma(h12, r + s, v[j][d]);
            }
          }
      }
    }
    if (REC && Y && save_every > 0 && (k + 1) % save_every == 0) write_frame((k + 1) / save_every - 1, q, v);
  }
  if (C.active) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = rix(j, d);
        if (!NC || r >= 0) {
          xm[r] = q[j][d];
          xm[n + r] = v[j][d];
        }
      }
  }
  if (REC) {
    FW.stage = stage_rows;
    FW.drain();
  }
}

// ==========================================================================================
// Persistent form of the paired kernel (shared mass factors): every warp loops over tiles of `mpw` consecutive
// members; the next tile's state rows and stiffness coefficients arrive by 1-D bulk copies (crb_tile.cuh) while the
// current tile integrates in registers, results leave through a shared-memory buffer as one bulk store per tile
// (and per recorded frame).  Same arithmetic, statement by statement, as crb_rk4_lin2_kernel.
//
// Lean recording (crb_system_t.out_sel_inv): frames are Y[T, B, out_n_sel] holding only the selected entries (tip
// trace, node shapes: what the reference's callers read from sol.y, examples/lqr_control.py:166-183,
// examples/example_utilities.py:173-205), written straight from the registers.
// ==========================================================================================
#define CRB_PERSIST_WARPS 1  // warps per block of the persistent kernels (see the barrier note in the kernel)
template <int M, int LV>
struct FastTileGeom {
  static constexpr int G = 1 << LV, mpw = 32 / G, LVE = LV > 0 ? LV : 1;
  static constexpr int FAST_DOUBLES = crb_compact_doubles(M, G, LVE);
  static constexpr int FAC_PAD = (FAST_DOUBLES + 15) & ~15;  // the tile buffers behind it stay 128-byte aligned
  static constexpr int ROW_MAX = 2 * 3 * M * G;              // doubles per state row (= 2n on full contiguous plans)
  static constexpr int KC_ROW = 4 * M * G;                   // stiffness coefficients per member
  static constexpr int WARP_DOUBLES = mpw * (2 * ROW_MAX + KC_ROW);  // in: state | in: kcoef | out: state
  static constexpr int SEL_DOUBLES = 3 * M * 32;  // lean recording: output columns [6M entries][32 lanes] as int32
  static constexpr size_t smem_bytes(int warps) {  // factors | tile buffers | mbarrier (+ pad) | column table
    return sizeof(double) * ((size_t)FAC_PAD + (size_t)warps * (WARP_DOUBLES + 2 + SEL_DOUBLES));
  }
};

// Persistent tile driver shared by the fused integrators of this file.  One warp per block, one tile (mpw
// consecutive members) at a time:
//   * the first tile of a block is blockIdx.x, the following ones come from a global ticket counter (tile_counter,
//     zeroed by the launcher) when the caller provides one: the warp schedulers are not fair (a favoured warp finishes
//     a static share of the tiles up to 25 % early and leaves its sub-partition with a single warp and no latency
//     hiding for the rest of the launch -- measured 7.0 instead of 7.9 warps active per SM and 7 % more time per
//     step); tickets keep every warp busy to the end.  Without a counter the tiles are dealt round-robin;
//   * the next tile's state rows and stiffness coefficients arrive by bulk copies while the current tile integrates;
//   * the final state leaves as one bulk store per tile; recorded full-state frames go registers -> shared memory ->
//     128-bit stores that are contiguous across the warp (generic proxy: no async-proxy fence per frame); lean frames
//     (crb_system_t.out_sel_inv) are written straight from the registers of the few lanes that own a selected entry.
// tile_setup(q, v): per-tile preparation (forcing vectors ...); step(k, q, v): one time step in registers.
template <int M, int LV, bool NC, typename Ctx, typename TileSetup, typename Step>
__device__ __forceinline__ void fast_persistent_run(const KPlan& P, const crb_system_t& S, const double* __restrict__ fac_set,
                                                    double* __restrict__ X, int nsteps, double* __restrict__ Y, int save_every,
                                                    int* __restrict__ tile_counter, double* smem, Ctx& C, TileSetup&& tile_setup,
                                                    Step&& step) {
  typedef FastTileGeom<M, LV> TG;
  constexpr int G = TG::G, mpw = TG::mpw;
  constexpr bool VEC = !NC && (3 * M) % 2 == 0;  // 128-bit shared-memory accesses of the lane's 3M-double runs
  const int lane = threadIdx.x & 31;
  const int ml = lane / G;
  double* const in_x = smem + TG::FAC_PAD;
  double* const in_k = in_x + mpw * TG::ROW_MAX;
  double* const out_x = in_k + mpw * TG::KC_ROW;
  const unsigned bar = crb_smem_u32(smem + TG::FAC_PAD + TG::WARP_DOUBLES);
  const int n = P.n_free;
  const unsigned row_bytes = 16u * (unsigned)n;
  const int n_tiles = (S.n_members + mpw - 1) / mpw;

  auto tile_count = [&](int bt) -> int {  // members of tile bt
    return max(0, min(mpw, S.n_members - bt * mpw));
  };
  auto prefetch = [&](int bt) {  // one lane: the tile's state rows (+ stiffness coefficients) -> `in` buffers
    const int m0 = bt * mpw;
    const unsigned cnt = (unsigned)tile_count(bt);
    const unsigned xb = cnt * row_bytes, kb = S.stiff_shared ? 0u : cnt * (unsigned)(TG::KC_ROW * 8);
    crb_mbar_expect_tx(bar, xb + kb);
    crb_bulk_load(crb_smem_u32(in_x), X + (long long)m0 * 2 * n, xb, bar);
    if (kb) crb_bulk_load(crb_smem_u32(in_k), S.kcoef + (long long)m0 * TG::KC_ROW, kb, bar);
  };
  if (lane == 0) {
    crb_mbar_init(bar, 1);
    crb_fence_mbar_init();
    if ((int)blockIdx.x < n_tiles) prefetch(blockIdx.x);
  }
  for (int k = threadIdx.x; k < TG::FAST_DOUBLES; k += blockDim.x) smem[k] = fac_set[k];  // compact factor copy, once per launch
  __syncthreads();
  C.g = lane % G;
  C.n = n;
  C.fslot = smem;
  C.fscan = smem + crb_compact_slot_doubles(M, G);
  fast_pin_load<M, G, Ctx>(C);
  C.imp_amp = 0.0;
  C.imp_dur = S.imp_duration;
  C.imp_local = -1;
  const int s0 = C.g * M;
  if (NC) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) C.ri[NC ? j : 0][d] = S.red_index[3 * (s0 + j) + d];
  }
  if (S.stiff_shared) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
      const double2 k0 = *reinterpret_cast<const double2*>(S.kcoef + 4 * (s0 + j));
      const double2 k1 = *reinterpret_cast<const double2*>(S.kcoef + 4 * (s0 + j) + 2);
      C.kc[j] = make_double4(k0.x, k0.y, k1.x, k1.y);
    }
  }
  auto rix = [&](int j, int d) -> int { return NC ? C.ri[NC ? j : 0][d] : 3 * (s0 + j) + d; };
  // recorded frames and the final state leave through the `out` rows (FrameWriter); lean recordings keep their column
  // table in a region of its own, because the `out` rows are rewritten by every tile's final store
  FrameWriter<M, NC> FW, FS;
  FW.init(S, Y, n, (Y && S.out_sel_inv) ? smem + TG::FAC_PAD + TG::WARP_DOUBLES + 2 : out_x, rix);
  FS.init(S, nullptr, n, out_x, rix);
  unsigned phase = 0;
  int nxt_ticket = 0;  // CRB_TICKET_AHEAD: the tile after the current one (first ticket drawn before the loop)
  if (CRB_TICKET_AHEAD && tile_counter) {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(tile_counter, 1);
    nxt_ticket = __shfl_sync(CRB_FULL_MASK, ticket, 0) + (int)gridDim.x;
  }

  for (int bt = blockIdx.x; bt < n_tiles;) {
    const int m0 = bt * mpw;
    const int member = m0 + ml;
    C.active = member < S.n_members;
    C.member = C.active ? member : S.n_members - 1;
    const unsigned cnt = (unsigned)tile_count(bt);
    crb_mbar_wait(bar, phase);
    phase ^= 1u;
    double q[M][3], v[M][3];
    {
      const double* xs = in_x + ml * 2 * n;
      if (VEC) {
        const double2* q2 = reinterpret_cast<const double2*>(xs + 3 * s0);
        const double2* v2 = reinterpret_cast<const double2*>(xs + n + 3 * s0);
#pragma unroll
        for (int i = 0; i < (3 * M) / 2; ++i) {
          const double2 a = q2[i], b = v2[i];
          (&q[0][0])[2 * i] = a.x;
          (&q[0][0])[2 * i + 1] = a.y;
          (&v[0][0])[2 * i] = b.x;
          (&v[0][0])[2 * i + 1] = b.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int r = rix(j, d);
            q[j][d] = (!NC || r >= 0) ? xs[r] : 0.0;
            v[j][d] = (!NC || r >= 0) ? xs[n + r] : 0.0;
          }
      }
      if (!S.stiff_shared) {
        const double2* k2 = reinterpret_cast<const double2*>(in_k + ml * TG::KC_ROW + 4 * s0);
#pragma unroll
        for (int j = 0; j < M; ++j) {
          const double2 k0 = k2[2 * j], k1 = k2[2 * j + 1];
          C.kc[j] = make_double4(k0.x, k0.y, k1.x, k1.y);
        }
      }
    }
    __syncwarp();  // every lane has its tile in registers: the `in` buffers are free for the next tile
    int nxt = bt + (int)gridDim.x;
    int ticket2 = 0;
    if (tile_counter) {
      if (CRB_TICKET_AHEAD) {
        // the ticket drawn during the PREVIOUS tile names the tile to prefetch now; the one drawn here is read after
        // the step loop, so the atomic's round trip overlaps the integration instead of delaying the prefetch
        nxt = nxt_ticket;
        if (lane == 0) ticket2 = atomicAdd(tile_counter, 1);
      } else {
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(tile_counter, 1);
        nxt = __shfl_sync(CRB_FULL_MASK, ticket, 0) + (int)gridDim.x;
      }
    }
    if (lane == 0 && nxt < n_tiles) prefetch(nxt);
    // One-warp blocks: this block barrier is a WARP convergence point that costs nothing, and it is there for the
    // compiler -- after the barrier spin and the one-lane prefetch it cannot prove that the warp is converged and
    // would guard every shuffle of the step loop (BRA.DIV + register copies: +8 % instructions, measured 15 % slower).
    // Blocks of several warps would be tied together by it and run their identical instruction streams in lockstep,
    // colliding on the FP64 and shuffle pipes instead of interleaving (measured 8 % slower than free-running warps).
    __syncthreads();
    tile_setup(q, v);

    for (int k = 0; k < nsteps; ++k) {
      step(k, q, v);
      if (Y && save_every > 0 && (k + 1) % save_every == 0) {
        FW.write(S, Y, (k + 1) / save_every - 1, m0, ml, (int)cnt, n, q, v, rix);
        __syncthreads();  // convergence point for the compiler (see above)
      }
    }
    // final state of the tile: one bulk store
    FS.put_rows(ml, n, q, v, rix);
    FS.send(X + (long long)m0 * 2 * n, (int)cnt, n);
    if (CRB_TICKET_AHEAD && tile_counter) nxt_ticket = __shfl_sync(CRB_FULL_MASK, ticket2, 0) + (int)gridDim.x;
    __syncthreads();
    bt = nxt;
  }
  if (lane == 0) crb_bulk_wait<0>();  // the last store has left shared memory before the block retires
}

template <int M, int LV, bool UC, bool IMP, bool NC = false>
__global__ void __launch_bounds__(32 * CRB_PERSIST_WARPS, CRB_FAST_MINBLOCKS * CRB_FAST_WARPS / CRB_PERSIST_WARPS)
crb_rk4_lin2p_kernel(KPlan P, crb_system_t S, double* __restrict__ X, double t0, double h, int nsteps,
                     double* __restrict__ Y, int save_every, int* __restrict__ tile_counter) {
  typedef FastTileGeom<M, LV> TG;
  extern __shared__ __align__(128) double smem[];
  typedef FastCtx<M, ((UC || IMP) ? 0 : 4), ((UC || IMP) ? 0 : 1), ((UC || IMP || NC) ? 0 : 2), NC> Ctx;
  Ctx C;
  const int n = P.n_free;
  const double hh = 0.5 * h, h2 = 0.5 * h * h, h3 = h * h * h / 6.0, h4 = h * h * h * h / 24.0;
  const double h6 = h / 6.0, hx = h * h / 6.0, h12 = h * h * h / 12.0;
  // forcing in acceleration space: cu = M^-1 (u_const + f_ext), ci = amp * M^-1 e_k (per tile)
  double cu[UC ? M : 1][3], ci[IMP ? M : 1][3];
  auto tile_setup = [&](double (&q)[M][3], double (&v)[M][3]) {
    if (UC || IMP) {
      double w[2][M][3];
      const int s0 = C.g * M;
      const long long mo = (long long)C.member * n;
      const double amp = IMP ? S.imp_amp[C.member] : 0.0;
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const int r = NC ? C.ri[NC ? j : 0][d] : 3 * (s0 + j) + d;
          double u = 0.0;
          if (!NC || r >= 0) {
            if (UC && S.u_const) u += S.u_const[mo + r];
            if (UC && S.f_ext) u += S.f_ext[mo + r];
          }
          w[0][j][d] = u;
          w[1][j][d] = (IMP && r == S.imp_dof) ? amp : 0.0;
        }
      fast_solve_r<M, LV, 2>(w, C);
#pragma unroll
      for (int j = 0; j < M; ++j)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          if (UC) cu[j][d] = w[0][j][d];
          if (IMP) ci[j][d] = w[1][j][d];
        }
    }
  };
  auto step = [&](int k, double (&q)[M][3], double (&v)[M][3]) {
    const double t = t0 + k * h;
    const double g1 = (IMP && t < S.imp_duration) ? 1.0 : 0.0;
    const double g2 = (IMP && t + hh < S.imp_duration) ? 1.0 : 0.0;
    const double g4 = (IMP && t + h < S.imp_duration) ? 1.0 : 0.0;
    double w[2][M][3];
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        w[0][j][d] = q[j][d];
        w[1][j][d] = v[j][d];
      }
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
      fast_apply2<M, LV>(C, w);
      if (round == 0) {  // w = (L, p)
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double Lq = w[0][j][d], p = w[1][j][d];
            if (!UC && !IMP) {
              q[j][d] = fma(h3, p, fma(h2, Lq, fma(h, v[j][d], q[j][d])));
              v[j][d] = fma(h2, p, fma(h, Lq, v[j][d]));
            } else {
              const double cuv = UC ? cu[j][d] : 0.0, civ = IMP ? ci[j][d] : 0.0;
              const double a1 = Lq + fma(g1, civ, cuv);
              const double a2 = fma(hh, p, Lq) + fma(g2, civ, cuv);
              q[j][d] = fma(hx, fma(2.0, a2, a1), fma(h, v[j][d], q[j][d]));
              v[j][d] = fma(h6, fma(5.0, a2, a1) + fma(hh, p, (g4 - g2) * civ), v[j][d]);
              w[0][j][d] = a1;
              w[1][j][d] = a2;
            }
          }
      } else {  // w = (r, s) or (r1, r2)
#pragma unroll
        for (int j = 0; j < M; ++j)
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double r = w[0][j][d], s = w[1][j][d];
            if (!UC && !IMP) {
              q[j][d] = fma(h4, r, q[j][d]);
              v[j][d] = fma(h4, s, fma(h3, r, v[j][d]));
            } else {
              q[j][d] = fma(h4, r, q[j][d]);
              v[j][d] = fma(h12, r + s, v[j][d]);
            }
          }
      }
    }
  };
  const double* fac_set = S.mfac + 2 * CRB_SLOT_PAIRS * (M * TG::G) + 2 * CRB_SCAN_PAIRS * TG::LVE * TG::G;
  fast_persistent_run<M, LV, NC>(P, S, fac_set, X, nsteps, Y, save_every, tile_counter, smem, C, tile_setup, step);
}

// ==========================================================================================
// Implicit midpoint rule (Newmark average acceleration) for the same beams: ONE operator
// application and ONE solve with the factors of M + h^2/4 K per step (crb_assemble_shifted),
//     dv = h (M + h^2/4 K)^-1 ( u(t + h/2) - K (q + h/2 v) ),  v+ = v + dv,  q+ = q + h v + h/2 dv
// Unconditionally stable: the step is chosen for accuracy, not for the highest element frequency
// (RK4 needs h < 2.8 / omega_max; the reference's examples use LSODA for that reason,
// examples/example_utilities.py:153-159).  Not a reference code path: parity is against the oracle's
// restatement of the rule on the reference's M and K.
// ==========================================================================================
#ifndef CRB_MID_MINBLOCKS
#define CRB_MID_MINBLOCKS CRB_FAST_MINBLOCKS
#endif
#ifndef CRB_MID_PINS
#define CRB_MID_PINS 4
#endif
template <int M, int LV, bool UC, bool IMP, bool PM, bool NC = false, bool REC = false>
__global__ void __launch_bounds__(CRB_FAST_THREADS, CRB_MID_MINBLOCKS)
crb_midpoint_kernel(KPlan P, crb_system_t S, const double* __restrict__ afac, double* __restrict__ X, double t0,
                    double h, int nsteps, double* __restrict__ Y, int save_every, int stage_off, int stage_stride) {
  extern __shared__ __align__(16) double smem[];
  double* const stage_rows = stage_off ? smem + stage_off : nullptr;  // staging rows of recorded frames (FrameWriter)
  typedef FastCtx<M, CRB_MID_PINS, (CRB_MID_PINS > 0 ? 1 : 0), (NC ? 0 : CRB_MID_PINS), NC> Ctx;  // one solve per step: the constants of the solve stay in registers
  Ctx C;
  const int s0 = fast_ctx_init<M, LV, PM>(C, P, S, afac, smem);
  if (IMP) fast_ctx_impulse<M>(C, S, s0);
  const int n = C.n;
  // reduced index of own DOF (j, d); NC: the plan's table (-1: constrained or phantom, held at zero)
  auto rix = [&](int j, int d) -> int { return NC ? C.ri[NC ? j : 0][d] : 3 * (s0 + j) + d; };
  // recorded frames: the warp's members are consecutive rows of Y (FrameWriter; only the selection mask lives across
  // the step loop, everything else is recomputed when a frame is written)
  FrameWriter<M, NC> FW;
  if (REC) FW.init(S, Y, n, stage_rows ? stage_rows + (threadIdx.x >> 5) * stage_stride : nullptr, rix);
  auto write_frame = [&](long long frame, const double (&fq)[M][3], const double (&fv)[M][3]) {
    constexpr int fw_mpw = 32 >> LV;
    const int fw_m0 = blockIdx.x * (CRB_FAST_WARPS * fw_mpw) + (threadIdx.x >> 5) * fw_mpw;
    FW.stage = stage_rows ? stage_rows + (threadIdx.x >> 5) * stage_stride : nullptr;
    FW.write(S, Y, frame, fw_m0, (threadIdx.x & 31) >> LV, max(0, min(fw_mpw, S.n_members - fw_m0)), n, fq, fv, rix);
  };
  double* xm = X + (long long)C.member * 2 * n;
  double q[M][3], v[M][3], w[M][3], uc[UC ? M : 1][3];
#pragma unroll
  for (int j = 0; j < M; ++j)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int r = rix(j, d);
      const bool ok = !NC || r >= 0;
      q[j][d] = ok ? xm[r] : 0.0;
      v[j][d] = ok ? xm[n + r] : 0.0;
      if (UC) {
        const long long mo = (long long)C.member * n + r;
        uc[j][d] = ok ? (S.u_const ? S.u_const[mo] : 0.0) + (S.f_ext ? S.f_ext[mo] : 0.0) : 0.0;
      }
    }
  const double hh = 0.5 * h;
  for (int k = 0; k < nsteps; ++k) {
    const double tm = t0 + (k + 0.5) * h;  // inputs at the midpoint
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) w[j][d] = fma(hh, v[j][d], q[j][d]);
    fast_accel<M, LV, IMP, Ctx>(C, w, tm, UC ? uc : nullptr);  // w <- (M + h^2/4 K)^-1 (u - K w)
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double dv = h * w[j][d];
        q[j][d] = fma(hh, dv, fma(h, v[j][d], q[j][d]));
        v[j][d] += dv;
      }
    if (REC && Y && save_every > 0 && (k + 1) % save_every == 0) write_frame((k + 1) / save_every - 1, q, v);
  }
  if (C.active) {
#pragma unroll
    for (int j = 0; j < M; ++j)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const int r = rix(j, d);
        if (!NC || r >= 0) {
          xm[r] = q[j][d];
          xm[n + r] = v[j][d];
        }
      }
  }
  if (REC) {
    FW.stage = stage_rows;
    FW.drain();
  }
}
