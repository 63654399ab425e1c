// crb_rk45_steps.cu -- building blocks of the UNFUSED adaptive Dormand-Prince driver.
//
// When the right-hand side contains user code (torch force plug-ins, AbstractForce.compute_forces at
// models/abstractions.py:153-173, or a free-form input callable u(t) as in the reference's own tests,
// tests/test_dynamic_beam.py:201-244), the RHS cannot live inside one kernel.  The host then evaluates the six
// stage derivatives of an attempt with crb_rhs (+ the user's torch code) and everything else stays on the device:
//
//   crb_rk45_stage    Ys = Y + h * sum_{l<s} a[s][l] K[l],  ts = t + c[s] h        (rk.py:14-72 rk_step)
//   crb_rk45_control  error norm, accept / reject, step-size update, dense output at t_eval, commit (FSAL) and
//                     the set-up of the NEXT attempt (min_step test, clipping to t_bound)   (rk.py:111-176)
//
// every member with its own (t, h); there is no host synchronisation inside an attempt.  The controller restates
// SciPy's RK45 (scipy/integrate/_ivp/rk.py, common.py:63-65 norm) like the fused kernel crb_rk45.cuh does.
#include "crb_internal.h"
#include "crb_rk45.cuh"

// K: [7, B, n2] stage derivatives; Ys[b, i] = Y[b, i] + h[b] * sum_l a[s][l] K[l][b, i]
__global__ void crb_rk45_stage_kernel(int n2, int B, int s, const double* __restrict__ Y, const double* __restrict__ K,
                                      const double* __restrict__ t, const double* __restrict__ h_step,
                                      double* __restrict__ Ys, double* __restrict__ ts, DpTab T) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * n2;
  if (idx >= total) return;
  const int b = (int)(idx / n2);
  const double h = h_step[b];
  double acc = 0.0;
  for (int l = 0; l < s; ++l) acc = fma(T.a[s][l], K[(long long)l * total + idx], acc);
  Ys[idx] = fma(h, acc, Y[idx]);
  if (idx % n2 == 0) ts[b] = t[b] + T.c[s] * h;
}

struct Rk45Ctl {
  double* Y;          // [B, n2] current state (updated on accept)
  double* K;          // [7, B, n2]; K[0] = f(t, y) on entry of an attempt, K[6] = f(t + h, y_new)
  const double* Ynew; // [B, n2] y + h sum b_l K_l (stage 6 of crb_rk45_stage)
  double* t;          // [B]
  double* t_next;     // [B] end of the attempt in flight
  double* h_abs;      // [B]
  double* h_step;     // [B] signed step of the attempt in flight (0: member idle)
  double t_bound, rtol, atol;
  const double* t_eval;
  int n_eval;
  int* ie;            // [B] next t_eval index
  double* Y_eval;     // [n_eval, B, n2]
  int* status;        // [B] 0 ok, -1 step too small
  long long* counters;// [B, 3] nfev, accepted, rejected
  int* flags;         // [B] bit 0 running, bit 1 previous attempt rejected, bit 2 first attempt of a step
};

// one warp per member
__global__ void crb_rk45_control_kernel(int n2, int B, Rk45Ctl A, int begin_only, DpTab T) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= B) return;
  const long long total = (long long)B * n2, row = (long long)b * n2;
  int fl = A.flags[b];
  bool running = fl & 1, rejected = fl & 2, new_step = fl & 4;
  double t = A.t[b], h_abs = A.h_abs[b];
  if (!begin_only && running) {
    const double h = A.h_step[b], t_new = A.t_next[b];
    double se = 0.0;
    for (int i = lane; i < n2; i += 32) {
      double e = 0.0;
#pragma unroll
      for (int l = 0; l < 7; ++l) e = fma(T.e[l], A.K[l * total + row + i], e);
      const double sc = fma(fmax(fabs(A.Y[row + i]), fabs(A.Ynew[row + i])), A.rtol, A.atol);
      const double r = e * h / sc;
      se = fma(r, r, se);
    }
    for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(CRB_FULL_MASK, se, d);
    const double err = sqrt(se / n2);
    if (lane == 0) A.counters[3ll * b + 0] += 6;
    if (err < 1.0) {
      double factor = (err == 0.0) ? 10.0 : fmin(10.0, 0.9 * pow(err, -0.2));
      if (rejected) factor = fmin(1.0, factor);
      h_abs *= factor;
      int ie = A.ie[b];
      while (ie < A.n_eval && A.t_eval[ie] <= t_new) {  // dense output (rk.py:178-180, ivp.py side='right')
        const double x = (A.t_eval[ie] - t) / h;
        const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
        double* out = A.Y_eval + ((long long)ie * B + b) * n2;
        for (int i = lane; i < n2; i += 32) {
          double p = 0.0;
#pragma unroll
          for (int l = 0; l < 7; ++l) {
            const double w = fma(T.p[l][0], x, fma(T.p[l][1], x2, fma(T.p[l][2], x3, T.p[l][3] * x4)));
            p = fma(w, A.K[l * total + row + i], p);
          }
          out[i] = fma(h, p, A.Y[row + i]);
        }
        ++ie;
      }
      __syncwarp();
      for (int i = lane; i < n2; i += 32) {  // commit, FSAL
        A.Y[row + i] = A.Ynew[row + i];
        A.K[row + i] = A.K[6 * total + row + i];
      }
      if (lane == 0) {
        A.ie[b] = ie;
        A.counters[3ll * b + 1] += 1;
      }
      t = t_new;
      rejected = false;
      new_step = true;
      if (t >= A.t_bound) running = false;
    } else {
      h_abs *= fmax(0.2, 0.9 * pow(err, -0.2));
      rejected = true;
      if (lane == 0) A.counters[3ll * b + 2] += 1;
    }
  }
  // ---- set up the next attempt (rk.py:111-140) ----
  double h = 0.0, t_new = t;
  if (running) {
    const double min_step = 10.0 * fabs(next_up(t) - t);
    if (new_step && h_abs < min_step) h_abs = min_step;
    new_step = false;
    if (h_abs < min_step) {  // TOO_SMALL_STEP
      if (lane == 0) A.status[b] = -1;
      running = false;
    } else {
      h = h_abs;
      t_new = t + h;
      if (t_new - A.t_bound > 0.0) t_new = A.t_bound;
      h = t_new - t;
      h_abs = fabs(h);
    }
  }
  if (lane == 0) {
    A.t[b] = t;
    A.t_next[b] = t_new;
    A.h_abs[b] = h_abs;
    A.h_step[b] = h;
    A.flags[b] = (running ? 1 : 0) | (rejected ? 2 : 0) | (new_step ? 4 : 0);
  }
}

extern "C" int crb_rk45_stage(int32_t n2, int32_t n_members, int32_t stage, const double* Y, const double* K,
                              const double* t, const double* h_step, double* Ys, double* ts, void* stream) {
  if (!Y || !K || !t || !h_step || !Ys || !ts) return crb_fail(CRB_E_ARG, "crb_rk45_stage: null argument");
  if (n2 < 1 || n_members < 1) return crb_fail(CRB_E_ARG, "crb_rk45_stage: n2 and n_members must be >= 1");
  if (stage < 1 || stage > 6) return crb_fail(CRB_E_ARG, "crb_rk45_stage: stage must be 1..6 (6 = y_new), got %d", stage);
  const long long total = (long long)n_members * n2;
  const int threads = 256;
  const DpTab T = make_dp_tab();
  crb_rk45_stage_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      n2, n_members, stage, Y, K, t, h_step, Ys, ts, T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_rk45_stage: launch failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int crb_rk45_control(int32_t n2, int32_t n_members, double* Y, double* K, const double* Ynew, double* t,
                                double* t_next, double* h_abs, double* h_step, double t_bound, double rtol, double atol,
                                const double* t_eval, int32_t n_eval, int32_t* ie, double* Y_eval, int32_t* status,
                                int64_t* counters, int32_t* flags, int32_t begin_only, void* stream) {
  if (!Y || !K || !Ynew || !t || !t_next || !h_abs || !h_step || !ie || !status || !counters || !flags)
    return crb_fail(CRB_E_ARG, "crb_rk45_control: null argument");
  if (n2 < 1 || n_members < 1) return crb_fail(CRB_E_ARG, "crb_rk45_control: n2 and n_members must be >= 1");
  if (n_eval < 0 || (n_eval > 0 && (!t_eval || !Y_eval))) return crb_fail(CRB_E_ARG, "crb_rk45_control: bad t_eval / Y_eval");
  if (!(rtol > 0.0) || !(atol >= 0.0)) return crb_fail(CRB_E_ARG, "crb_rk45_control: rtol must be > 0 and atol >= 0");
  Rk45Ctl A;
  A.Y = Y; A.K = K; A.Ynew = Ynew; A.t = t; A.t_next = t_next; A.h_abs = h_abs; A.h_step = h_step;
  A.t_bound = t_bound; A.rtol = rtol; A.atol = atol; A.t_eval = t_eval; A.n_eval = n_eval; A.ie = ie;
  A.Y_eval = Y_eval; A.status = status; A.counters = reinterpret_cast<long long*>(counters); A.flags = flags;
  const int threads = 128;  // 4 members per block
  const DpTab T = make_dp_tab();
  crb_rk45_control_kernel<<<(n_members + 3) / 4, threads, 0, (cudaStream_t)stream>>>(n2, n_members, A, begin_only, T);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_rk45_control: launch failed: %s", cudaGetErrorString(e));
  return 0;
}
