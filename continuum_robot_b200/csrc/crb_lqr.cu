// crb_lqr.cu -- batched LQR synthesis on the device: per-member A / B build and the
// continuous-time algebraic Riccati equation, one thread block per design.
//
// Reference behaviour (control/linear_quadratic_regulator.py:84-191, examples/lqr_control.py:46-84):
//   A = [[0, I], [-M^-1 K, 0]],  B = [[0], [M^-1]],  K_gain, S, E = control.lqr(A, B, Q, R),
//   ValueError when the closed loop A - B K_gain has an eigenvalue with Re >= 0.
// control.lqr (python-control / slycot) is a third-party dependency that is absent from this image; its result is
// the unique stabilising solution S of  A^T S + S A - S B R^-1 B^T S + Q = 0  and  K_gain = R^-1 B^T S.
//
// Method (chosen for the GPU: only inversions and products of small dense matrices in shared memory):
//   * sign(H) of the Hamiltonian H = [[A, -G], [-Q, -A^T]], G = B R^-1 B^T, by the Newton iteration
//     Z <- (c Z + (c Z)^-1) / 2 with determinant scaling c = |det Z|^(-1/D) (Roberts 1980; Byers 1987);
//   * the stable invariant subspace is the range of (I - sign H) / 2, whose first block column gives
//     S = W21 (W11 - I)^-1;
//   * iterative refinement (`refine_passes` times): with the residual Res = A^T S + S A - S G S + Q, one
//     Newton-Kleinman step (A - G S)^T dS + dS (A - G S) + Res = 0, solved by the coupled sign iteration on the
//     closed loop (lyapunov_sign_iteration: matrices of half the Hamiltonian's size).  One pass brings the gain from
//     ~1e-6 to ~2e-10 of the Newton-Kleinman fixed point (SciPy's own solve_continuous_are is at ~1.5e-8 on these
//     beams).  (The error equation is itself a Riccati equation; solving it with the Hamiltonian solver gave the
//     same accuracy at 2.7 x the cost of this step.)
//   * closed-loop check: sign(A - G S) = -I  <=>  every closed-loop eigenvalue has Re < 0.
// All inversions are Gauss-Jordan eliminations with row pivoting: register-resident (gj_inverse_reg: the matrix lives
// in the registers of a 16 x 16 thread grid, only the pivot row / column go through shared memory) up to 80 x 80,
// in place in shared memory (gj_inverse) beyond.  Measured on B200, 8192 designs with 72 x 72 Hamiltonians:
// shared-memory elimination 132 ms (LDS/STS wavefronts 64 % of peak), register-resident 110 ms (instruction-issue
// bound: ~290 -> ~200 instructions per thread and pivot for 25 FMAs; 3 blocks of 256 threads per SM).
#include "crb_internal.h"

#define CRB_LQR_THREADS 256
#define CRB_LQR_MAX_ITERS 80

namespace {

struct LqrArgs {
  int n, n_members;
  const double* Mb;
  const double* Kb;
  int m_shared, k_shared;
  const double* Q;
  const double* R;
  int passes;
  double* gain;
  double* S_out;
  double* residual;
  int* status;
  double* ws;
  long long ws_stride;  // doubles per block
  int big;              // 1: the 4n x 4n work matrix lives in the block's global workspace (it does not fit shared memory)
};

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // red may still be read from a previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < CRB_LQR_THREADS / 32; ++w) s += red[w];
  return s;
}

// Elimination step of gj_inverse for pivot p: warp w owns rows w, w + 8, ..., lane l owns columns l + 32 c
// (c < CH): the pivot-row values of the lane's columns stay in registers for all of its rows, so that the inner
// loop is LDS / DFMA / STS.  Column p was zeroed when it was saved, which makes the update of column p
// (a[i][p] <- -col[i] / pivot) the same FMA as every other column (prow[p] = 1).
template <int CH>
__device__ __forceinline__ void gj_eliminate(double* a, int m, int p, const double* col, const double* prow, double ipv) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double pr[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) pr[c] = (lane + 32 * c < m) ? prow[lane + 32 * c] : 0.0;
  for (int i = warp; i < m; i += CRB_LQR_THREADS / 32) {
    double* row = a + i * m + lane;
    if (i == p) {
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (lane + 32 * c < m) row[32 * c] = pr[c] * ipv;
    } else {
      const double f = -col[i] * ipv;
#pragma unroll
      for (int c = 0; c < CH; ++c)
        if (lane + 32 * c < m) row[32 * c] = fma(f, pr[c], row[32 * c]);
    }
  }
}

// In-place inverse of the m x m row-major matrix `a` (shared memory, m <= 192) by Gauss-Jordan elimination with
// partial pivoting.  Returns sum log|pivot| = log|det a|; a zero pivot column sets *singular (block-uniform).
__device__ double gj_inverse(double* a, int m, double* col, double* prow, int* piv, double* red, bool* singular) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks = (m + 31) >> 5;
  double logdet = 0.0;
  *singular = false;
  for (int p = 0; p < m; ++p) {
    if (warp == 0) {  // pivot search in column p
      double best = -1.0;
      int bi = p;
      for (int i = p + lane; i < m; i += 32) {
        const double v = fabs(a[i * m + p]);
        if (v > best) { best = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { piv[p] = bi; red[16] = best; }
    }
    __syncthreads();
    const int r = piv[p];
    const double best = red[16];
    if (!(best > 0.0) || !isfinite(best)) { *singular = true; return logdet; }
    if (r != p) {  // block-uniform
      for (int j = tid; j < m; j += CRB_LQR_THREADS) {
        const double x = a[p * m + j];
        a[p * m + j] = a[r * m + j];
        a[r * m + j] = x;
      }
      __syncthreads();
    }
    const double pv = a[p * m + p];
    __syncthreads();  // every thread has read the pivot before column p is zeroed
    for (int i = tid; i < m; i += CRB_LQR_THREADS) {
      col[i] = a[i * m + p];
      prow[i] = (i == p) ? 1.0 : a[p * m + i];
      a[i * m + p] = 0.0;
    }
    __syncthreads();
    const double ipv = 1.0 / pv;
    switch (chunks) {
      case 1: gj_eliminate<1>(a, m, p, col, prow, ipv); break;
      case 2: gj_eliminate<2>(a, m, p, col, prow, ipv); break;
      case 3: gj_eliminate<3>(a, m, p, col, prow, ipv); break;
      case 4: gj_eliminate<4>(a, m, p, col, prow, ipv); break;
      case 5: gj_eliminate<5>(a, m, p, col, prow, ipv); break;
      case 6: gj_eliminate<6>(a, m, p, col, prow, ipv); break;
      default:  // wider than 192 columns (matrix in global memory): pivot row read from shared memory
        for (int i = warp; i < m; i += CRB_LQR_THREADS / 32) {
          double* row = a + (long long)i * m;
          if (i == p) {
            for (int j = lane; j < m; j += 32) row[j] = prow[j] * ipv;
          } else {
            const double f = -col[i] * ipv;
            for (int j = lane; j < m; j += 32) row[j] = fma(f, prow[j], row[j]);
          }
        }
        break;
    }
    logdet += log(fabs(pv));
    // the next pivot search reads column p + 1 of the updated matrix
    __syncthreads();
  }
  for (int p = m - 1; p >= 0; --p) {  // undo the row exchanges as column exchanges, in reverse
    const int r = piv[p];
    if (r != p) {
      for (int i = tid; i < m; i += CRB_LQR_THREADS) {
        const double x = a[i * m + p];
        a[i * m + p] = a[i * m + r];
        a[i * m + r] = x;
      }
      __syncthreads();
    }
  }
  return logdet;
}

// Register-resident Gauss-Jordan inverse for m <= 16 T (T <= 5, i.e. m <= 80: the 72 x 72 Hamiltonian of the 6-element
// example).  The 256 threads form a 16 x 16 grid, thread (ty, tx) holds the elements (ty + 16 a, tx + 16 b) of the
// matrix in registers for the whole elimination; per pivot only the pivot column and the pivot row travel through
// shared memory (2 T loads per thread for T^2 FMAs, against one load and one store per FMA in gj_inverse).
// Row pivoting is IMPLICIT (no exchanges): the pivot of column p is the largest entry among the rows not used yet,
// piv[p] = its row; at the end inverse(p, piv[c]) = storage(piv[p], c), which is undone when the result is scattered
// back to shared memory.  Padding rows / columns (>= m) are zeros and never become pivots.
template <int T>
__device__ double gj_inverse_reg(double* a, int m, double* colb, double* prowb, int* piv, int* rinv, int* pinfo, bool* singular) {
  const int tid = threadIdx.x, ty = tid & 15, tx = tid >> 4;
  constexpr int LD = 16 * T;
  double W[T][T];
#pragma unroll
  for (int x = 0; x < T; ++x)
#pragma unroll
    for (int y = 0; y < T; ++y) {
      const int i = ty + 16 * x, c = tx + 16 * y;
      W[x][y] = (i < m && c < m) ? a[i * m + c] : 0.0;
    }
  unsigned used = 0;  // bit x: own row ty + 16 x has been a pivot row
  double dmant = 1.0;  // |det| = dmant * 2^dexp, tracked by the thread that publishes the pivots of a column class
  int dexp = 0;
  bool bad = false;
#pragma unroll
  for (int b0 = 0; b0 < T; ++b0) {
    for (int p16 = 0; p16 < 16; ++p16) {
      const int p = 16 * b0 + p16;
      if (p >= m) break;
      const int buf = p & 1;
      double* col = colb + buf * LD;
      double* prow = prowb + buf * LD;
      const bool own_col = tx == p16;
      // phase 1: the owners of column p (one half-warp) publish it and find the pivot row
      double best = -1.0;
      int bi = 0;
#pragma unroll
      for (int x = 0; x < T; ++x) {
        const int i = ty + 16 * x;
        const double v = W[x][b0];
        if (own_col) col[i] = v;
        const double av = (i < m && !((used >> x) & 1u)) ? fabs(v) : -1.0;
        if (av > best) { best = av; bi = i; }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o, 16);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o, 16);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (own_col) {
        if (ty == 0) {
          pinfo[2 * buf] = bi;
          pinfo[2 * buf + 1] = (best > 0.0 && isfinite(best)) ? 1 : 0;
          piv[p] = bi;
          rinv[bi] = p;
          // |det| as mantissa x 2^exponent (a log per pivot costs more than the elimination step of a thread)
          int ex;
          dmant *= frexp(best, &ex);
          dexp += ex;
          int ex2;
          dmant = frexp(dmant, &ex2);
          dexp += ex2;
        }
#pragma unroll
        for (int x = 0; x < T; ++x) W[x][b0] = 0.0;  // column p restarts from zero (it receives the inverse column)
      }
      __syncthreads();
      const int r = pinfo[2 * buf];
      if (!pinfo[2 * buf + 1]) { bad = true; break; }
      // phase 2: the owners of row r publish it (x0 is block-uniform: the branch on it costs no divergence)
      const bool own_row = ty == (r & 15);
      const int x0 = r >> 4;
      if (own_row) used |= 1u << x0;
#pragma unroll
      for (int x = 0; x < T; ++x)
        if (x == x0 && own_row) {
#pragma unroll
          for (int y = 0; y < T; ++y) prow[tx + 16 * y] = W[x][y];
        }
      __syncthreads();
      // phase 3: rank-1 update in registers
      const double ipv = 1.0 / col[r];
      double pr[T];
#pragma unroll
      for (int y = 0; y < T; ++y) pr[y] = (y == b0 && own_col) ? 1.0 : prow[tx + 16 * y];
#pragma unroll
      for (int x = 0; x < T; ++x) {
        const double f = -col[ty + 16 * x] * ipv;
        if (x == x0) {
#pragma unroll
          for (int y = 0; y < T; ++y) W[x][y] = own_row ? pr[y] * ipv : fma(f, pr[y], W[x][y]);
        } else {
#pragma unroll
          for (int y = 0; y < T; ++y) W[x][y] = fma(f, pr[y], W[x][y]);
        }
      }
      // col / prow / pinfo are double-buffered: the next pivot writes the other buffer, and its first barrier
      // orders these reads before the buffer is reused two pivots later
    }
    if (bad) break;
  }
  *singular = bad;
  // log|det|: the 16 publishing threads (ty == 0, one per column class tx) hold partial products
  double part = (ty == 0) ? log(dmant) + dexp * 0.69314718055994530942 : 0.0;
  __syncthreads();  // every thread has loaded its tile, piv / rinv are complete
  if (ty == 0) colb[tx] = part;
  __syncthreads();
  double logdet = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) logdet += colb[k];
  if (!bad) {
#pragma unroll
    for (int x = 0; x < T; ++x)
#pragma unroll
      for (int y = 0; y < T; ++y) {
        const int i = ty + 16 * x, c = tx + 16 * y;
        if (i < m && c < m) a[rinv[i] * m + piv[c]] = W[x][y];
      }
  }
  __syncthreads();
  return logdet;
}

// ------------------------------------------------------------------------------------------
// Blocked Gauss-Jordan inverse for matrices that do not fit shared memory (long beams: the 4n x 4n Hamiltonian of a
// 32-element cantilever is 384 x 384 = 1.2 MB).  The unblocked elimination streams the whole matrix once per pivot
// (384 passes: HBM / L2 bound, 122 designs/s at 32 elements); here the matrix is eliminated in panels of NB columns:
//   1. the panel (all m rows x NB columns) is copied to shared memory (transposed: Wt[c][i]) and eliminated there with
//      implicit row pivoting over the rows not used yet -- afterwards it holds the NB special columns W of the
//      accumulated transformation T = I + (W - E_R) E_R^T (E_R selects the panel's pivot rows R);
//   2. the OLD pivot rows A[R, :] are saved in shared memory (Rr[c][j]);
//   3. every other column is updated with ONE pass over the matrix, a rank-NB product from shared memory:
//      A[i, j] <- (i in R ? 0 : A[i, j]) + sum_c W[i, c] A[r_c, j]        (register tile 4 x 4 per thread);
//   4. the panel is written back.
// m / NB passes over the matrix instead of m, 2 m^3 flops as before.  Implicit pivoting leaves
// inverse(p, piv[c]) = storage(piv[p], c), undone while the result is written to the scratch matrix `tmp`.
// ------------------------------------------------------------------------------------------
#define CRB_LQR_NB 32
__host__ __device__ inline size_t lqr_blk_doubles(int m) {  // Wt | Rr | prow | flag (int) of gj_inverse_blocked
  const size_t mp = ((size_t)m + 7) & ~(size_t)7, mr = ((size_t)m + 127) & ~(size_t)127;
  return CRB_LQR_NB * (mp + mr) + CRB_LQR_NB + (mp + 1) / 2 + 2;
}

// Streaming helpers for the long-beam path: with one block of 256 threads per SM a plain `dst[k] = src[k]` loop keeps
// too few bytes in flight (measured 5 GB/s per SM); 128-bit accesses, eight per thread and trip.
__device__ __forceinline__ void big_copy(double* __restrict__ dst, const double* __restrict__ src, int count) {
  const int tid = threadIdx.x, n2 = count >> 1;
  double2* d2 = reinterpret_cast<double2*>(dst);
  const double2* s2 = reinterpret_cast<const double2*>(src);
  for (int k0 = 0; k0 < n2; k0 += 8 * CRB_LQR_THREADS) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * CRB_LQR_THREADS + tid;
      if (k < n2) v[u] = s2[k];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * CRB_LQR_THREADS + tid;
      if (k < n2) d2[k] = v[u];
    }
  }
  if ((count & 1) && tid == 0) dst[count - 1] = src[count - 1];
}

// Returns log|det a|; the INVERSE IS LEFT IN `tmp` (one pass less than copying it back), `a` is destroyed.
__device__ __noinline__ double gj_inverse_blocked(double* a, double* tmp, int m, double* blk, int* piv, int* pinfo, double* red,
                                                  bool* singular) {
  constexpr int NB = CRB_LQR_NB, T = CRB_LQR_THREADS;
  constexpr int RPT = (4 * CRB_LQR_MAX_N + T - 1) / T;  // rows a thread owns during the panel elimination (i = tid + T k)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int mp = (m + 7) & ~7;      // row stride of Wt: whole 8-row tiles of the update
  const int mr = (m + 127) & ~127;  // row stride of Rr: whole 128-column sweeps of the update, no index clamps
  double* Wt = blk;               // [NB][mp]
  double* Rr = Wt + NB * mp;      // [NB][mr]
  double* prs = Rr + NB * mr;     // [NB] panel part of the pivot row
  int* flag = reinterpret_cast<int*>(prs + NB);  // [mp]: 0 = row not used yet, else 1 + the column it is the pivot row of
  double det_m = 1.0;  // |det| = det_m * 2^det_e (one log at the end instead of one per pivot and thread)
  int det_e = 0;
  *singular = false;
  for (int i = tid; i < mp; i += T) flag[i] = 0;
  for (int i = tid; i < NB * mr; i += T) Rr[i] = 0.0;
  __syncthreads();
  for (int k0 = 0; k0 < m; k0 += NB) {
    const int nb = min(NB, m - k0);
    // ---- 1. panel elimination.  Thread t owns rows t, t + T, ... and keeps its NB panel entries of each IN REGISTERS
    // for the whole panel: a pivot is 2 block barriers, one broadcast row in shared memory and NB register FMAs per
    // owned row (the shared-memory form spent 30 % of the kernel here) ----
    double W[RPT][NB];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {  // the thread's own rows: 256 contiguous bytes each, straight into registers
      const int i = tid + T * k;
      const double* prow_g = a + (long long)i * m + k0;
      if (i < m && nb == NB && (m & 1) == 0) {  // 16-byte aligned (m even, k0 a multiple of NB)
#pragma unroll
        for (int c2 = 0; c2 < NB; c2 += 2) {
          const double2 v = *reinterpret_cast<const double2*>(prow_g + c2);
          W[k][c2] = v.x;
          W[k][c2 + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int c2 = 0; c2 < NB; ++c2) W[k][c2] = (i < m && c2 < nb) ? prow_g[c2] : 0.0;
      }
    }
    bool mine_used[RPT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) mine_used[k] = (tid + T * k < m) ? flag[tid + T * k] != 0 : true;
    // The pivot loop runs 4 pivots per trip on the FIRST four register columns and then rotates the register panel by
    // four columns (static register indices without unrolling 32 pivot bodies: that much code thrashed the
    // instruction cache); after the 8 trips the columns are back in place.
#pragma unroll 1
    for (int o = 0; o < NB / 4; ++o) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = 4 * o + cc;
        if (c < nb) {  // block-uniform
          double best = -1.0;
          int bi = m;
#pragma unroll
          for (int k = 0; k < RPT; ++k)
            if (!mine_used[k]) {
              const double v = fabs(W[k][cc]);
              if (v > best) { best = v; bi = tid + T * k; }
            }
#pragma unroll
          for (int sh = 16; sh > 0; sh >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, sh);
            const int oi = __shfl_down_sync(0xffffffffu, bi, sh);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
          }
          if (lane == 0) {
            red[warp] = best;
            reinterpret_cast<int*>(red + 8)[warp] = bi;
          }
          __syncthreads();  // (1) warp maxima visible
          best = red[0];
          int r = reinterpret_cast<int*>(red + 8)[0];
#pragma unroll
          for (int w = 1; w < T / 32; ++w) {
            const double ob = red[w];
            const int oi = reinterpret_cast<int*>(red + 8)[w];
            if (ob > best || (ob == best && oi < r)) { best = ob; r = oi; }
          }
          if (!(best > 0.0) || !isfinite(best) || r >= m) { *singular = true; return 0.0; }  // block-uniform
#pragma unroll
          for (int k = 0; k < RPT; ++k)
            if (tid + T * k == r) {  // the owner publishes the pivot row (panel part, rotated frame) and retires the row
#pragma unroll
              for (int c2 = 0; c2 < NB; ++c2) prs[c2] = W[k][c2];
              mine_used[k] = true;
              flag[r] = k0 + c + 1;
              piv[k0 + c] = r;
            }
          __syncthreads();  // (2) pivot row visible (red may be rewritten after this point)
          const double pv = prs[cc], ipv = 1.0 / pv;
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            const bool isr = (tid + T * k == r);
            const double fi = isr ? 0.0 : -W[k][cc] * ipv;
#pragma unroll
            for (int c2 = 0; c2 < NB; ++c2) {
              const double pr = prs[c2];
              if (c2 == cc) W[k][c2] = isr ? ipv : fi;
              else W[k][c2] = isr ? pr * ipv : fma(fi, pr, W[k][c2]);
            }
          }
          {
            int e;
            det_m = frexp(det_m * fabs(pv), &e);
            det_e += e;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < RPT; ++k) {  // rotate the register panel left by four columns
        const double t0 = W[k][0], t1 = W[k][1], t2 = W[k][2], t3 = W[k][3];
#pragma unroll
        for (int c2 = 0; c2 < NB - 4; ++c2) W[k][c2] = W[k][c2 + 4];
        W[k][NB - 4] = t0;
        W[k][NB - 3] = t1;
        W[k][NB - 2] = t2;
        W[k][NB - 1] = t3;
      }
    }
    // the eliminated panel, transposed, for the update (Wt[c][i])
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int i = tid + T * k;
      if (i < m) {
#pragma unroll
        for (int c2 = 0; c2 < NB; ++c2) Wt[c2 * mp + i] = W[k][c2];
      }
    }
    __syncthreads();
    // ---- 2. the panel's pivot rows as they are BEFORE this block step ----
    for (int c = warp; c < nb; c += T / 32) {
      const double* srow = a + (long long)piv[k0 + c] * m;
      if ((m & 1) == 0) {  // rows are 16-byte aligned
        const double2* src = reinterpret_cast<const double2*>(srow);
        double2* dst = reinterpret_cast<double2*>(Rr + c * mr);
        for (int j = lane; j < (m >> 1); j += 32) dst[j] = src[j];
      } else {
        for (int j = lane; j < m; j += 32) Rr[c * mr + j] = srow[j];
      }
    }
    __syncthreads();
    // ---- 3. rank-nb update of every column outside the panel: warp = 8 rows, lane = 4 columns (stride 32); the
    // next tile's matrix entries are loaded while the current tile is being multiplied ----
    {
      constexpr int TR = 8;
      const int n_jb = mr / 128, rows_per_sweep = TR * (T / 32);
      const int n_ib = (m + rows_per_sweep - 1) / rows_per_sweep, n_tiles = n_ib * n_jb;
      double nxt[TR][4];
      auto load_tile = [&](int tile, double (&dst)[TR][4]) {
        const int i0 = (tile / n_jb) * rows_per_sweep + TR * warp, jb = (tile % n_jb) * 128;
#pragma unroll
        for (int ii = 0; ii < TR; ++ii) {
          const int i = i0 + ii;
          const int fl = (i < m) ? flag[i] : 0;
          const bool isp = fl > k0 && fl <= k0 + nb;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = jb + lane + 32 * jj;
            dst[ii][jj] = (i < m && j < m && !isp) ? a[(long long)i * m + j] : 0.0;
          }
        }
      };
      load_tile(0, nxt);
      for (int tile = 0; tile < n_tiles; ++tile) {
        const int i0 = (tile / n_jb) * rows_per_sweep + TR * warp, jb = (tile % n_jb) * 128;
        double acc[TR][4];
#pragma unroll
        for (int ii = 0; ii < TR; ++ii)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) acc[ii][jj] = nxt[ii][jj];
        if (tile + 1 < n_tiles) load_tile(tile + 1, nxt);
        if (i0 < m) {
          const double* wp = Wt + i0;  // (i0 is a multiple of 8 <= mp - 8; rows past m hold garbage that is never stored)
          const double* rp = Rr + jb + lane;
#pragma unroll 4
          for (int c = 0; c < nb; ++c) {
            const double4 w0 = *reinterpret_cast<const double4*>(wp + c * mp);      // broadcast loads
            const double4 w1 = *reinterpret_cast<const double4*>(wp + c * mp + 4);
            double rv[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) rv[jj] = rp[c * mr + 32 * jj];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              acc[0][jj] = fma(w0.x, rv[jj], acc[0][jj]);
              acc[1][jj] = fma(w0.y, rv[jj], acc[1][jj]);
              acc[2][jj] = fma(w0.z, rv[jj], acc[2][jj]);
              acc[3][jj] = fma(w0.w, rv[jj], acc[3][jj]);
              acc[4][jj] = fma(w1.x, rv[jj], acc[4][jj]);
              acc[5][jj] = fma(w1.y, rv[jj], acc[5][jj]);
              acc[6][jj] = fma(w1.z, rv[jj], acc[6][jj]);
              acc[7][jj] = fma(w1.w, rv[jj], acc[7][jj]);
            }
          }
#pragma unroll
          for (int ii = 0; ii < TR; ++ii)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int i = i0 + ii, j = jb + lane + 32 * jj;
              if (i < m && j < m && (j < k0 || j >= k0 + nb)) a[(long long)i * m + j] = acc[ii][jj];
            }
        }
      }
    }
    __syncthreads();
    // ---- 4. panel back ----
    for (int idx = tid; idx < nb * m; idx += T) {
      const int i = idx / nb, c = idx - i * nb;
      a[(long long)i * m + k0 + c] = Wt[c * mp + i];
    }
    __syncthreads();
  }
  // implicit pivoting: inverse(p, piv[c]) = storage(piv[p], c); one row per warp, coalesced reads
  for (int p = warp; p < m; p += T / 32) {
    const double* src = a + (long long)piv[p] * m;
    double* dst = tmp + (long long)p * m;
    for (int c0 = 0; c0 < m; c0 += 128) {
      double v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < m) v[u] = src[c];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + lane + 32 * u;
        if (c < m) dst[piv[c]] = v[u];
      }
    }
  }
  __syncthreads();
  return log(det_m) + det_e * 0.69314718055994530942;
}

// The inverse is left in `a`, or (BIG) in `gtmp`: inv_of<BIG>(a, gtmp) names the place.
template <bool BIG>
__device__ __forceinline__ double* inv_of(double* a, double* gtmp) { return BIG ? gtmp : a; }
template <bool BIG = false>
__device__ double gj_inverse_any(double* a, int m, double* col, double* prow, int* piv, int* rinv, int* pinfo, double* red,
                                 bool* singular, double* blk = nullptr, double* gtmp = nullptr) {
  if (BIG) return gj_inverse_blocked(a, gtmp, m, blk, piv, pinfo, red, singular);
  switch ((m + 15) >> 4) {
    case 1: return gj_inverse_reg<1>(a, m, col, prow, piv, rinv, pinfo, singular);
    case 2: return gj_inverse_reg<2>(a, m, col, prow, piv, rinv, pinfo, singular);
    case 3: return gj_inverse_reg<3>(a, m, col, prow, piv, rinv, pinfo, singular);
    case 4: return gj_inverse_reg<4>(a, m, col, prow, piv, rinv, pinfo, singular);
    case 5: return gj_inverse_reg<5>(a, m, col, prow, piv, rinv, pinfo, singular);
    default: return gj_inverse(a, m, col, prow, piv, red, singular);
  }
}

// Z (global, m x m) <- sign(Z).  Returns the number of iterations, or -1 (singular iterate / no convergence).
template <bool BIG>
__device__ int sign_iteration(double* Z, int m, double* sm, double* col, double* prow, int* piv, int* rinv, int* pinfo, double* red,
                              double* blk, double* gtmp) {
  const int tid = threadIdx.x;
  double dprev = 1e300;
  bool scaling = true;
  for (int it = 1; it <= CRB_LQR_MAX_ITERS; ++it) {
    if (BIG) big_copy(sm, Z, m * m);
    else
      for (int k = tid; k < m * m; k += CRB_LQR_THREADS) sm[k] = Z[k];
    __syncthreads();
    bool singular;
    const double logdet = gj_inverse_any<BIG>(sm, m, col, prow, piv, rinv, pinfo, red, &singular, blk, gtmp);
    if (singular) return -1;
    const double c = scaling ? exp(-logdet / m) : 1.0, ic = 1.0 / c;
    const double* inv = inv_of<BIG>(sm, gtmp);
    double dd = 0.0, nn = 0.0;
    if (BIG && ((m * m) & 1) == 0) {  // 128-bit accesses, four per thread and trip (bytes in flight, see big_copy)
      double2* Z2 = reinterpret_cast<double2*>(Z);
      const double2* I2 = reinterpret_cast<const double2*>(inv);
      const int cnt = (m * m) >> 1;
      for (int k0 = 0; k0 < cnt; k0 += 4 * CRB_LQR_THREADS) {
        double2 zv[4], iv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + u * CRB_LQR_THREADS + tid;
          if (k < cnt) { zv[u] = Z2[k]; iv[u] = I2[k]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = k0 + u * CRB_LQR_THREADS + tid;
          if (k < cnt) {
            const double z0 = 0.5 * fma(c, zv[u].x, ic * iv[u].x), z1 = 0.5 * fma(c, zv[u].y, ic * iv[u].y);
            const double e0 = z0 - zv[u].x, e1 = z1 - zv[u].y;
            dd = fma(e0, e0, fma(e1, e1, dd));
            nn = fma(z0, z0, fma(z1, z1, nn));
            Z2[k] = make_double2(z0, z1);
          }
        }
      }
    } else {
      for (int k = tid; k < m * m; k += CRB_LQR_THREADS) {
        const double z = Z[k];
        const double zn = 0.5 * fma(c, z, ic * inv[k]);
        const double e = zn - z;
        dd = fma(e, e, dd);
        nn = fma(zn, zn, nn);
        Z[k] = zn;
      }
    }
    dd = block_sum(dd, red);
    nn = block_sum(nn, red);
    if (!isfinite(nn) || !(nn > 0.0)) return -1;
    const double d = sqrt(dd / nn);
    if (d < 1e-14) return it;
    if (dprev < 1e-6 && d >= dprev) return it;  // quadratic phase over: rounding level reached
    if (d < 1e-2) scaling = false;
    dprev = d;
  }
  return -1;
}

// Correction pass = one Newton-Kleinman step: Ac^T dS + dS Ac + W = 0 for the (stable) closed loop Ac, solved by the
// coupled sign iteration (Roberts 1980)
//     A <- (c A + A^-1 / c) / 2,   W <- (c W + A^-T W A^-1 / c) / 2,   dS = W_inf / 2
// on m x m matrices: one inversion of the closed loop plus two products per iteration, against one inversion of the
// twice as large Hamiltonian for a full Riccati solve.  A (global, in: Ac, destroyed), W (global, in: residual, out:
// 2 dS), Tm: scratch.  Returns the iteration count or -1.
template <bool BIG>
__device__ int lyapunov_sign_iteration(double* A, double* W, double* Tm, int m, double* sm, double* col, double* prow, int* piv,
                                       int* rinv, int* pinfo, double* red, double* blk, double* gtmp) {
  const int tid = threadIdx.x;
  double dprev = 1e300;
  bool scaling = true;
  for (int it = 1; it <= CRB_LQR_MAX_ITERS; ++it) {
    for (int k = tid; k < m * m; k += CRB_LQR_THREADS) sm[k] = A[k];
    __syncthreads();
    bool singular;
    const double logdet = gj_inverse_any<BIG>(sm, m, col, prow, piv, rinv, pinfo, red, &singular, blk, gtmp);
    if (singular) return -1;
    const double c = scaling ? exp(-logdet / m) : 1.0, ic = 1.0 / c;
    const double* inv = inv_of<BIG>(sm, gtmp);
    // Tm = W A^-1
    for (int idx = tid; idx < m * m; idx += CRB_LQR_THREADS) {
      const int i = idx / m, j = idx - i * m;
      double acc = 0.0;
      for (int k = 0; k < m; ++k) acc = fma(W[i * m + k], inv[k * m + j], acc);
      Tm[idx] = acc;
    }
    __syncthreads();
    // W <- (c W + A^-T Tm / c) / 2,  A <- (c A + A^-1 / c) / 2
    double dd = 0.0, nn = 0.0;
    for (int idx = tid; idx < m * m; idx += CRB_LQR_THREADS) {
      const int i = idx / m, j = idx - i * m;
      double acc = 0.0;
      for (int k = 0; k < m; ++k) acc = fma(inv[k * m + i], Tm[k * m + j], acc);
      W[idx] = 0.5 * fma(c, W[idx], ic * acc);
      const double z = A[idx];
      const double zn = 0.5 * fma(c, z, ic * inv[idx]);
      const double e = zn - z;
      dd = fma(e, e, dd);
      nn = fma(zn, zn, nn);
      A[idx] = zn;
    }
    dd = block_sum(dd, red);
    nn = block_sum(nn, red);
    if (!isfinite(nn) || !(nn > 0.0)) return -1;
    const double d = sqrt(dd / nn);
    if (d < 1e-14) return it;
    if (dprev < 1e-6 && d >= dprev) return it;
    if (d < 1e-2) scaling = false;
    dprev = d;
  }
  return -1;
}

// C(i,j) = sum_k A(i,k) B(k,j), i < m, j < n (accessors are lambdas; C is written through `store`)
template <typename FA, typename FB, typename FS>
__device__ __forceinline__ void mat_mul(int m, int n, int kk, FA a, FB b, FS store) {
  for (int idx = threadIdx.x; idx < m * n; idx += CRB_LQR_THREADS) {
    const int i = idx / n, j = idx - i * n;
    double acc = 0.0;
    for (int k = 0; k < kk; ++k) acc = fma(a(i, k), b(k, j), acc);
    store(i, j, acc);
  }
}

// Res = A^T X + X A - X G X + Q (n2 x n2, all dense); T, T2 scratch
__device__ void riccati_residual(int n2, const double* A, const double* G, const double* Q, const double* X, double* T,
                                 double* T2, double* Res) {
  mat_mul(n2, n2, n2, [&](int i, int k) { return X[i * n2 + k]; }, [&](int k, int j) { return A[k * n2 + j]; },
          [&](int i, int j, double v) { T[i * n2 + j] = v; });
  mat_mul(n2, n2, n2, [&](int i, int k) { return G[i * n2 + k]; }, [&](int k, int j) { return X[k * n2 + j]; },
          [&](int i, int j, double v) { T2[i * n2 + j] = v; });
  __syncthreads();
  mat_mul(n2, n2, n2, [&](int i, int k) { return X[i * n2 + k]; }, [&](int k, int j) { return T2[k * n2 + j]; },
          [&](int i, int j, double v) { Res[i * n2 + j] = T[i * n2 + j] + T[j * n2 + i] - v + Q[i * n2 + j]; });
  __syncthreads();
}

// BIG: the 4n x 4n work matrix does not fit shared memory (n > 42): it lives in the block's global workspace and is
// inverted by the blocked elimination (gj_inverse_blocked), whose panels take the block's shared memory (one block per SM).
template <bool BIG>
__global__ void __launch_bounds__(CRB_LQR_THREADS, BIG ? 1 : 3) crb_lqr_kernel(LqrArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int n = a.n, n2 = 2 * n, D = 4 * n, tid = threadIdx.x;
  double* ws = a.ws + (long long)blockIdx.x * a.ws_stride;
  // D x D work matrix: shared memory up to n = 42, else the tail of the block's global workspace, with a second
  // D x D scratch matrix in front of it (un-permutation of the blocked inverse)
  double* sm = BIG ? ws + (a.ws_stride - (long long)D * D) : smem;
  double* gtmp = BIG ? sm - (long long)D * D : nullptr;
  double* col = BIG ? smem : smem + D * D;  // 2 (D + 16): double-buffered pivot column (register-resident elimination)
  double* prow = col + 2 * (D + 16);  // 2 (D + 16): pivot row
  double* red = prow + 2 * (D + 16);  // 32
  int* piv = reinterpret_cast<int*>(red + 32);  // D
  int* rinv = piv + D;             // D
  int* pinfo = rinv + D;           // 4
  double* blk = BIG ? reinterpret_cast<double*>(pinfo + 8 + ((2 * D) & 1)) : nullptr;  // panels of the blocked elimination
  double* Z = ws;                 // D^2
  double* A = Z + D * D;          // n2^2 each
  double* G = A + n2 * n2;
  double* X = G + n2 * n2;
  double* Ac = X + n2 * n2;
  double* Qc = Ac + n2 * n2;
  double* T = Qc + n2 * n2;
  double* T2 = T + n2 * n2;
  double* Minv = T2 + n2 * n2;    // n^2 each
  double* Rinv = Minv + n * n;
  double* tmp = Rinv + n * n;
  bool singular;

  // R^-1 once per block
  for (int k = tid; k < n * n; k += CRB_LQR_THREADS) sm[k] = a.R[k];
  __syncthreads();
  gj_inverse_any<BIG>(sm, n, col, prow, piv, rinv, pinfo, red, &singular, blk, gtmp);
  const bool r_singular = singular;
  for (int k = tid; k < n * n; k += CRB_LQR_THREADS) Rinv[k] = inv_of<BIG>(sm, gtmp)[k];
  __syncthreads();

  for (int member = blockIdx.x; member < a.n_members; member += gridDim.x) {
    const double* Mb = a.Mb + (a.m_shared ? 0ll : (long long)member * n * n);
    const double* Kb = a.Kb + (a.k_shared ? 0ll : (long long)member * n * n);
    int status = 0;
    double resid = 0.0;
    // M^-1 (linear_quadratic_regulator.py:100-104)
    __syncthreads();
    for (int k = tid; k < n * n; k += CRB_LQR_THREADS) sm[k] = Mb[k];
    __syncthreads();
    gj_inverse_any<BIG>(sm, n, col, prow, piv, rinv, pinfo, red, &singular, blk, gtmp);
    if (singular || r_singular) status = 1;
    for (int k = tid; k < n * n; k += CRB_LQR_THREADS) Minv[k] = inv_of<BIG>(sm, gtmp)[k];
    __syncthreads();
    if (status == 0) {
      // A = [[0, I], [-M^-1 K, 0]]   (:84-118);   G = B R^-1 B^T = [[0, 0], [0, M^-1 R^-1 M^-T]]   (:120-146)
      for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) {
        const int i = k / n2, j = k - i * n2;
        A[k] = (i < n && j == i + n) ? 1.0 : 0.0;
        G[k] = 0.0;
        X[k] = 0.0;
        Qc[k] = a.Q[k];
      }
      __syncthreads();
      mat_mul(n, n, n, [&](int i, int k) { return Minv[i * n + k]; }, [&](int k, int j) { return Kb[k * n + j]; },
              [&](int i, int j, double v) { A[(n + i) * n2 + j] = -v; });
      mat_mul(n, n, n, [&](int i, int k) { return Minv[i * n + k]; }, [&](int k, int j) { return Rinv[k * n + j]; },
              [&](int i, int j, double v) { tmp[i * n + j] = v; });
      __syncthreads();
      mat_mul(n, n, n, [&](int i, int k) { return tmp[i * n + k]; }, [&](int k, int j) { return Minv[j * n + k]; },
              [&](int i, int j, double v) { G[(n + i) * n2 + n + j] = v; });
      __syncthreads();
      for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) Ac[k] = A[k];
      __syncthreads();

      for (int pass = 0; pass <= a.passes && status == 0; ++pass) {
        if (pass > 0) {
          // correction: Newton-Kleinman step on the closed loop of the current solution (Qc holds its residual)
          for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) Z[k] = Ac[k];
          __syncthreads();
          if (lyapunov_sign_iteration<BIG>(Z, Qc, T, n2, sm, col, prow, piv, rinv, pinfo, red, blk, gtmp) < 0) { status = 2; break; }
          __syncthreads();
          for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) {
            const int i = k / n2, j = k - i * n2;
            X[k] += 0.25 * (Qc[k] + Qc[j * n2 + i]);
          }
          __syncthreads();
          mat_mul(n2, n2, n2, [&](int i, int k) { return G[i * n2 + k]; }, [&](int k, int j) { return X[k * n2 + j]; },
                  [&](int i, int j, double v) { Ac[i * n2 + j] = A[i * n2 + j] - v; });
          __syncthreads();
          riccati_residual(n2, A, G, a.Q, X, T, T2, Qc);
          continue;
        }
        // H = [[Ac, -G], [-Qc, -Ac^T]]
        for (int k = tid; k < D * D; k += CRB_LQR_THREADS) {
          const int i = k / D, j = k - i * D;
          double v;
          if (i < n2) v = (j < n2) ? Ac[i * n2 + j] : -G[i * n2 + (j - n2)];
          else v = (j < n2) ? -Qc[(i - n2) * n2 + j] : -Ac[(j - n2) * n2 + (i - n2)];
          Z[k] = v;
        }
        __syncthreads();
        if (sign_iteration<BIG>(Z, D, sm, col, prow, piv, rinv, pinfo, red, blk, gtmp) < 0) { status = 2; break; }
        __syncthreads();
        // dS = W21 (W11 - I)^-1
        for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) {
          const int i = k / n2, j = k - i * n2;
          sm[k] = Z[i * D + j] - (i == j ? 1.0 : 0.0);
        }
        __syncthreads();
        gj_inverse_any<BIG>(sm, n2, col, prow, piv, rinv, pinfo, red, &singular, blk, gtmp);
        if (singular) { status = 2; break; }
        const double* xinv = inv_of<BIG>(sm, gtmp);
        mat_mul(n2, n2, n2, [&](int i, int k) { return Z[(n2 + i) * D + k]; }, [&](int k, int j) { return xinv[k * n2 + j]; },
                [&](int i, int j, double v) { T[i * n2 + j] = v; });
        __syncthreads();
        for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) {
          const int i = k / n2, j = k - i * n2;
          X[k] += 0.5 * (T[k] + T[j * n2 + i]);
        }
        __syncthreads();
        // closed loop and residual of the updated solution: next pass solves for the correction
        mat_mul(n2, n2, n2, [&](int i, int k) { return G[i * n2 + k]; }, [&](int k, int j) { return X[k * n2 + j]; },
                [&](int i, int j, double v) { Ac[i * n2 + j] = A[i * n2 + j] - v; });
        __syncthreads();
        riccati_residual(n2, A, G, a.Q, X, T, T2, Qc);
      }
    }
    if (status == 0) {
      // relative residual ||Res||_F / ||Q||_F of the returned solution
      double rr = 0.0, qq = 0.0;
      for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) {
        rr = fma(Qc[k], Qc[k], rr);
        qq = fma(a.Q[k], a.Q[k], qq);
      }
      rr = block_sum(rr, red);
      qq = block_sum(qq, red);
      resid = sqrt(rr / qq);
      // closed-loop eigenvalues (:185-189): sign(A - B K) must be -I
      for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) Z[k] = Ac[k];
      __syncthreads();
      if (sign_iteration<BIG>(Z, n2, sm, col, prow, piv, rinv, pinfo, red, blk, gtmp) < 0) {
        status = 3;
      } else {
        double tr = 0.0;
        for (int i = tid; i < n2; i += CRB_LQR_THREADS) tr += Z[i * n2 + i];
        tr = block_sum(tr, red);
        if (fabs(tr + n2) > 0.5) status = 3;
      }
      // K_gain = R^-1 B^T S = R^-1 M^-T S[n:, :]
      mat_mul(n, n2, n, [&](int i, int k) { return Minv[k * n + i]; }, [&](int k, int j) { return X[(n + k) * n2 + j]; },
              [&](int i, int j, double v) { T[i * n2 + j] = v; });
      __syncthreads();
      double* gout = a.gain + (long long)member * n * n2;
      mat_mul(n, n2, n, [&](int i, int k) { return Rinv[i * n + k]; }, [&](int k, int j) { return T[k * n2 + j]; },
              [&](int i, int j, double v) { gout[i * n2 + j] = v; });
      if (a.S_out)
        for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) a.S_out[(long long)member * n2 * n2 + k] = X[k];
    } else {
      const double qnan = __longlong_as_double(0x7ff8000000000000ll);
      for (int k = tid; k < n * n2; k += CRB_LQR_THREADS) a.gain[(long long)member * n * n2 + k] = qnan;
      if (a.S_out)
        for (int k = tid; k < n2 * n2; k += CRB_LQR_THREADS) a.S_out[(long long)member * n2 * n2 + k] = qnan;
      resid = qnan;
    }
    if (tid == 0) {
      a.status[member] = status;
      if (a.residual) a.residual[member] = resid;
    }
    __syncthreads();
  }
}

inline size_t lqr_smem_small(int n) {  // everything but the D x D work matrix
  const size_t D = 4 * (size_t)n;
  return sizeof(double) * (4 * (D + 16) + 32) + sizeof(int) * (2 * D + 8);
}
inline bool lqr_big(int n) { return lqr_smem_small(n) + sizeof(double) * 16 * (size_t)n * n > 227 * 1024; }
inline size_t lqr_smem_bytes(int n) {
  return lqr_smem_small(n) + 16 + sizeof(double) * (lqr_big(n) ? lqr_blk_doubles(4 * n) : 16 * (size_t)n * n);
}
inline long long lqr_ws_doubles(int n) {  // (even: the two D x D matrices at the tail of a block's workspace stay 16-byte aligned)
  return (16ll * n * n + 7 * 4ll * n * n + 3ll * n * n + (lqr_big(n) ? 32ll * n * n : 0) + 1) & ~1ll;
}

int lqr_grid(int n, int n_members, int* out) {
  int dev = 0, sms = 0, per_sm = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_lqr: %s", cudaGetErrorString(e));
  const bool big = lqr_big(n);
  if (int rc = big ? set_smem(crb_lqr_kernel<true>, lqr_smem_bytes(n), "crb_lqr_gains")
                   : set_smem(crb_lqr_kernel<false>, lqr_smem_bytes(n), "crb_lqr_gains"))
    return rc;
  e = big ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, crb_lqr_kernel<true>, CRB_LQR_THREADS, lqr_smem_bytes(n))
          : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, crb_lqr_kernel<false>, CRB_LQR_THREADS, lqr_smem_bytes(n));
  if (e != cudaSuccess || per_sm < 1) return crb_fail(CRB_E_CUDA, "crb_lqr: occupancy query failed (%s)", cudaGetErrorString(e));
  const long long cap = (long long)sms * per_sm;
  *out = (int)(n_members < cap ? n_members : cap);
  return 0;
}

// ------------------------------------------------------------------------------------------
// batched dense M, K (device): one thread per (set, element)
// ------------------------------------------------------------------------------------------
struct RedMap { int r[3 * (CRB_LQR_MAX_ELEMENTS + 1)]; };

__global__ void crb_dense_batched_kernel(int N, int n, int n_sets, const double* __restrict__ params, RedMap map,
                                         double* __restrict__ Mo, double* __restrict__ Ko) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_sets * N) return;
  const int set = (int)(gid / N), e = (int)(gid - (long long)set * N);
  const double* q = params + gid * CRB_NPARAM;
  const double L = q[CRB_P_LENGTH], mu = q[CRB_P_RHO] * q[CRB_P_AREA] * L / 420;
  const double EI = q[CRB_P_E] * q[CRB_P_I], EA = q[CRB_P_E] * q[CRB_P_AREA];
  const double ax = EA / L, c1 = EI / L, c2 = EI / (L * L), c3 = EI / (L * L * L);
  const double L2 = L * L;
  // models/segments.py:32-78 (DOF order u1, w1, phi1, u2, w2, phi2)
  const double me[6][6] = {{140, 0, 0, 70, 0, 0},
                           {0, 156, -22 * L, 0, 54, 13 * L},
                           {0, -22 * L, 4 * L2, 0, -13 * L, -3 * L2},
                           {70, 0, 0, 140, 0, 0},
                           {0, 54, -13 * L, 0, 156, 22 * L},
                           {0, 13 * L, -3 * L2, 0, 22 * L, 4 * L2}};
  const double ke[6][6] = {{ax, 0, 0, -ax, 0, 0},
                           {0, 12 * c3, -6 * c2, 0, -12 * c3, -6 * c2},
                           {0, -6 * c2, 4 * c1, 0, 6 * c2, 2 * c1},
                           {-ax, 0, 0, ax, 0, 0},
                           {0, -12 * c3, 6 * c2, 0, 12 * c3, 6 * c2},
                           {0, -6 * c2, 2 * c1, 0, 6 * c2, 4 * c1}};
  double* Ms = Mo + (long long)set * n * n;
  double* Ks = Ko ? Ko + (long long)set * n * n : nullptr;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      const int ri = map.r[3 * e + i], rj = map.r[3 * e + j];
      if (ri < 0 || rj < 0) continue;
      // neighbouring elements add into the shared node block: two addends per entry at most, so the
      // sum does not depend on the order of the atomics
      atomicAdd(Ms + (long long)ri * n + rj, me[i][j] * mu);
      if (Ks) atomicAdd(Ks + (long long)ri * n + rj, ke[i][j]);
    }
}

// ------------------------------------------------------------------------------------------
// per-member closed-loop operators (one block per member, grid-stride)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CRB_LQR_THREADS, 3)
crb_member_operator_kernel(int n, int n_members, const double* __restrict__ Mb, int m_shared, const double* __restrict__ Kb,
                           int k_shared, const double* __restrict__ gain, const double* __restrict__ ref,
                           double* __restrict__ op, int* __restrict__ status) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, n2 = 2 * n, RL = 3 * n + 1;
  double* sm = smem;                    // n x n: M^-1
  double* col = sm + n * n;             // 2 (n + 16)
  double* prow = col + 2 * (n + 16);    // 2 (n + 16)
  double* red = prow + 2 * (n + 16);    // 32
  double* gr = red + 32;                // n: gain . ref
  int* piv = reinterpret_cast<int*>(gr + n);
  int* rinv = piv + n;
  int* pinfo = rinv + n;
  for (int member = blockIdx.x; member < n_members; member += gridDim.x) {
    const double* M = Mb + (m_shared ? 0ll : (long long)member * n * n);
    const double* K = Kb + (k_shared ? 0ll : (long long)member * n * n);
    const double* G = gain + (long long)member * n * n2;
    double* out = op + (long long)member * n * RL;
    __syncthreads();
    for (int k = tid; k < n * n; k += CRB_LQR_THREADS) sm[k] = M[k];
    for (int i = tid; i < n; i += CRB_LQR_THREADS) {
      double acc = 0.0;
      if (ref)
        for (int c = 0; c < n2; ++c) acc = fma(G[i * n2 + c], ref[c], acc);
      gr[i] = acc;
    }
    __syncthreads();
    bool singular;
    gj_inverse_any(sm, n, col, prow, piv, rinv, pinfo, red, &singular);
    if (tid == 0) status[member] = singular ? 1 : 0;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    for (int idx = tid; idx < n * RL; idx += CRB_LQR_THREADS) {
      const int l = idx / RL, j = idx - l * RL;
      double v;
      if (singular) {
        v = qnan;
      } else if (j < n2) {  // -M^-1 (K + G_q) | -M^-1 G_v
        double acc = 0.0;
        for (int k = 0; k < n; ++k) {
          const double g = G[k * n2 + j] + (j < n ? K[k * n + j] : 0.0);
          acc = fma(sm[l * n + k], g, acc);
        }
        v = -acc;
      } else if (j < 3 * n) {  // M^-1
        v = sm[l * n + (j - n2)];
      } else {  // M^-1 G ref
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc = fma(sm[l * n + k], gr[k], acc);
        v = acc;
      }
      out[idx] = v;
    }
  }
}

}  // namespace

extern "C" int crb_dense_matrices_batched(const crb_plan_t* plan, const double* params, int32_t n_param_sets,
                                          const uint8_t* elem_type_host, const uint8_t* bc_host, double* M_out,
                                          double* K_out, void* stream) {
  if (!plan || !params || !elem_type_host || !bc_host || !M_out)
    return crb_fail(CRB_E_ARG, "crb_dense_matrices_batched: null argument");
  if (n_param_sets < 1) return crb_fail(CRB_E_ARG, "crb_dense_matrices_batched: n_param_sets must be positive");
  const int N = plan->n_elements, n = plan->n_free;
  if (N > CRB_LQR_MAX_ELEMENTS)
    return crb_fail(CRB_E_LIMIT, "crb_dense_matrices_batched: at most %d elements (got %d)", CRB_LQR_MAX_ELEMENTS, N);
  if (K_out)
    for (int e = 0; e < N; ++e)
      if (elem_type_host[e] != CRB_ELEM_LINEAR)
        return crb_fail(CRB_E_ARG, "Cannot extract stiffness matrix from beam with nonlinear segments. Segment %d is nonlinear.", e);
  RedMap map;
  int r = 0;
  for (int node = 0; node <= N; ++node)
    for (int d = 0; d < 3; ++d) {
      const bool c = bc_host[node] == CRB_BC_FIXED || (bc_host[node] == CRB_BC_PINNED && d < 2);
      map.r[3 * node + d] = c ? -1 : r++;
    }
  if (r != n) return crb_fail(CRB_E_ARG, "crb_dense_matrices_batched: plan / bc mismatch (%d vs %d free DOFs)", r, n);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = sizeof(double) * (size_t)n_param_sets * n * n;
  cudaError_t e = cudaMemsetAsync(M_out, 0, bytes, st);
  if (e == cudaSuccess && K_out) e = cudaMemsetAsync(K_out, 0, bytes, st);
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_dense_matrices_batched: %s", cudaGetErrorString(e));
  const long long work = (long long)n_param_sets * N;
  crb_dense_batched_kernel<<<(unsigned)((work + 127) / 128), 128, 0, st>>>(N, n, n_param_sets, params, map, M_out, K_out);
  e = cudaGetLastError();
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_dense_matrices_batched: %s", cudaGetErrorString(e));
  return 0;
}

static int lqr_check_n(int32_t n, const char* who) {
  if (n < 1) return crb_fail(CRB_E_ARG, "%s: n must be positive", who);
  if (n > CRB_LQR_MAX_N)
    return crb_fail(CRB_E_LIMIT, "%s: n = %d free DOFs exceed the limit of %d (%d x %d Hamiltonian)", who, n, CRB_LQR_MAX_N, 4 * n, 4 * n);
  return 0;
}

extern "C" int crb_lqr_workspace_bytes(int32_t n, int32_t n_members, size_t* out) {
  if (!out) return crb_fail(CRB_E_ARG, "crb_lqr_workspace_bytes: null out");
  if (int rc = lqr_check_n(n, "crb_lqr_workspace_bytes")) return rc;
  if (n_members < 1) return crb_fail(CRB_E_ARG, "crb_lqr_workspace_bytes: n_members must be positive");
  int grid = 0;
  if (int rc = lqr_grid(n, n_members, &grid)) return rc;
  *out = sizeof(double) * (size_t)lqr_ws_doubles(n) * (size_t)grid;
  return 0;
}

extern "C" int crb_lqr_gains(int32_t n, int32_t n_members, const double* M_beam, int32_t m_shared, const double* K_beam,
                             int32_t k_shared, const double* Q, const double* R, int32_t refine_passes, double* gain_out,
                             double* S_out, double* residual_out, int32_t* status_out, void* workspace,
                             size_t workspace_bytes, void* stream) {
  if (!M_beam || !K_beam || !Q || !R || !gain_out || !status_out || !workspace)
    return crb_fail(CRB_E_ARG, "crb_lqr_gains: null argument");
  if (int rc = lqr_check_n(n, "crb_lqr_gains")) return rc;
  if (n_members < 1) return crb_fail(CRB_E_ARG, "crb_lqr_gains: n_members must be positive");
  if (refine_passes < 0 || refine_passes > 8) return crb_fail(CRB_E_ARG, "crb_lqr_gains: refine_passes must be in [0, 8]");
  int grid = 0;
  if (int rc = lqr_grid(n, n_members, &grid)) return rc;
  const size_t need = sizeof(double) * (size_t)lqr_ws_doubles(n) * (size_t)grid;
  if (workspace_bytes < need)
    return crb_fail(CRB_E_ARG, "crb_lqr_gains: workspace of %zu bytes, %zu needed (crb_lqr_workspace_bytes)", workspace_bytes, need);
  LqrArgs a;
  a.n = n;
  a.n_members = n_members;
  a.Mb = M_beam;
  a.Kb = K_beam;
  a.m_shared = m_shared;
  a.k_shared = k_shared;
  a.Q = Q;
  a.R = R;
  a.passes = refine_passes;
  a.gain = gain_out;
  a.S_out = S_out;
  a.residual = residual_out;
  a.status = status_out;
  a.ws = static_cast<double*>(workspace);
  a.ws_stride = lqr_ws_doubles(n);
  a.big = lqr_big(n) ? 1 : 0;
  if (a.big) crb_lqr_kernel<true><<<grid, CRB_LQR_THREADS, lqr_smem_bytes(n), static_cast<cudaStream_t>(stream)>>>(a);
  else crb_lqr_kernel<false><<<grid, CRB_LQR_THREADS, lqr_smem_bytes(n), static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_lqr_gains: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int crb_member_operators(int32_t n, int32_t n_members, const double* M_beam, int32_t m_shared, const double* K_beam,
                                    int32_t k_shared, const double* gain, const double* ref, double* op_out,
                                    int32_t* status_out, void* stream) {
  if (!M_beam || !K_beam || !gain || !op_out || !status_out) return crb_fail(CRB_E_ARG, "crb_member_operators: null argument");
  if (n < 1 || n > 32) return crb_fail(CRB_E_LIMIT, "crb_member_operators: n = %d free DOFs outside [1, 32] (one lane per DOF)", n);
  if (n_members < 1) return crb_fail(CRB_E_ARG, "crb_member_operators: n_members must be positive");
  const size_t bytes = sizeof(double) * ((size_t)n * n + 4 * (n + 16) + 32 + n) + sizeof(int) * (2 * n + 8);
  const int grid = n_members < 148 * 6 ? n_members : 148 * 6;
  crb_member_operator_kernel<<<grid, CRB_LQR_THREADS, bytes, static_cast<cudaStream_t>(stream)>>>(
      n, n_members, M_beam, m_shared, K_beam, k_shared, gain, ref, op_out, status_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_member_operators: %s", cudaGetErrorString(e));
  return 0;
}
