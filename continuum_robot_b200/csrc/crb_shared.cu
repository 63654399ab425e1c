// crb_shared.cu -- shared-operator RK4: the closed-loop step as a dense FP64 tensor-core contraction.
//
// When ONE linear design (mass, stiffness, boundary conditions), ONE gravity setting and ONE
// feedback gain are shared by every member of the ensemble (BASELINE config 5: the LQR rollout of
// examples/lqr_control.py:87-130, members differing in disturbance and initial state only), the
// right-hand side of models/dynamic_beam_model.py:256-272, 343-362 with control/full_state_linear.py:58
//
//     a = M^-1 ( -K q + f_grav(q) + e_k imp(t) + G (r - x) )
//       = W [q ; v] + (M^-1 Gc) cos(phibar) + (M^-1 Gs) sin(phibar) + c0 + imp(t) M^-1 e_k
//
// is a product of the member states with operators that do not depend on the member:
//     W  = [ -M^-1 (K + G_q) | -M^-1 G_v ]   (n x 2n),   c0 = M^-1 G r,
//     phibar = P q  (segment-average rotations, models/gravity_forces.py:104-115, reduced indices,
//                    SURVEY Q2),  Gc / Gs = the gravity placement of gravity_forces.py:117-146.
// Eight members form the rows of mma.sync.m8n8k4.f64; lane j of a member's 4 lanes owns the reduced
// DOFs 4 i + j, which is at the same time the A-fragment layout (k-tile i = DOFs 4i .. 4i+3) and,
// because the OUTPUT columns of every operator are permuted on the host (column 2j+e of n-tile nt
// <-> DOF 4 (2 nt + e) + j), the C-fragment layout: stage inputs and accelerations never leave the
// lane that owns them -- no shuffles, no shared-memory transposes, no banded solve.
//
// The operators are built once on the host (crb_shared_operator, long-double Cholesky of M) and
// staged in shared memory as B fragments (one double per lane and tile).
#include <algorithm>
#include <cstring>
#include <vector>

#include "crb_internal.h"

#define CRB_SH_WARPS 4
#define CRB_SH_THREADS (32 * CRB_SH_WARPS)
#ifndef CRB_SH_MINBLOCKS
#define CRB_SH_MINBLOCKS 3  // resident blocks per SM the register allocation is sized for (170 registers: no spills)
#endif
#define CRB_SH_MAX_KQ 6      // n_free <= 24
#define CRB_SH_HEADER 8

// blob layout (doubles): header[8] = {magic, n, KQ, NT, GKP, Ns, imp_dof, total}, then
//   Wq[KQ][NT][32], Wv[KQ][NT][32], Pf[KQ][32], Gc[GKP][NT][32], Gs[GKP][NT][32], c0[4 KQ], mi[4 KQ],
//   order[4 KQ] as int32 (2 KQ doubles), tile masks[8] as int32 (4 doubles).
// `order[r]` = the reduced DOF that INTERNAL index r (k-tile r / 4, lane r % 4 of the member's quad) stands for, or -1
// for a padding entry; every per-DOF array of the blob (k-tiles, output columns, c0, mi) is in internal order.  The
// host picks the order that leaves the most operator tiles entirely zero -- e.g. a straight beam's axial DOFs and its
// bending DOFs do not couple through M, K or an LQR gain designed on them (examples/lqr_control.py:46-84), so with the
// axial DOFs grouped into their own k-tiles the cross tiles of W vanish, and with gravity along -y the cos / sin
// operators each fill only the bending / only the axial output tiles -- and `masks` = {Wq, Wv, Pf, Gc, Gs} say which
// tiles (bit i * NT + nt, resp. i, resp. p * NT + nt) hold a nonzero: the kernel issues no DMMA for the others
// (config 5: 24 instead of 47 per right-hand side).  A dense random gain keeps the natural order and every tile.
static const double kSharedMagic = 5171.0;

struct SharedDims {
  int n, KQ, NT, GKP, Ns;
  long long o_wq, o_wv, o_pf, o_gc, o_gs, o_c0, o_mi, o_ord, o_msk, total;
};

static SharedDims shared_dims(int n, int Ns, bool grav) {
  SharedDims d;
  d.n = n;
  d.Ns = Ns;
  d.KQ = (n + 3) / 4;
  d.NT = (d.KQ + 1) / 2;
  d.GKP = grav ? (Ns + 3) / 4 : 0;
  long long o = CRB_SH_HEADER;
  d.o_wq = o; o += 32ll * d.KQ * d.NT;
  d.o_wv = o; o += 32ll * d.KQ * d.NT;
  d.o_pf = o; o += 32ll * d.KQ;
  d.o_gc = o; o += 32ll * d.GKP * d.NT;
  d.o_gs = o; o += 32ll * d.GKP * d.NT;
  d.o_c0 = o; o += 4ll * d.KQ;
  d.o_mi = o; o += 4ll * d.KQ;
  d.o_ord = o; o += 2ll * d.KQ;
  d.o_msk = o; o += 4;
  d.total = o;
  return d;
}

// ------------------------------------------------------------------------------------------
// internal DOF order and tile masks (host)
// ------------------------------------------------------------------------------------------
namespace {
typedef long double ldbl;
struct SharedOps {  // dense operators in reduced-DOF order: W [n,2n], Pm [N,n], Gc / Gs [n,N]
  int n, N, KQ, NT, GKP;
  const std::vector<ldbl>*W, *Pm, *Gc, *Gs;
};
struct SharedMasks {
  unsigned wq, wv, pf, gc, gs;
  int count() const { return __builtin_popcount(wq) + __builtin_popcount(wv) + __builtin_popcount(pf) + __builtin_popcount(gc) + __builtin_popcount(gs); }
};

// which tiles hold a nonzero when internal index r stands for reduced DOF ord[r] (fragment positions as in the kernel:
// lane -> k = lane % 4, output column 2 jo + e = lane / 4, output index 4 (2 nt + e) + jo)
SharedMasks shared_tile_masks(const SharedOps& O, const std::vector<int>& ord) {
  SharedMasks m{0, 0, 0, 0, 0};
  const int n = O.n;
  for (int lane = 0; lane < 32; ++lane) {
    const int k = lane % 4, ncol = lane / 4, jo = ncol / 2, e = ncol % 2;
    for (int nt = 0; nt < O.NT; ++nt) {
      const int oi = 4 * (2 * nt + e) + jo, o = oi < 4 * O.KQ ? ord[oi] : -1;
      if (o < 0) continue;
      for (int i = 0; i < O.KQ; ++i) {
        const int c = ord[4 * i + k];
        if (c < 0) continue;
        if ((*O.W)[(size_t)o * 2 * n + c] != 0.0L) m.wq |= 1u << (i * O.NT + nt);
        if ((*O.W)[(size_t)o * 2 * n + n + c] != 0.0L) m.wv |= 1u << (i * O.NT + nt);
      }
      for (int p = 0; p < O.GKP; ++p) {
        const int seg = 4 * p + k;
        if (seg >= O.N) continue;
        if ((*O.Gc)[(size_t)o * O.N + seg] != 0.0L) m.gc |= 1u << (p * O.NT + nt);
        if ((*O.Gs)[(size_t)o * O.N + seg] != 0.0L) m.gs |= 1u << (p * O.NT + nt);
      }
    }
    const int seg = 4 * e + jo;
    if (O.GKP > 0 && seg < O.N)
      for (int i = 0; i < O.KQ; ++i) {
        const int c = ord[4 * i + k];
        if (c >= 0 && (*O.Pm)[(size_t)seg * n + c] != 0.0L) m.pf |= 1u << i;
      }
  }
  return m;
}

// Candidate orders: the natural one, and the connected components of W's coupling pattern laid out one after the
// other (every sequence of up to 4 components; each padded to whole k-tiles when that fits 4 KQ entries, and unpadded;
// inside a component in natural order, or with the DOFs that enter the segment rotations P q last).
std::vector<std::vector<int>> shared_candidate_orders(const SharedOps& O) {
  const int n = O.n, cap = 4 * O.KQ;
  std::vector<std::vector<int>> out;
  std::vector<int> nat(cap, -1);
  for (int r = 0; r < n; ++r) nat[r] = r;
  out.push_back(nat);
  std::vector<int> root(n);
  for (int r = 0; r < n; ++r) root[r] = r;
  auto find = [&](int a) { while (root[a] != a) a = root[a] = root[root[a]]; return a; };
  for (int o = 0; o < n; ++o)
    for (int c = 0; c < n; ++c)
      if ((*O.W)[(size_t)o * 2 * n + c] != 0.0L || (*O.W)[(size_t)o * 2 * n + n + c] != 0.0L) root[find(o)] = find(c);
  std::vector<std::vector<int>> comps;
  for (int r = 0; r < n; ++r) {
    if (find(r) != r) continue;
    std::vector<int> cmp;
    for (int a = 0; a < n; ++a) if (find(a) == r) cmp.push_back(a);
    comps.push_back(cmp);
  }
  if (comps.size() < 2) return out;
  auto in_p = [&](int dof) {
    for (int s = 0; s < O.N && O.GKP > 0; ++s) if ((*O.Pm)[(size_t)s * n + dof] != 0.0L) return true;
    return false;
  };
  std::vector<int> seq(comps.size());
  for (size_t a = 0; a < seq.size(); ++a) seq[a] = (int)a;
  do {
    for (int plast = 0; plast < 2; ++plast)
      for (int padded = 0; padded < 2; ++padded) {
        std::vector<int> ord;
        for (int ci : seq) {
          std::vector<int> cmp = comps[ci];
          if (plast) std::stable_sort(cmp.begin(), cmp.end(), [&](int a, int b) { return in_p(a) < in_p(b); });
          ord.insert(ord.end(), cmp.begin(), cmp.end());
          while (padded && ord.size() % 4) ord.push_back(-1);
        }
        if ((int)ord.size() > cap) continue;
        ord.resize(cap, -1);
        out.push_back(ord);
      }
  } while (comps.size() <= 4 && std::next_permutation(seq.begin(), seq.end()));
  return out;
}
}  // namespace

extern "C" int64_t crb_shared_operator(const crb_plan_t* plan, const double* params_host, const uint8_t* elem_type_host,
                                       const uint8_t* bc_host, const double* gain_host, const double* ref_host,
                                       double gx, double gy, int32_t gravity_on, int32_t imp_dof, double* out_host) {
  if (!plan) return crb_fail(CRB_E_ARG, "crb_shared_operator: null plan");
  const int n = plan->n_free, N = plan->n_elements;
  if (n < 1 || n > 4 * CRB_SH_MAX_KQ) return crb_fail(CRB_E_LIMIT, "crb_shared_operator: n_free %d outside [1, %d]", n, 4 * CRB_SH_MAX_KQ);
  if (gravity_on && N > 8) return crb_fail(CRB_E_LIMIT, "crb_shared_operator: gravity needs <= 8 segments, got %d", N);
  if (imp_dof >= n) return crb_fail(CRB_E_ARG, "crb_shared_operator: imp_dof %d outside [0,%d)", imp_dof, n);
  const SharedDims D = shared_dims(n, N, gravity_on != 0);
  if (!out_host) return D.total;
  if (!params_host || !elem_type_host || !bc_host) return crb_fail(CRB_E_ARG, "crb_shared_operator: null argument");
  std::vector<double> M((size_t)n * n), K((size_t)n * n);
  if (int rc = crb_dense_matrices(plan, params_host, elem_type_host, bc_host, M.data(), K.data())) return rc;
  // M^-1 by Cholesky in extended precision (M is SPD, cond ~ 1e4)
  typedef long double ld;
  std::vector<ld> Lc((size_t)n * n, 0.0L), Minv((size_t)n * n, 0.0L);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      ld s = M[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) s -= Lc[(size_t)i * n + k] * Lc[(size_t)j * n + k];
      if (i == j) {
        if (!(s > 0.0L)) return crb_fail(CRB_E_ARG, "crb_shared_operator: mass matrix is not positive definite");
        Lc[(size_t)i * n + i] = sqrtl(s);
      } else {
        Lc[(size_t)i * n + j] = s / Lc[(size_t)j * n + j];
      }
    }
  for (int c = 0; c < n; ++c) {  // solve L L^T x = e_c
    std::vector<ld> y(n);
    for (int i = 0; i < n; ++i) {
      ld s = (i == c) ? 1.0L : 0.0L;
      for (int k = 0; k < i; ++k) s -= Lc[(size_t)i * n + k] * y[k];
      y[i] = s / Lc[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      ld s = y[i];
      for (int k = i + 1; k < n; ++k) s -= Lc[(size_t)k * n + i] * Minv[(size_t)k * n + c];
      Minv[(size_t)i * n + c] = s / Lc[(size_t)i * n + i];
    }
  }
  // W = [-Minv (K + Gq) | -Minv Gv],  c0 = Minv (G r)
  std::vector<ld> W((size_t)n * 2 * n, 0.0L), c0(n, 0.0L);
  for (int o = 0; o < n; ++o)
    for (int c = 0; c < 2 * n; ++c) {
      ld s = 0.0L;
      for (int k = 0; k < n; ++k) {
        ld m = (c < n) ? (ld)K[(size_t)k * n + c] : 0.0L;
        if (gain_host) m += (ld)gain_host[(size_t)k * 2 * n + c];
        s += Minv[(size_t)o * n + k] * m;
      }
      W[(size_t)o * 2 * n + c] = -s;
    }
  if (gain_host && ref_host)
    for (int o = 0; o < n; ++o) {
      ld s = 0.0L;
      for (int k = 0; k < n; ++k) {
        ld gr = 0.0L;
        for (int c = 0; c < 2 * n; ++c) gr += (ld)gain_host[(size_t)k * 2 * n + c] * (ld)ref_host[c];
        s += Minv[(size_t)o * n + k] * gr;
      }
      c0[o] = s;
    }
  // gravity: phibar = P q and the placement vectors (gravity_forces.py:97-146, reduced indices)
  std::vector<ld> Pm((size_t)(N > 0 ? N : 1) * n, 0.0L), Gc((size_t)n * (N > 0 ? N : 1), 0.0L), Gs((size_t)n * (N > 0 ? N : 1), 0.0L);
  if (gravity_on)
    for (int i = 0; i < N; ++i) {
      const int a = 3 * i + 2, b = 3 * (i + 1) + 2;
      if (a < n && b < n) { Pm[(size_t)i * n + a] = 0.5L; Pm[(size_t)i * n + b] = 0.5L; }
      else if (a < n) Pm[(size_t)i * n + a] = 1.0L;
      else if (b < n) Pm[(size_t)i * n + b] = 1.0L;
      const double* q = params_host + (size_t)i * CRB_NPARAM;
      const ld hm = 0.5L * (ld)(q[CRB_P_RHO] * q[CRB_P_AREA] * q[CRB_P_LENGTH]);  // segment mass as the reference rounds it
      std::vector<ld> fc(n, 0.0L), fs(n, 0.0L);  // force = fc cos + fs sin
      const int ia[2] = {3 * i, 3 * (i + 1)}, it[2] = {3 * i + 1, 3 * (i + 1) + 1};
      for (int e = 0; e < 2; ++e) {
        if (ia[e] < n) { fc[ia[e]] += hm * gx; fs[ia[e]] += hm * gy; }
        if (it[e] < n) { fc[it[e]] += hm * gy; fs[it[e]] -= hm * gx; }
      }
      for (int o = 0; o < n; ++o) {
        ld sc = 0.0L, ss = 0.0L;
        for (int k = 0; k < n; ++k) { sc += Minv[(size_t)o * n + k] * fc[k]; ss += Minv[(size_t)o * n + k] * fs[k]; }
        Gc[(size_t)o * N + i] = sc;
        Gs[(size_t)o * N + i] = ss;
      }
    }
  // internal DOF order: the candidate that leaves the fewest nonzero tiles (ties: the natural order)
  const SharedOps O{n, N, D.KQ, D.NT, D.GKP, &W, &Pm, &Gc, &Gs};
  std::vector<int> ord;
  SharedMasks best{};
  for (const std::vector<int>& cand : shared_candidate_orders(O)) {
    const SharedMasks m = shared_tile_masks(O, cand);
    if (ord.empty() || m.count() < best.count()) { ord = cand; best = m; }
  }
  for (long long k = 0; k < D.total; ++k) out_host[k] = 0.0;
  out_host[0] = kSharedMagic; out_host[1] = n; out_host[2] = D.KQ; out_host[3] = D.NT; out_host[4] = D.GKP;
  out_host[5] = N; out_host[6] = imp_dof; out_host[7] = (double)D.total;
  for (int lane = 0; lane < 32; ++lane) {
    const int k = lane % 4, ncol = lane / 4, jo = ncol / 2, e = ncol % 2;
    for (int i = 0; i < D.KQ; ++i) {
      const int c = ord[4 * i + k];
      for (int nt = 0; nt < D.NT; ++nt) {
        const int oi = 4 * (2 * nt + e) + jo, o = oi < 4 * D.KQ ? ord[oi] : -1;
        if (o >= 0 && c >= 0) {
          out_host[D.o_wq + ((long long)i * D.NT + nt) * 32 + lane] = (double)W[(size_t)o * 2 * n + c];
          out_host[D.o_wv + ((long long)i * D.NT + nt) * 32 + lane] = (double)W[(size_t)o * 2 * n + n + c];
        }
      }
      const int seg = 4 * e + jo;  // phibar tile: column 2 jo + e <-> segment 4 e + jo
      if (gravity_on && seg < N && c >= 0) out_host[D.o_pf + (long long)i * 32 + lane] = (double)Pm[(size_t)seg * n + c];
    }
    for (int p = 0; p < D.GKP; ++p) {
      const int seg = 4 * p + k;
      for (int nt = 0; nt < D.NT; ++nt) {
        const int oi = 4 * (2 * nt + e) + jo, o = oi < 4 * D.KQ ? ord[oi] : -1;
        if (seg < N && o >= 0) {
          out_host[D.o_gc + ((long long)p * D.NT + nt) * 32 + lane] = (double)Gc[(size_t)o * N + seg];
          out_host[D.o_gs + ((long long)p * D.NT + nt) * 32 + lane] = (double)Gs[(size_t)o * N + seg];
        }
      }
    }
  }
  int32_t tail[4 * CRB_SH_MAX_KQ + 8] = {0};
  for (int r = 0; r < 4 * D.KQ; ++r) {
    const int o = ord[r];
    tail[r] = o;
    if (o < 0) continue;
    out_host[D.o_c0 + r] = (double)c0[o];
    if (imp_dof >= 0) out_host[D.o_mi + r] = (double)Minv[(size_t)o * n + imp_dof];
  }
  int32_t* msk = tail + 4 * D.KQ;
  msk[0] = (int32_t)best.wq; msk[1] = (int32_t)best.wv; msk[2] = (int32_t)best.pf; msk[3] = (int32_t)best.gc; msk[4] = (int32_t)best.gs;
  memcpy(out_host + D.o_ord, tail, sizeof(int32_t) * (4 * D.KQ + 8));  // o_msk follows o_ord directly
  return D.total;
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}


struct SharedArgs {
  const double* blob;
  const double* imp_amp;  // [B] or NULL
  double imp_duration;
  int n_members, n, blob_doubles;
  int o_wq, o_wv, o_pf, o_gc, o_gs, o_c0, o_mi, o_ord;
  const int* sel_inv;  // lean recording (crb_system_t.out_sel_inv) or NULL
  int n_sel;
};

// Tile masks the specialised code path of crb_rk4_shared_kernel<KQ, GKP, .> is compiled for: the straight FIXED-root
// beam (n = 3 N, examples/lqr_control.py:26-44) with a gain that does not couple axial and bending DOFs and gravity
// along y, in the internal order crb_shared_operator picks for it (union over the element counts that share KQ).  A
// predicated-off DMMA still occupies the FP64 pipe for its 16 cycles (measured: masks as run-time predicates left the
// kernel at 13.3 ms), so the masks must be compile-time constants; a blob whose masks are not a subset of these runs
// the all-tiles path of the same kernel.
struct SharedPattern { unsigned wq, wv, pf, gc, gs; };
__host__ __device__ constexpr SharedPattern shared_sparse_pattern(int KQ, int GKP) {
  return KQ == 3 && GKP == 0 ? SharedPattern{0x25, 0x25, 0, 0, 0}
       : KQ == 3 && GKP == 1 ? SharedPattern{0x25, 0x25, 0x3, 0x1, 0x2}
       : KQ == 4 && GKP == 0 ? SharedPattern{0xfd, 0xfd, 0, 0, 0}
       : KQ == 4 && GKP == 2 ? SharedPattern{0xfd, 0xfd, 0xc, 0xf, 0x5}
       : KQ == 5 && GKP == 0 ? SharedPattern{0x6d89, 0x6d89, 0, 0, 0}
       : KQ == 5 && GKP == 2 ? SharedPattern{0x6d89, 0x6d89, 0x18, 0x36, 0x9}
       : KQ == 6 && GKP == 0 ? SharedPattern{0x36d89, 0x36d89, 0, 0, 0}
       : KQ == 6 && GKP == 2 ? SharedPattern{0x36d89, 0x36d89, 0x38, 0x36, 0x9}
                             : SharedPattern{~0u, ~0u, ~0u, ~0u, ~0u};
}

extern "C" int crb_shared_sparse_masks(int32_t KQ, int32_t GKP, uint32_t* out5) {
  if (!out5 || KQ < 1 || KQ > CRB_SH_MAX_KQ || GKP < 0 || GKP > 2) return crb_fail(CRB_E_ARG, "crb_shared_sparse_masks: bad argument");
  const SharedPattern p = shared_sparse_pattern(KQ, GKP);
  out5[0] = p.wq; out5[1] = p.wv; out5[2] = p.pf; out5[3] = p.gc; out5[4] = p.gs;
  return 0;
}

// nsteps RK4 steps of the 8 members of a warp; only the tiles named by the compile-time masks are multiplied.
template <int KQ, int GKP, bool IMP, unsigned MWQ, unsigned MWV, unsigned MPF, unsigned MGC, unsigned MGS>
__device__ __forceinline__ void shared_steps(const SharedArgs& A, const double* smem, const int* order, int lane, int mem, bool active,
                                             double (&q)[KQ], double (&v)[KQ], double t0, double h, int nsteps,
                                             double* __restrict__ Y, int save_every) {
  constexpr int NT = (KQ + 1) / 2;
  const int n = A.n;
  const double* Wq = smem + A.o_wq + lane;
  const double* Wv = smem + A.o_wv + lane;
  const double* Pf = smem + A.o_pf + lane;
  const double* Gc = smem + A.o_gc + lane;
  const double* Gs = smem + A.o_gs + lane;
  double c0[KQ], mi[KQ];
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    c0[i] = smem[A.o_c0 + 4 * i + (lane & 3)];
    mi[i] = IMP ? smem[A.o_mi + 4 * i + (lane & 3)] : 0.0;
  }
  const double amp = IMP ? A.imp_amp[mem] : 0.0;
  const double hh = 0.5 * h, h6 = h / 6.0, h3 = h / 3.0;
  for (int k = 0; k < nsteps; ++k) {
    const double t = t0 + k * h;
    double qs[KQ], vs[KQ], aq[KQ], av[KQ];
#pragma unroll
    for (int i = 0; i < KQ; ++i) {
      qs[i] = q[i];
      vs[i] = v[i];
      aq[i] = q[i];
      av[i] = v[i];
    }
#pragma unroll 1
    for (int st = 0; st < 4; ++st) {
      const double ts = t + (st == 0 ? 0.0 : (st == 3 ? h : hh));
      const double gate = (IMP && ts < A.imp_duration) ? amp : 0.0;
      double acc[NT][2];
#pragma unroll
      for (int i = 0; i < 2 * NT; ++i) acc[i / 2][i % 2] = i < KQ ? fma(gate, mi[i < KQ ? i : 0], c0[i < KQ ? i : 0]) : 0.0;
      if (GKP > 0) {
        double ph[2] = {0.0, 0.0};  // segment-average rotations of segments j and 4 + j
#pragma unroll
        for (int i = 0; i < KQ; ++i)
          if (MPF >> i & 1) dmma884(ph[0], ph[1], qs[i], Pf[i * 32]);
#pragma unroll
        for (int p = 0; p < GKP; ++p) {
          double sn, cs;
          crb_sincos<true, true>(ph[p], sn, cs);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            if (MGC >> (p * NT + nt) & 1) dmma884(acc[nt][0], acc[nt][1], cs, Gc[(p * NT + nt) * 32]);
            if (MGS >> (p * NT + nt) & 1) dmma884(acc[nt][0], acc[nt][1], sn, Gs[(p * NT + nt) * 32]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < KQ; ++i)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          if (MWQ >> (i * NT + nt) & 1) dmma884(acc[nt][0], acc[nt][1], qs[i], Wq[(i * NT + nt) * 32]);
          if (MWV >> (i * NT + nt) & 1) dmma884(acc[nt][0], acc[nt][1], vs[i], Wv[(i * NT + nt) * 32]);
        }
      const double wgt = (st == 0 || st == 3) ? h6 : h3;  // b = (1/6, 1/3, 1/3, 1/6)
      const double cn = st == 2 ? h : hh;                 // next stage: x + c k
#pragma unroll
      for (int i = 0; i < KQ; ++i) {
        const double a = acc[i / 2][i % 2];
        aq[i] = fma(wgt, vs[i], aq[i]);
        av[i] = fma(wgt, a, av[i]);
        qs[i] = fma(cn, vs[i], q[i]);
        vs[i] = fma(cn, a, v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < KQ; ++i) {
      q[i] = aq[i];
      v[i] = av[i];
    }
    if (Y && save_every > 0 && (k + 1) % save_every == 0 && active) {
      double* ym = Y + ((long long)((k + 1) / save_every - 1) * A.n_members + mem) * (A.sel_inv ? A.n_sel : 2 * n);
#pragma unroll
      for (int i = 0; i < KQ; ++i) {
        const int r = order[4 * i];
        if (r >= 0) frame_put(A.sel_inv, ym, n, r, q[i], v[i]);
      }
    }
  }
}

template <int KQ, int GKP, bool IMP>
__global__ void __launch_bounds__(CRB_SH_THREADS, CRB_SH_MINBLOCKS)
crb_rk4_shared_kernel(SharedArgs A, double* __restrict__ X, double t0, double h, int nsteps, double* __restrict__ Y,
                      int save_every) {
  extern __shared__ __align__(16) double smem[];
  for (int k = threadIdx.x; k < A.blob_doubles; k += blockDim.x) smem[k] = A.blob[k];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 3;
  const int member = (blockIdx.x * CRB_SH_WARPS + warp) * 8 + (lane >> 2);
  const bool active = member < A.n_members;
  const int mem = active ? member : A.n_members - 1;
  const int n = A.n;
  // internal index 4 i + j <-> reduced DOF order[4 i + j] (-1: padding); tile masks behind the order table
  const int* const order = reinterpret_cast<const int*>(smem + A.o_ord) + j;
  const unsigned* const masks = reinterpret_cast<const unsigned*>(smem + A.o_ord) + 4 * KQ;
  double q[KQ], v[KQ];
  double* xm = X + (long long)mem * 2 * n;
#pragma unroll
  for (int i = 0; i < KQ; ++i) {
    const int r = order[4 * i];
    const bool ok = r >= 0;
    q[i] = ok ? xm[r] : 0.0;
    v[i] = ok ? xm[n + r] : 0.0;
  }
  constexpr SharedPattern SP = shared_sparse_pattern(KQ, GKP);
  if (SP.wq != ~0u && !(masks[0] & ~SP.wq) && !(masks[1] & ~SP.wv) && !(masks[2] & ~SP.pf) && !(masks[3] & ~SP.gc) && !(masks[4] & ~SP.gs))
    shared_steps<KQ, GKP, IMP, SP.wq, SP.wv, SP.pf, SP.gc, SP.gs>(A, smem, order, lane, mem, active, q, v, t0, h, nsteps, Y, save_every);
  else
    shared_steps<KQ, GKP, IMP, ~0u, ~0u, ~0u, ~0u, ~0u>(A, smem, order, lane, mem, active, q, v, t0, h, nsteps, Y, save_every);
  if (active) {
#pragma unroll
    for (int i = 0; i < KQ; ++i) {
      const int r = order[4 * i];
      if (r >= 0) {
        xm[r] = q[i];
        xm[n + r] = v[i];
      }
    }
  }
}

bool crb_shared_eligible(const crb_plan_t* plan, const crb_system_t* sys) {
  return sys->shared_op && !sys->force_general && sys->all_linear && sys->mass_shared && sys->stiff_shared && sys->gain_stride == 0 &&
         !sys->drag && !sys->u_const && !sys->f_ext && !crb_time_varying_input(sys) && plan->n_free <= 4 * CRB_SH_MAX_KQ;
}

int crb_launch_rk4_shared(const crb_plan_t* plan, const crb_system_t* sys, double* X, double t0, double h, int nsteps,
                          double* Y_out, int save_every, cudaStream_t stream) {
  const int n = plan->n_free;
  const bool grav = sys->grav_mode != 0;
  const SharedDims D = shared_dims(n, plan->n_elements, grav);
  if (sys->shared_op_doubles != D.total)
    return crb_fail(CRB_E_ARG, "crb_rk4: shared operator has %lld doubles, this system needs %lld (gravity %s)",
                    (long long)sys->shared_op_doubles, D.total, grav ? "on" : "off");
  SharedArgs A;
  A.blob = sys->shared_op;
  A.imp_amp = sys->imp_amp;
  A.imp_duration = sys->imp_duration;
  A.n_members = sys->n_members;
  A.sel_inv = sys->out_sel_inv;
  A.n_sel = sys->out_n_sel;
  A.n = n;
  A.blob_doubles = (int)D.total;
  A.o_wq = (int)D.o_wq; A.o_wv = (int)D.o_wv; A.o_pf = (int)D.o_pf; A.o_gc = (int)D.o_gc; A.o_gs = (int)D.o_gs;
  A.o_c0 = (int)D.o_c0; A.o_mi = (int)D.o_mi; A.o_ord = (int)D.o_ord;
  const size_t bytes = sizeof(double) * (size_t)D.total;
  const int grid = (sys->n_members + 8 * CRB_SH_WARPS - 1) / (8 * CRB_SH_WARPS);
  const bool imp = sys->imp_amp != nullptr;
#define CRB_SH_LAUNCH(KQV, GV, IV)                                                                       \
  {                                                                                                      \
    if (int rc = set_smem(crb_rk4_shared_kernel<KQV, GV, IV>, bytes, "crb_rk4")) return rc;              \
    crb_rk4_shared_kernel<KQV, GV, IV><<<grid, CRB_SH_THREADS, bytes, stream>>>(A, X, t0, h, nsteps, Y_out, save_every); \
    return 0;                                                                                            \
  }
#define CRB_SH_CASE(KQV)                                              \
  if (D.KQ == KQV) {                                                  \
    if (D.GKP == 0) { if (imp) CRB_SH_LAUNCH(KQV, 0, true) else CRB_SH_LAUNCH(KQV, 0, false) }  \
    if (D.GKP == 1) { if (imp) CRB_SH_LAUNCH(KQV, 1, true) else CRB_SH_LAUNCH(KQV, 1, false) }  \
    if (D.GKP == 2) { if (imp) CRB_SH_LAUNCH(KQV, 2, true) else CRB_SH_LAUNCH(KQV, 2, false) }  \
  }
  CRB_SH_CASE(1) CRB_SH_CASE(2) CRB_SH_CASE(3) CRB_SH_CASE(4) CRB_SH_CASE(5) CRB_SH_CASE(6)
#undef CRB_SH_CASE
#undef CRB_SH_LAUNCH
  return crb_fail(CRB_E_LIMIT, "crb_rk4: shared-operator shape KQ=%d GKP=%d not instantiated", D.KQ, D.GKP);
}
