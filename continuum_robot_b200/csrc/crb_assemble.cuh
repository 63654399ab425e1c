// crb_assemble.cuh -- per-member element assembly and mass factorisation (runs once per ensemble).
//
// One thread per parameter set, a thread block per batch of sets.  Produces, in the lane layout
// consumed by crb_device.cuh:
//   mfac   block-LDL^T factors of the BC-reduced consistent mass matrix plus the partition
//          (SPIKE) and scan constants -- replaces scipy.sparse.linalg.inv(M)
//          (models/dynamic_beam_model.py:60) without ever forming the 5/9-dense inverse;
//   kcoef  4 stiffness coefficients per element (models/segments.py:32-62, 121-157);
//   drag   0.5*rho_f*C_d*A_w per slot with the per-node lookup of fluid_forces.py:56-61,83-90;
//   grav   half segment masses placed by the reduced-index rule of gravity_forces.py:97-146.
// Element matrices: models/segments.py:64-78 (mass); assembly: euler_bernoulli_beam.py:139-161;
// boundary conditions: euler_bernoulli_beam.py:221-298 (constrained DOFs become identity rows).
#pragma once
#include "crb_device.cuh"

struct AsmTopo {
  uint8_t free_bits[CRB_MAX_SLOTS];  // bit d set: DOF d of the slot is free
  uint8_t etype[CRB_MAX_SLOTS];      // type of the element left of slot s (CRB_ELEM_*)
};

struct M2 {
  double a, b, c, d;  // [[a b],[c d]]
};
__device__ __forceinline__ M2 mul(const M2& x, const M2& y) {
  return {x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
__device__ __forceinline__ M2 tr(const M2& x) { return {x.a, x.c, x.b, x.d}; }
__device__ __forceinline__ M2 neg(const M2& x) { return {-x.a, -x.b, -x.c, -x.d}; }
__device__ __forceinline__ M2 sub(const M2& x, const M2& y) { return {x.a - y.a, x.b - y.b, x.c - y.c, x.d - y.d}; }
__device__ __forceinline__ M2 inv(const M2& x) {
  const double det = x.a * x.d - x.b * x.c;
  const double r = 1.0 / det;
  return {x.d * r, -x.b * r, -x.c * r, x.a * r};
}

// address of constant `idx` (0..25) of slot s inside one mfac set
__device__ __forceinline__ long long mf_addr(int idx, int s, int m, int G) {
  const int g = s / m, j = s - g * m;
  return ((((long long)(idx >> 1) * m + j) * G + g) << 1) + (idx & 1);
}
__device__ __forceinline__ long long sc_addr(int level, int idx, int g, int G) {
  return ((((long long)level * CRB_SCAN_PAIRS + (idx >> 1)) * G + g) << 1) + (idx & 1);
}
__device__ __forceinline__ void st_m2(double* mf, int base, int s, int m, int G, const M2& x) {
  mf[mf_addr(base, s, m, G)] = x.a;
  mf[mf_addr(base + 1, s, m, G)] = x.b;
  mf[mf_addr(base + 2, s, m, G)] = x.c;
  mf[mf_addr(base + 3, s, m, G)] = x.d;
}
__device__ __forceinline__ M2 ld_m2(const double* mf, int base, int s, int m, int G) {
  return {mf[mf_addr(base, s, m, G)], mf[mf_addr(base + 1, s, m, G)], mf[mf_addr(base + 2, s, m, G)],
          mf[mf_addr(base + 3, s, m, G)]};
}

__global__ void __launch_bounds__(128)
crb_assemble_kernel(KPlan P, AsmTopo T, const double* __restrict__ params, int n_param_sets,
                    int n_mass, int n_stiff, int n_force, double fluid_density, double shift,
                    double* __restrict__ mfac, double* __restrict__ kcoef,
                    uint8_t* __restrict__ etype_out, double* __restrict__ drag,
                    double* __restrict__ grav, double* __restrict__ seg_half_mass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = P.m, G = P.g, Pn = P.p, N = P.N;
  if (i == 0 && etype_out)
    for (int s = 0; s < Pn; ++s) etype_out[s] = T.etype[s];
  const double* par = params + (n_param_sets == 1 ? 0ll : (long long)i * N * CRB_NPARAM);
  auto elem_of_slot = [&](int s) -> int {  // physical element left of slot s, or -1
    if (s >= P.p_act) return -1;
    const int e = P.n0 + s - 1;
    return (e >= 0 && e < N) ? e : -1;
  };

  // ---------------- stiffness coefficients ----------------
  if (i < n_stiff) {
    double* kc = kcoef + (long long)i * Pn * 4;
    for (int s = 0; s < Pn; ++s) {
      const int e = elem_of_slot(s);
      double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
      if (e >= 0) {
        const double* q = par + e * CRB_NPARAM;
        const double L = q[CRB_P_LENGTH];
        const double EI = q[CRB_P_E] * q[CRB_P_I];
        const double EA = q[CRB_P_E] * q[CRB_P_AREA];
        if (T.etype[s] == CRB_ELEM_LINEAR) {
          c0 = EA / L;
          c1 = 12 * EI / (L * L * L);
          c2 = 6 * EI / (L * L);
          c3 = 2 * EI / L;
        } else {
          c0 = EA / (L * L);
          c1 = EI / (L * L);
          c2 = L;
          c3 = 1.0 / L;
        }
      }
      kc[4 * s + 0] = c0;
      kc[4 * s + 1] = c1;
      kc[4 * s + 2] = c2;
      kc[4 * s + 3] = c3;
    }
  }

  // ---------------- built-in force coefficients ----------------
  if (i < n_force) {
    if (drag) {
      double* dr = drag + (long long)i * Pn;
      for (int s = 0; s < Pn; ++s) {
        double f = 0.0;
        if (s < P.p_act && (T.free_bits[s] & 2)) {
          const int node = s + P.n0;
          const int row = node < N ? node : N - 1;  // fluid_forces.py:60-61
          const double* q = par + row * CRB_NPARAM;
          f = 0.5 * fluid_density * q[CRB_P_CD] * q[CRB_P_WETTED];
        }
        dr[s] = f;
      }
    }
    if (grav) {
      double* gr = grav + (long long)i * Pn * 2;
      for (int s = 0; s < Pn; ++s) {
        double gl = 0.0, gt = 0.0;
        if (P.contiguous) {
          // pseudo-segment k joins reduced node slots k and k+1 and carries the mass of CSV row k
          if (s >= 1 && s < P.p_act && s - 1 < N) {
            const double* q = par + (s - 1) * CRB_NPARAM;
            gl = 0.5 * (q[CRB_P_RHO] * q[CRB_P_AREA] * q[CRB_P_LENGTH]);
          }
          if (s == P.p_act - 1 && s < N) {
            const double* q = par + s * CRB_NPARAM;
            gt = 0.5 * (q[CRB_P_RHO] * q[CRB_P_AREA] * q[CRB_P_LENGTH]);
          }
        }
        gr[2 * s] = gl;
        gr[2 * s + 1] = gt;
      }
    }
    if (seg_half_mass) {
      double* hm = seg_half_mass + (long long)i * N;
      for (int e = 0; e < N; ++e) {
        const double* q = par + e * CRB_NPARAM;
        hm[e] = 0.5 * (q[CRB_P_RHO] * q[CRB_P_AREA] * q[CRB_P_LENGTH]);
      }
    }
  }

  // ---------------- mass factorisation ----------------
  if (i >= n_mass) return;
  double* mf = mfac + (long long)i * P.mfac_doubles;
  double* sc = mf + 2 * CRB_SLOT_PAIRS * Pn;
  for (long long k = 0; k < P.mfac_doubles; ++k) mf[k] = 0.0;

  // element mass blocks for the element left of slot s (models/segments.py:64-78)
  auto mass_of = [&](int e, double& mu, double& L) {
    const double* q = par + e * CRB_NPARAM;
    L = q[CRB_P_LENGTH];
    mu = q[CRB_P_RHO] * q[CRB_P_AREA] * L / 420;
  };

  // shift != 0: factor M + shift * K instead of M (implicit midpoint, shift = h^2 / 4); linear element
  // stiffness blocks of models/segments.py:32-62 in the same node-block form as the mass blocks
  auto stiff_of = [&](int e, double& ka, double& c1, double& c2, double& c3) {
    const double* q = par + e * CRB_NPARAM;
    const double L = q[CRB_P_LENGTH];
    const double EI = q[CRB_P_E] * q[CRB_P_I];
    ka = shift * (q[CRB_P_E] * q[CRB_P_AREA] / L);
    c1 = shift * (EI / L);
    c2 = shift * (EI / (L * L));
    c3 = shift * (EI / (L * L * L));
  };

  M2 Sinv_prev = {0, 0, 0, 0};
  double sinvu_prev = 0.0;
  for (int s = 0; s < Pn; ++s) {
    const int el = elem_of_slot(s);
    const int er = (s + 1 < P.p_act) ? elem_of_slot(s + 1) : -1;
    const int fb = (s < P.p_act) ? T.free_bits[s] : 0;
    const int fbp = (s >= 1 && s - 1 < P.p_act) ? T.free_bits[s - 1] : 0;
    M2 D = {0, 0, 0, 0}, O = {0, 0, 0, 0};
    double Du = 0.0, Ou = 0.0;
    if (el >= 0) {
      double mu, L;
      mass_of(el, mu, L);
      D.a += 156 * mu;
      D.b += 22 * L * mu;
      D.c += 22 * L * mu;
      D.d += 4 * L * L * mu;
      Du += 140 * mu;
      if (s >= 1) {
        O = {54 * mu, -13 * L * mu, 13 * L * mu, -3 * L * L * mu};
        Ou = 70 * mu;
      }
      if (shift != 0.0) {
        double ka, c1, c2, c3;
        stiff_of(el, ka, c1, c2, c3);
        D.a += 12 * c3;
        D.b += 6 * c2;
        D.c += 6 * c2;
        D.d += 4 * c1;
        Du += ka;
        if (s >= 1) {
          O.a += -12 * c3;
          O.b += 6 * c2;
          O.c += -6 * c2;
          O.d += 2 * c1;
          Ou += -ka;
        }
      }
    }
    if (er >= 0) {
      double mu, L;
      mass_of(er, mu, L);
      D.a += 156 * mu;
      D.b += -22 * L * mu;
      D.c += -22 * L * mu;
      D.d += 4 * L * L * mu;
      Du += 140 * mu;
      if (shift != 0.0) {
        double ka, c1, c2, c3;
        stiff_of(er, ka, c1, c2, c3);
        D.a += 12 * c3;
        D.b += -6 * c2;
        D.c += -6 * c2;
        D.d += 4 * c1;
        Du += ka;
      }
    }
    // boundary conditions: constrained DOFs become identity rows / columns
    const bool fu = fb & 1, fw = fb & 2, fp = fb & 4;
    const bool pu = fbp & 1, pw = fbp & 2, pp = fbp & 4;
    if (!fu) { Du = 1.0; Ou = 0.0; }
    if (!pu) Ou = 0.0;
    if (!fw) { D.a = 1.0; D.b = 0.0; D.c = 0.0; O.a = 0.0; O.b = 0.0; }
    if (!fp) { D.d = 1.0; D.b = 0.0; D.c = 0.0; O.c = 0.0; O.d = 0.0; }
    if (!pw) { O.a = 0.0; O.c = 0.0; }
    if (!pp) { O.b = 0.0; O.d = 0.0; }

    M2 Lm = {0, 0, 0, 0};
    double lu = 0.0;
    M2 S = D;
    double Su = Du;
    if (s >= 1) {
      Lm = mul(O, Sinv_prev);
      S = sub(D, mul(Lm, tr(O)));
      lu = Ou * sinvu_prev;
      Su = Du - lu * Ou;
      // U_{s-1} = Sinv_{s-1} O_s^T
      st_m2(mf, 8, s - 1, m, G, mul(Sinv_prev, tr(O)));
      mf[mf_addr(22, s - 1, m, G)] = sinvu_prev * Ou;
    }
    const M2 Sinv = inv(S);
    const double sinvu = 1.0 / Su;
    st_m2(mf, 0, s, m, G, Lm);
    st_m2(mf, 4, s, m, G, Sinv);
    mf[mf_addr(20, s, m, G)] = lu;
    mf[mf_addr(21, s, m, G)] = sinvu;
    Sinv_prev = Sinv;
    sinvu_prev = sinvu;
  }
  // chunk-local transfer products: T_s (forward), Psi_s, Phi_s (backward)
  for (int g = 0; g < G; ++g) {
    const int s0 = g * m;
    M2 Tm[8];
    double Tu[8];
    for (int j = 0; j < m; ++j) {
      const M2 Lm = ld_m2(mf, 0, s0 + j, m, G);
      const double lu = mf[mf_addr(20, s0 + j, m, G)];
      if (j == 0) { Tm[0] = neg(Lm); Tu[0] = -lu; }
      else { Tm[j] = neg(mul(Lm, Tm[j - 1])); Tu[j] = -lu * Tu[j - 1]; }
    }
    M2 Psi = {0, 0, 0, 0}, Phi = {0, 0, 0, 0};
    double psiu = 0.0, phiu = 0.0;
    for (int j = m - 1; j >= 0; --j) {
      const int s = s0 + j;
      const M2 U = ld_m2(mf, 8, s, m, G);
      const M2 Sinv = ld_m2(mf, 4, s, m, G);
      const double uu = mf[mf_addr(22, s, m, G)];
      const double sinvu = mf[mf_addr(21, s, m, G)];
      if (j == m - 1) {
        Psi = neg(U);
        psiu = -uu;
        Phi = mul(Sinv, Tm[j]);
        phiu = sinvu * Tu[j];
      } else {
        Psi = neg(mul(U, Psi));
        psiu = -uu * psiu;
        Phi = sub(mul(Sinv, Tm[j]), mul(U, Phi));
        phiu = sinvu * Tu[j] - uu * phiu;
      }
      st_m2(mf, 12, s, m, G, Phi);
      st_m2(mf, 16, s, m, G, Psi);
      mf[mf_addr(23, s, m, G)] = phiu;
      mf[mf_addr(24, s, m, G)] = psiu;
      if (j == 0) {  // level-0 backward scan coefficient: Psi of the chunk's first slot
        sc[sc_addr(0, 6, g, G)] = Psi.a;
        sc[sc_addr(0, 7, g, G)] = Psi.b;
        sc[sc_addr(0, 8, g, G)] = Psi.c;
        sc[sc_addr(0, 9, g, G)] = Psi.d;
        sc[sc_addr(0, 10, g, G)] = psiu;
      }
    }
    // level-0 forward scan coefficient: T of the chunk's last slot
    sc[sc_addr(0, 0, g, G)] = Tm[m - 1].a;
    sc[sc_addr(0, 1, g, G)] = Tm[m - 1].b;
    sc[sc_addr(0, 2, g, G)] = Tm[m - 1].c;
    sc[sc_addr(0, 3, g, G)] = Tm[m - 1].d;
    sc[sc_addr(0, 4, g, G)] = Tu[m - 1];
  }
  // Kogge-Stone products for the higher scan levels
  for (int l = 1; l < P.levels; ++l) {
    const int h = 1 << (l - 1);
    for (int g = 0; g < G; ++g) {
      M2 C = {0, 0, 0, 0}, Cb = {0, 0, 0, 0};
      double cu = 0.0, cbu = 0.0;
      if (g - h >= 0) {
        const M2 x = {sc[sc_addr(l - 1, 0, g, G)], sc[sc_addr(l - 1, 1, g, G)], sc[sc_addr(l - 1, 2, g, G)],
                      sc[sc_addr(l - 1, 3, g, G)]};
        const M2 y = {sc[sc_addr(l - 1, 0, g - h, G)], sc[sc_addr(l - 1, 1, g - h, G)],
                      sc[sc_addr(l - 1, 2, g - h, G)], sc[sc_addr(l - 1, 3, g - h, G)]};
        C = mul(x, y);
        cu = sc[sc_addr(l - 1, 4, g, G)] * sc[sc_addr(l - 1, 4, g - h, G)];
      }
      if (g + h < G) {
        const M2 x = {sc[sc_addr(l - 1, 6, g, G)], sc[sc_addr(l - 1, 7, g, G)], sc[sc_addr(l - 1, 8, g, G)],
                      sc[sc_addr(l - 1, 9, g, G)]};
        const M2 y = {sc[sc_addr(l - 1, 6, g + h, G)], sc[sc_addr(l - 1, 7, g + h, G)],
                      sc[sc_addr(l - 1, 8, g + h, G)], sc[sc_addr(l - 1, 9, g + h, G)]};
        Cb = mul(x, y);
        cbu = sc[sc_addr(l - 1, 10, g, G)] * sc[sc_addr(l - 1, 10, g + h, G)];
      }
      sc[sc_addr(l, 0, g, G)] = C.a;
      sc[sc_addr(l, 1, g, G)] = C.b;
      sc[sc_addr(l, 2, g, G)] = C.c;
      sc[sc_addr(l, 3, g, G)] = C.d;
      sc[sc_addr(l, 4, g, G)] = cu;
      sc[sc_addr(l, 6, g, G)] = Cb.a;
      sc[sc_addr(l, 7, g, G)] = Cb.b;
      sc[sc_addr(l, 8, g, G)] = Cb.c;
      sc[sc_addr(l, 9, g, G)] = Cb.d;
      sc[sc_addr(l, 10, g, G)] = cbu;
    }
  }
  // ---- compact copy for the compact solve (crb_device.cuh::fast_solve_r):
  //   slot part  [pair 0..3][j][g]: (s00, s01), (s11, sinv_u), (t00, t01), (t10, t11) with T = -Lm,
  //              then [jj][g]: (tu_{2jj}, tu_{2jj+1}) with tu = -lu
  //   scan part  [level][pair 0..4][g]: C fwd (2 pairs), C bwd (2 pairs), (cu_fwd, cu_bwd)
  {
    const int lv = P.levels > 0 ? P.levels : 1;
    double* fs = sc + 2 * CRB_SCAN_PAIRS * lv * G;
    double* fc = fs + crb_compact_slot_doubles(m, G);
    for (int s = 0; s < Pn; ++s) {
      const int g = s / m, j = s - g * m;
      const M2 Sinv = ld_m2(mf, 4, s, m, G);
      const M2 Lm = ld_m2(mf, 0, s, m, G);
      const double live = s < P.p_act ? 1.0 : 0.0;  // phantom slots: Sinv = 0 keeps their solution at exactly 0
      fs[(((0 * m + j) * G + g) << 1) + 0] = live * Sinv.a;
      fs[(((0 * m + j) * G + g) << 1) + 1] = live * Sinv.b;
      fs[(((1 * m + j) * G + g) << 1) + 0] = live * Sinv.d;
      fs[(((1 * m + j) * G + g) << 1) + 1] = live * mf[mf_addr(21, s, m, G)];
      fs[(((2 * m + j) * G + g) << 1) + 0] = -Lm.a;
      fs[(((2 * m + j) * G + g) << 1) + 1] = -Lm.b;
      fs[(((3 * m + j) * G + g) << 1) + 0] = -Lm.c;
      fs[(((3 * m + j) * G + g) << 1) + 1] = -Lm.d;
      fs[(((4 * m + (j >> 1)) * G + g) << 1) + (j & 1)] = -mf[mf_addr(20, s, m, G)];
    }
    for (int l = 0; l < P.levels; ++l)
      for (int g = 0; g < G; ++g) {
        double* o = fc + ((long long)(l * 5) * G << 1);
        for (int k = 0; k < 4; ++k) {
          o[(((k >> 1) * G + g) << 1) + (k & 1)] = sc[sc_addr(l, k, g, G)];
          o[(((2 + (k >> 1)) * G + g) << 1) + (k & 1)] = sc[sc_addr(l, 6 + k, g, G)];
        }
        o[((4 * G + g) << 1) + 0] = sc[sc_addr(l, 4, g, G)];
        o[((4 * G + g) << 1) + 1] = sc[sc_addr(l, 10, g, G)];
      }
  }
}
