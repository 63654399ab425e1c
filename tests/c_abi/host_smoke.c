/* Plain-C client of include/crb.h (test infrastructure): proves that the header compiles as C99 and that the
 * host-only entry points work without Python or PyTorch.  Prints a few values that tests/test_host_api.py
 * compares with the same calls made through ctypes.  No device work is done here. */
#include <stdio.h>
#include <string.h>

#include "crb.h"

int main(void) {
  enum { N = 4 };
  uint8_t bc[N + 1] = {CRB_BC_FIXED, CRB_BC_NONE, CRB_BC_NONE, CRB_BC_NONE, CRB_BC_NONE};
  uint8_t et[N] = {CRB_ELEM_LINEAR, CRB_ELEM_LINEAR, CRB_ELEM_LINEAR, CRB_ELEM_LINEAR};
  double par[N][CRB_NPARAM];
  double M[12 * 12], K[12 * 12];
  crb_plan_t plan;
  int i, j, rc;
  double trM = 0.0, trK = 0.0, asym = 0.0;

  printf("version %d\n", crb_version());
  rc = crb_plan(N, bc, 0, &plan);
  if (rc) { printf("crb_plan failed: %s\n", crb_last_error()); return 1; }
  printf("plan n_free %d m %d g %d p %d contiguous %d\n", plan.n_free, plan.m, plan.g, plan.p, plan.contiguous);
  for (i = 0; i < N; ++i) { /* examples/example_utilities.py:25-34 */
    par[i][CRB_P_LENGTH] = 0.25;
    par[i][CRB_P_E] = 75e9;
    par[i][CRB_P_I] = 4.908738521234052e-10;
    par[i][CRB_P_RHO] = 6450.0;
    par[i][CRB_P_AREA] = 7.853981633974483e-05;
    par[i][CRB_P_WETTED] = 0.007853981633974483;
    par[i][CRB_P_CD] = 0.82;
  }
  rc = crb_dense_matrices(&plan, &par[0][0], et, bc, M, K);
  if (rc) { printf("crb_dense_matrices failed: %s\n", crb_last_error()); return 1; }
  for (i = 0; i < 12; ++i) {
    trM += M[i * 12 + i];
    trK += K[i * 12 + i];
    for (j = 0; j < 12; ++j) {
      double d = K[i * 12 + j] - K[j * 12 + i];
      if (d < 0) d = -d;
      if (d > asym) asym = d;
    }
  }
  printf("trace_M %.17g\ntrace_K %.17g\nasym_K %.3g\n", trM, trK, asym);
  /* error path: a bad boundary code is refused with a message, not a crash */
  bc[2] = 7;
  rc = crb_plan(N, bc, 0, &plan);
  printf("bad_bc rc %d msg_len %d\n", rc, (int)strlen(crb_last_error()));
  return 0;
}
