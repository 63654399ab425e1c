// crb_probe.cu -- machine probe used by bench.py: the FP64 FMA rate of the device, measured in the same run as the
// kernels it is the roofline denominator of (MEASURED_PEAKS.json holds HBM and bf16 figures only; FP64 vector math and
// the FP64 tensor cores share one pipe on B200, see DESIGN.md, so this one number bounds every kernel of the library).
#include "crb_internal.h"

#define CRB_PROBE_CHAINS 8
#define CRB_PROBE_THREADS 256
#define CRB_PROBE_BLOCKS_PER_SM 8

__global__ void __launch_bounds__(CRB_PROBE_THREADS) crb_dfma_probe_kernel(int iters, double* __restrict__ out) {
  double a[CRB_PROBE_CHAINS];
#pragma unroll
  for (int k = 0; k < CRB_PROBE_CHAINS; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + k);
  const double x = 1.0 - 1e-9, y = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CRB_PROBE_CHAINS; ++k) a[k] = fma(a[k], x, y);  // independent dependency chains
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CRB_PROBE_CHAINS; ++k) s += a[k];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int crb_probe_dfma(int32_t iters, double* scratch, int64_t scratch_doubles, int64_t* flops_out, void* stream) {
  if (iters < 1 || !flops_out) return crb_fail(CRB_E_ARG, "crb_probe_dfma: iters >= 1 and flops_out are required");
  const int blocks = crb_sm_count() * CRB_PROBE_BLOCKS_PER_SM;
  const long long threads = (long long)blocks * CRB_PROBE_THREADS;
  *flops_out = threads * CRB_PROBE_CHAINS * 2ll * iters;
  if (!scratch) return 0;  // query: scratch doubles needed = flops / (16 iters)
  if (scratch_doubles < threads)
    return crb_fail(CRB_E_ARG, "crb_probe_dfma: scratch holds %lld doubles, %lld needed", (long long)scratch_doubles, threads);
  crb_dfma_probe_kernel<<<blocks, CRB_PROBE_THREADS, 0, (cudaStream_t)stream>>>(iters, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return crb_fail(CRB_E_CUDA, "crb_probe_dfma: launch failed: %s", cudaGetErrorString(e));
  return 0;
}
