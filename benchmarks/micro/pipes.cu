// pipes.cu -- B200 microbenchmark: do FP64 FMA (DFMA) and FP64 tensor MMA (DMMA) overlap?  What do
// SHFL / LDS cost next to them?  Standalone (nvcc -arch=sm_100a pipes.cu -o pipes); prints one JSON
// line per experiment.  Used to steer the mass-solve design.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// NF dfma per iteration per thread (independent chains), NM dmma.m8n8k4 per iteration per warp, NS 64-bit shuffles
template <int NF, int NM, int NS, int NL>
__global__ void __launch_bounds__(256) mix_kernel(double* out, int iters, double seed) {
  __shared__ double sm[256 * 2];
  sm[threadIdx.x] = seed + threadIdx.x;
  sm[threadIdx.x + 256] = seed;
  __syncthreads();
  double f[NF > 0 ? NF : 1];
  double m0[NM > 0 ? NM : 1], m1[NM > 0 ? NM : 1];
  double s[NS > 0 ? NS : 1];
  double l[NL > 0 ? NL : 1];
#pragma unroll
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { m0[i] = seed; m1[i] = seed + 1; }
#pragma unroll
  for (int i = 0; i < (NS > 0 ? NS : 1); ++i) s[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < (NL > 0 ? NL : 1); ++i) l[i] = 0.0;
  const double a = seed * 0.5, b = seed * 0.25;
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
    for (int i = 0; i < NM; ++i) dmma884(m0[i], m1[i], a, b);
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = __shfl_xor_sync(0xffffffffu, s[i], 1 + (i & 3));
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const double2 v = *reinterpret_cast<const double2*>(sm + 2 * ((idx + i) & 255));
      l[i] += v.x;
      idx = (idx + (int)v.y) & 255;   // v.y == seed == 0 at run time: keeps the load in the loop
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) r += f[i];
#pragma unroll
  for (int i = 0; i < NM; ++i) r += m0[i] + m1[i];
#pragma unroll
  for (int i = 0; i < NS; ++i) r += s[i];
#pragma unroll
  for (int i = 0; i < NL; ++i) r += l[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int NM>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double seed) {
  double d[NM][4];
#pragma unroll
  for (int i = 0; i < NM; ++i) { d[i][0] = seed; d[i][1] = seed; d[i][2] = seed; d[i][3] = seed; }
  const double a[4] = {seed, seed * 0.5, seed * 0.25, seed * 0.125};
  const double b[2] = {seed, seed * 0.5};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NM; ++i) dmma1688(d[i], a, b);
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) r += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F>
float time_it(F launch) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) launch();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 5;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int sms = p.multiProcessorCount;
  const int blocks = sms * 4, threads = 256, iters = 20000;
  double* out; CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
  const double warps_per_sm = 4.0 * threads / 32;
  auto report = [&](const char* name, float ms, double instr_per_iter_per_warp, const char* what) {
    // warp-instructions per clock per SM at the nominal max clock
    const double clk = ms * 1e-3 * clk_khz * 1e3;
    printf("{\"exp\": \"%s\", \"ms\": %.4f, \"clk_per_iter_per_sm\": %.3f, \"%s_per_clk_per_sm\": %.4f}\n", name, ms,
           clk / iters, what, instr_per_iter_per_warp * warps_per_sm * iters / clk);
  };
#define RUN(NF, NM, NS, NL, name, cnt, what) { float ms = time_it([&] { mix_kernel<NF, NM, NS, NL><<<blocks, threads>>>(out, iters, 0.0); }); report(name, ms, cnt, what); }
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, clk_khz);
  RUN(16, 0, 0, 0, "dfma16", 16, "dfma_warp_instr")
  RUN(0, 4, 0, 0, "dmma884x4", 4, "dmma884_warp_instr")
  RUN(0, 8, 0, 0, "dmma884x8", 8, "dmma884_warp_instr")
  RUN(16, 2, 0, 0, "dfma16+dmma2", 16, "dfma_warp_instr")
  RUN(16, 4, 0, 0, "dfma16+dmma4", 16, "dfma_warp_instr")
  RUN(32, 4, 0, 0, "dfma32+dmma4", 32, "dfma_warp_instr")
  RUN(0, 0, 8, 0, "shfl64x8", 8, "shfl64_warp_instr")
  RUN(16, 0, 4, 0, "dfma16+shfl4", 16, "dfma_warp_instr")
  RUN(0, 0, 0, 8, "lds128x8", 8, "lds128_warp_instr")
  RUN(16, 0, 4, 2, "dfma16+shfl4+lds2", 16, "dfma_warp_instr")
  { float ms = time_it([&] { dmma1688_kernel<4><<<blocks, threads>>>(out, iters, 0.0); }); report("dmma1688x4", ms, 4, "dmma1688_warp_instr"); }
  CK(cudaDeviceSynchronize());
  return 0;
}
