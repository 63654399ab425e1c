// lds.cu -- shared-memory load throughput on B200: LDS.128 / LDS.64 / LDS.32, distinct vs broadcast addresses.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// MODE 0: every lane its own 16 B (conflict-free); MODE 1: 8 distinct 16 B words per warp (4 lanes share: the
// lane-group layout of crb_rk4_fast.cuh); MODE 2: all lanes the same address
template <int W, int MODE>
__global__ void __launch_bounds__(256) lds_kernel(unsigned long long* out, int iters) {
  __shared__ __align__(16) unsigned long long sm[2048];
  for (int k = threadIdx.x; k < 2048; k += 256) sm[k] = k * 0x9E3779B97F4A7C15ull;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int idx = MODE == 0 ? lane : (MODE == 1 ? (lane & 7) : 0);
  unsigned long long acc = 0;
  int base = warp * 64;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int a = (base + i * 37 + idx) & 511;   // 16-byte word index
      if (W == 16) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(sm + 2 * a);
        acc ^= v.x + v.y;
      } else if (W == 8) {
        acc ^= sm[2 * a];
      } else {
        acc ^= reinterpret_cast<const unsigned*>(sm)[4 * a];
      }
    }
    base += 3;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
float time_it(F launch) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) launch();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 5;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int blocks = p.multiProcessorCount * 4, iters = 20000;
  unsigned long long* out; CK(cudaMalloc(&out, 8ull * blocks * 256));
#define RUN(W, MODE) { float ms = time_it([&] { lds_kernel<W, MODE><<<blocks, 256>>>(out, iters); }); \
    const double clk = ms * 1e-3 * clk_khz * 1e3; \
    printf("{\"width\": %d, \"mode\": %d, \"ms\": %.4f, \"clk_per_warp_load\": %.3f}\n", W, MODE, ms, clk / (8.0 * iters * 32)); }
  RUN(16, 0) RUN(16, 1) RUN(16, 2) RUN(8, 0) RUN(8, 1) RUN(8, 2) RUN(4, 0) RUN(4, 2)
  return 0;
}
