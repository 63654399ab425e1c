// crb_tile.cuh -- asynchronous state tiles for the persistent fused integrators (sm_100a).
//
// A warp integrates `mpw` consecutive members at a time (a "tile").  Because the ensemble state is member-major
// (X[B, 2n], the vector the reference integrates, dynamic_beam_model.py:120-149), a tile is ONE contiguous byte
// range of X (and of the per-member stiffness coefficients), so it moves with 1-D bulk copies of the TMA unit:
//
//   load   cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes   (global -> the warp's `in` buffer)
//   store  cp.async.bulk.global.shared::cta.bulk_group                         (the warp's `out` buffer -> global)
//
// Each warp owns one mbarrier, one `in` and one `out` buffer.  While tile i integrates in registers, tile i+1 is
// already in flight into the `in` buffer (it was freed when tile i was read into registers) and tile i-1's result
// drains from the `out` buffer: the per-tile prologue / epilogue that a one-tile-per-block kernel exposes on every
// wave is hidden behind the arithmetic of the neighbouring tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned crb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void crb_mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// make the initialised barrier visible to the async proxy
__device__ __forceinline__ void crb_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void crb_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void crb_mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void crb_mbar_wait(unsigned bar, unsigned parity) {
  unsigned done;
  do {  // try_wait blocks for a hardware-defined time slice before it reports failure
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared, completion counted in bytes on the mbarrier; src / dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void crb_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global as part of the thread's current bulk group
__device__ __forceinline__ void crb_bulk_store(void* dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void crb_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the newest N groups have finished READING shared memory (the buffer may be overwritten)
template <int N>
__device__ __forceinline__ void crb_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void crb_bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
