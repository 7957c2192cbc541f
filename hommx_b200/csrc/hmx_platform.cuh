// Thin platform layer for the hommx_b200 kernels.
//
// On the product path (nvcc / NVRTC, sm_100a) every macro maps 1:1 onto the
// CUDA intrinsic or PTX instruction named beside it.  With HMX_EMULATE defined
// (tests/cpu_emu only, never shipped) the same kernel source is compiled by
// g++ and run with one OS thread per CUDA thread, so the host-side logic of a
// kernel can be checked against the oracle on a machine without a GPU.
#pragma once

#ifdef HMX_EMULATE
#include "cuda_shim.h"  // tests/cpu_emu/cuda_shim.h
#else

#define HMX_DEV __device__ __forceinline__
#define HMX_HOSTDEV __host__ __device__ __forceinline__
#define HMX_KERNEL(maxthreads) extern "C" __global__ void __launch_bounds__(maxthreads, 1)
#define HMX_TKERNEL(maxthreads) __global__ void __launch_bounds__(maxthreads, 1)
#define HMX_SHARED_DECL extern __shared__ __align__(16) double hmx_smem[];
#define HMX_RESTRICT __restrict__

namespace hmx {
HMX_DEV int tid() { return threadIdx.x; }
HMX_DEV int nthreads() { return blockDim.x; }
HMX_DEV int bid() { return blockIdx.x; }
HMX_DEV int nblocks() { return gridDim.x; }
HMX_DEV void sync() { __syncthreads(); }
HMX_DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
HMX_DEV double shfl_down(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
HMX_DEV void atomic_add(double* p, double v) { atomicAdd(p, v); }  // REDG.E.ADD.F64 (global)

// thread-block cluster primitives (sm_90+ PTX; DSMEM between the CTAs that
// split the right-hand sides of one micro cell problem)
HMX_DEV unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
HMX_DEV unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
HMX_DEV void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// load a double from the shared memory of CTA `rank` of this cluster
HMX_DEV double ld_remote(const double* local_smem_ptr, unsigned rank) {
  unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(local_smem_ptr)), ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
}  // namespace hmx
#endif
