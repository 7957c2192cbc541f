// Thin platform layer for the hommx_b200 cell kernels.
//
// On the product path (nvcc, sm_100a) every function below maps 1:1 onto the
// CUDA intrinsic or PTX instruction named beside it.  With HMX_EMULATE defined
// (tests/cpu_emu only, never shipped, never linked into libhmx.so) the same
// kernel source is compiled by g++ and each CUDA thread runs as a fiber, so the
// logic of a kernel can be checked against the oracle on a machine without a
// GPU before GPU time is spent on it.
#pragma once

#ifdef HMX_EMULATE
#include "cuda_shim.h"  // tests/cpu_emu/cuda_shim.h
#else

#define HMX_DEV __device__ __forceinline__
#define HMX_DEV_NOINLINE __device__ __noinline__  // large set-up routines called from several places: one copy of the code
#define HMX_HOSTDEV __host__ __device__ __forceinline__
#define HMX_GLOBAL(maxthreads, minblocks) __global__ void __launch_bounds__(maxthreads, minblocks)
// a kernel launched as thread-block clusters of `cl` CTAs (compile-time cluster size: a plain launch forms the clusters)
#define HMX_GLOBAL_CLUSTER(maxthreads, cl) __global__ void __cluster_dims__(cl, 1, 1) __launch_bounds__(maxthreads, 1)
#define HMX_RESTRICT __restrict__
#define HMX_UNROLL _Pragma("unroll")

namespace hmx {
HMX_DEV int tid() { return threadIdx.x; }
HMX_DEV int bid() { return blockIdx.x; }
HMX_DEV int nblocks() { return gridDim.x; }
HMX_DEV void sync() { __syncthreads(); }  // BAR.SYNC
// named barrier `id` (1..15) over `count` threads (a multiple of 32): BAR.SYNC id, count
HMX_DEV void group_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
HMX_DEV void warp_sync() { __syncwarp(); }
// streaming load that bypasses L1 (the per-point matrix is far larger than L1): LDG.E.64.STRONG.GPU / ld.global.cg
HMX_DEV double ld_stream(const double* p) { return __ldcg(p); }
// 16-byte store (STS.128 / STG.E.128); the address must be 16-byte aligned
HMX_DEV void st_pair(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
// 16-byte load from shared or global memory through the generic path (LDS.128 when p is in shared memory)
HMX_DEV void ld_pair(const double* p, double& a, double& b) {
  const double2 v = *reinterpret_cast<const double2*>(p);
  a = v.x;
  b = v.y;
}
HMX_DEV double* dyn_smem() {
  extern __shared__ __align__(16) double hmx_smem_[];
  return hmx_smem_;
}
// butterfly sum over the 32 lanes of a warp: 5 x SHFL.BFLY (two 32-bit halves each) + DADD
HMX_DEV double warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// sum over aligned segments of `width` lanes (width = 1,2,4,8,16,32); every lane of the warp calls it
HMX_DEV double seg_sum(double v, int width) {
  for (int m = width >> 1; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// value of lane (lane ^ m) of the warp: SHFL.BFLY; every lane of the warp calls it
HMX_DEV double lane_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
HMX_DEV bool warp_any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// ---- mbarrier + bulk async copy (TMA 1-D): the producer/consumer pipeline of the TMA-staged kernel ----
// An MBar occupies HMX_MBAR_BYTES of shared memory (8 used on the device; the CPU emulation needs more).
#define HMX_MBAR_BYTES 32
struct MBar {
  unsigned long long v;
};
HMX_DEV unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
HMX_DEV void mbar_init(MBar* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(b)) : "memory");
}
// make the initialised barriers visible to the async proxy (TMA) before first use; follow with a CTA barrier
HMX_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
HMX_DEV void mbar_arrive(MBar* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
HMX_DEV void mbar_arrive_expect_tx(MBar* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(b)) : "memory");
}
// true when the phase with the given parity has completed
HMX_DEV bool mbar_try_wait(MBar* b, unsigned parity) {
  unsigned done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return done != 0;
}
HMX_DEV void mbar_wait(MBar* b, unsigned parity) {
  while (!mbar_try_wait(b, parity)) {
  }
}
// cp.async.bulk global -> shared (UBLKCP in SASS); bytes multiple of 16, both addresses 16-byte aligned;
// completion is signalled on `b` as transaction bytes
HMX_DEV void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, MBar* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
HMX_DEV void mbar_inval(MBar* b) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(b)) : "memory"); }
// order this thread's generic-proxy writes (global / shared) before later async-proxy (TMA) accesses
HMX_DEV void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }
// a / b for the PCG scalars: MUFU.RCP64H seed + two Newton steps (~6 instructions, <= 2 ulp) instead of the
// ~35-instruction IEEE division sequence with its slow-path branch -- the divisions were a fifth of the
// instructions of the Poisson PCG loop (profiles/r01_p2_inclusion16_raw.txt).  b must be finite and non-zero.
HMX_DEV double fast_div(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return a * r;
}
// 1 / sqrt(x) for finite x > 0: MUFU.RSQ64H seed + three Newton steps (quadratic: 2^-20 -> below 1 ulp)
HMX_DEV double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  HMX_UNROLL
  for (int k = 0; k < 3; ++k) {
    const double e = fma(-x * r, r, 1.0);
    r = fma(0.5 * r, e, r);
  }
  return r;
}
// ---- thread-block clusters + distributed shared memory ----
HMX_DEV int cluster_rank() {  // %cluster_ctarank: this CTA's rank in its cluster
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return (int)r;
}
HMX_DEV int cluster_id() {  // %clusterid.x
  unsigned r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return (int)r;
}
HMX_DEV int nclusters() {  // %nclusterid.x
  unsigned r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return (int)r;
}
// split cluster barrier (UCGABAR_ARV / UCGABAR_WAIT): arrive releases this thread's earlier writes -- also those into
// other CTAs' shared memory -- to the cluster, wait acquires everybody else's.  Every thread of every CTA calls both.
HMX_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
HMX_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
HMX_DEV void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}
// generic address of the same shared-memory location in the CTA of rank `rank` (mapa): plain loads / stores through it
// travel over the SM-to-SM network (DSMEM)
template <class T>
HMX_DEV T* cluster_map(T* p, int rank) {
  unsigned long long out;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((unsigned long long)p), "r"(rank));
  return reinterpret_cast<T*>(out);
}
// shared::cluster address (32-bit) of the location `p` of this CTA's shared memory in the CTA of rank `rank`
HMX_DEV unsigned cluster_map_u32(const void* p, int rank) {
  unsigned out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_u32(p)), "r"(rank));
  return out;
}
// bulk copy of this CTA's shared memory into the shared memory of CTA `rank` (DSMEM, issued to the copy engine by one
// thread: UBLKCP); dst / bar are given as the addresses of the same objects in THIS CTA and are mapped to the peer.
// Completion is signalled on the PEER's mbarrier as transaction bytes.  bytes multiple of 16, addresses 16-byte aligned.
HMX_DEV void bulk_s2c(void* dst, const void* src, unsigned bytes, MBar* bar, int rank) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   cluster_map_u32(dst, rank)),
               "r"(smem_u32(src)), "r"(bytes), "r"(cluster_map_u32(bar, rank))
               : "memory");
}
// one 8-byte store into the shared memory of CTA `rank` that is COUNTED on that CTA's mbarrier (st.async ...
// mbarrier::complete_tx::bytes): the receiver waits on its own mbarrier for the expected number of bytes -- no fence, no
// cluster barrier, no L1 invalidation on either side.  dst / bar: addresses of the same objects in THIS CTA.
HMX_DEV void st_async_f64(double* dst, double v, MBar* bar, int rank) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(cluster_map_u32(dst, rank)),
               "l"(__double_as_longlong(v)), "r"(cluster_map_u32(bar, rank))
               : "memory");
}
HMX_DEV void atomic_add_u64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }  // RED.E.ADD.64
// busy-wait for about `cycles` SM clocks (CS2R on the clock register): staggers the right-hand-side groups of a CTA
HMX_DEV void spin_cycles(long long cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}
}  // namespace hmx
#endif
