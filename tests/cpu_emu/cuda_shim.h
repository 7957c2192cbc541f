// CPU emulation of the few CUDA facilities the hommx_b200 cell kernels use.
// TEST INFRASTRUCTURE ONLY: lets `pytest -m "not gpu"` run the unmodified kernel source
// (hommx_b200/csrc/*.cuh) on the host and compare it with the oracle before GPU time is
// spent.  Nothing under hommx_b200/ includes or links this file.
//
// Every CUDA thread of a CTA is a ucontext fiber; fibers run round-robin, each up to its
// next barrier / warp collective, which reproduces __syncthreads() semantics exactly for
// kernels whose threads all execute the same sequence of barriers (true for ours).
#pragma once
#include <math.h>
#include <stddef.h>

#define HMX_DEV inline
#define HMX_DEV_NOINLINE inline
#define HMX_HOSTDEV inline
#define HMX_RESTRICT __restrict__
#define HMX_UNROLL
#define __device__
#define __forceinline__ inline
#define HMX_GLOBAL(maxthreads, minblocks) void
#define HMX_GLOBAL_CLUSTER(maxthreads, cl) void

namespace hmx {
namespace emu {
struct ClusterState {
  int arrived = 0;
  unsigned gen = 0;
};
struct Cta {
  int nthreads, bid, nblocks, cur;
  // thread-block cluster this CTA belongs to (size 1 for ordinary launches): its CTAs' fibers run in one scheduler
  int crank = 0, csize = 1, fiber_base = 0;
  Cta* peers = nullptr;           // [csize] the CTAs of the cluster, by rank
  ClusterState* cluster = nullptr;
  unsigned* cgen = nullptr;       // [nthreads] barrier generation each fiber arrived in
  double* smem;
  double* wbuf;  // [2][nthreads] warp collective exchange
  int* wpar;     // [nthreads] per-fiber parity
  int arrived[64];       // barrier id -> fibers waiting
  unsigned gen[64];      // barrier id -> generation
};
extern thread_local Cta* g_cta;
void yield();  // return to the scheduler until every fiber has arrived
}  // namespace emu

inline int tid() { return emu::g_cta->cur; }
inline int bid() { return emu::g_cta->bid; }
inline int nblocks() { return emu::g_cta->nblocks; }
// barrier `id` over `count` fibers (ids 0..15 mirror the hardware named barriers, 16.. are the
// per-warp rendezvous of the shuffle emulation)
inline void barrier(int id, int count) {
  emu::Cta* c = emu::g_cta;
  const unsigned g = c->gen[id];
  if (++c->arrived[id] == count) {
    c->arrived[id] = 0;
    c->gen[id] = g + 1;
  } else {
    while (c->gen[id] == g) emu::yield();
  }
}
inline void sync() { barrier(0, emu::g_cta->nthreads); }
inline void group_sync(int id, int count) { barrier(id, count); }
inline void warp_sync() { barrier(16 + emu::g_cta->cur / 32, 32); }
inline double ld_stream(const double* p) { return *p; }
inline void ld_pair(const double* p, double& a, double& b) {
  a = p[0];
  b = p[1];
}
inline void st_pair(double* p, double a, double b) {
  p[0] = a;
  p[1] = b;
}
inline double* dyn_smem() { return emu::g_cta->smem; }
inline double warp_sum(double v) {
  emu::Cta* c = emu::g_cta;
  const int me = c->cur, par = c->wpar[me];
  c->wbuf[par * c->nthreads + me] = v;
  c->wpar[me] = par ^ 1;
  barrier(16 + me / 32, 32);
  const int w0 = (me / 32) * 32;
  // same association as the xor butterfly of the device code
  double t[32];
  for (int l = 0; l < 32; ++l) t[l] = c->wbuf[par * c->nthreads + w0 + l];
  for (int m = 16; m > 0; m >>= 1) {
    double u[32];
    for (int l = 0; l < 32; ++l) u[l] = t[l] + t[l ^ m];
    for (int l = 0; l < 32; ++l) t[l] = u[l];
  }
  return t[me & 31];
}
inline double seg_sum(double v, int width) {
  emu::Cta* c = emu::g_cta;
  const int me = c->cur, par = c->wpar[me];
  c->wbuf[par * c->nthreads + me] = v;
  c->wpar[me] = par ^ 1;
  barrier(16 + me / 32, 32);
  const int w0 = (me / 32) * 32;
  double t[32];
  for (int l = 0; l < 32; ++l) t[l] = c->wbuf[par * c->nthreads + w0 + l];
  for (int m = width >> 1; m > 0; m >>= 1) {
    double u[32];
    for (int l = 0; l < 32; ++l) u[l] = t[l] + t[l ^ m];
    for (int l = 0; l < 32; ++l) t[l] = u[l];
  }
  return t[me & 31];
}
inline double lane_xor(double v, int m) {
  emu::Cta* c = emu::g_cta;
  const int me = c->cur, par = c->wpar[me];
  c->wbuf[par * c->nthreads + me] = v;
  c->wpar[me] = par ^ 1;
  barrier(16 + me / 32, 32);
  return c->wbuf[par * c->nthreads + (me / 32) * 32 + ((me & 31) ^ m)];
}
inline bool warp_any(bool p) {
  emu::Cta* c = emu::g_cta;
  const int me = c->cur, par = c->wpar[me];
  c->wbuf[par * c->nthreads + me] = p ? 1.0 : 0.0;
  c->wpar[me] = par ^ 1;
  barrier(16 + me / 32, 32);
  const int w0 = (me / 32) * 32;
  bool any = false;
  for (int l = 0; l < 32; ++l) any = any || c->wbuf[par * c->nthreads + w0 + l] != 0.0;
  return any;
}
// ---- mbarrier + bulk copy emulation: phase bit, pending arrivals, outstanding transaction bytes ----
#define HMX_MBAR_BYTES 32
struct MBar {
  int init_count, pending;
  long long tx;
  unsigned phase;
};
inline void mbar_maybe_complete(MBar* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->phase ^= 1u;
    b->pending = b->init_count;
  }
}
inline void mbar_init(MBar* b, int count) {
  b->init_count = b->pending = count;
  b->tx = 0;
  b->phase = 0;
}
inline void mbar_fence_init() {}
inline void mbar_arrive(MBar* b) {
  --b->pending;
  mbar_maybe_complete(b);
}
inline void mbar_arrive_expect_tx(MBar* b, unsigned bytes) {
  b->tx += bytes;
  --b->pending;
  mbar_maybe_complete(b);
}
inline bool mbar_try_wait(MBar* b, unsigned parity) { return b->phase != parity; }
inline void mbar_wait(MBar* b, unsigned parity) {
  while (!mbar_try_wait(b, parity)) emu::yield();
}
inline void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, MBar* b) {
  __builtin_memcpy(smem_dst, gmem_src, bytes);  // lands "instantly"; ordering bugs must be caught on the device
  b->tx -= bytes;
  mbar_maybe_complete(b);
}
inline void mbar_inval(MBar*) {}
inline void fence_async_proxy() {}
inline double fast_div(double a, double b) { return a / b; }
inline double fast_rsqrt(double x) { return 1.0 / sqrt(x); }
// ---- clusters: a cluster's CTAs live in one address space here, DSMEM is a pointer into the peer's buffer ----
inline int cluster_rank() { return emu::g_cta->crank; }
inline int cluster_id() { return emu::g_cta->bid / emu::g_cta->csize; }
inline int nclusters() { return emu::g_cta->nblocks / emu::g_cta->csize; }
inline void cluster_arrive() {
  emu::Cta* c = emu::g_cta;
  c->cgen[c->cur] = c->cluster->gen;
  if (++c->cluster->arrived == c->csize * c->nthreads) {
    c->cluster->arrived = 0;
    ++c->cluster->gen;
  }
}
inline void cluster_wait() {
  emu::Cta* c = emu::g_cta;
  while (c->cluster->gen == c->cgen[c->cur]) emu::yield();
}
inline void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}
template <class T>
inline T* cluster_map(T* p, int rank) {
  emu::Cta* c = emu::g_cta;
  const size_t off = (size_t)((const char*)p - (const char*)c->smem);
  return reinterpret_cast<T*>((char*)c->peers[rank].smem + off);
}
inline void bulk_s2c(void* dst, const void* src, unsigned bytes, MBar* bar, int rank) {
  __builtin_memcpy(cluster_map(static_cast<char*>(dst), rank), src, bytes);  // lands "instantly" (see bulk_g2s)
  MBar* rb = cluster_map(bar, rank);
  rb->tx -= bytes;
  mbar_maybe_complete(rb);
}
inline void st_async_f64(double* dst, double v, MBar* bar, int rank) {
  *cluster_map(dst, rank) = v;
  MBar* rb = cluster_map(bar, rank);
  rb->tx -= 8;
  mbar_maybe_complete(rb);
}
inline void atomic_add_u64(unsigned long long* p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline void spin_cycles(long long) {}  // timing only: nothing to emulate
}  // namespace hmx
