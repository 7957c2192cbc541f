// Fiber scheduler behind cuda_shim.h (TEST INFRASTRUCTURE ONLY).
#pragma once
#include <ucontext.h>

#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "cuda_shim.h"

namespace hmx {
namespace emu {
thread_local Cta* g_cta = nullptr;

struct Sched {
  ucontext_t main_ctx;
  std::vector<ucontext_t> ctx;
  std::vector<char> done;
  std::vector<char*> stacks;
  void (*body)(void*);
  void* arg;
};
thread_local Sched* g_sched = nullptr;

void yield() {
  Sched* s = g_sched;
  swapcontext(&s->ctx[g_cta->fiber_base + g_cta->cur], &s->main_ctx);
}

static void fiber_entry() {
  Sched* s = g_sched;
  s->body(s->arg);
  s->done[g_cta->fiber_base + g_cta->cur] = 1;
  swapcontext(&s->ctx[g_cta->fiber_base + g_cta->cur], &s->main_ctx);
}

// run one thread-block cluster (csize CTAs of nthreads fibers each, one scheduler, one address space); csize = 1 is
// an ordinary CTA.  bid0 = block index of rank 0.
inline void run_cluster(int csize, int nthreads, int bid0, int nblocks, size_t smem_doubles, void (*body)(void*), void* arg) {
  constexpr size_t STACK = 256 * 1024;
  const int nfib = csize * nthreads;
  // shared memory is NOT zero-initialised on the device: HMX_EMU_POISON=1 fills it with NaNs (0xfff7... bit
  // patterns, huge as integers) so that a read of a slot nobody wrote shows up as a NaN result or a wild index
  const char* poison = std::getenv("HMX_EMU_POISON");
  double fill = 0.0;
  if (poison && poison[0] == '1') {
    const unsigned long long bits = 0xfff7dead7ff7beefULL;
    std::memcpy(&fill, &bits, sizeof fill);
  }
  std::vector<Cta> ctas(csize);
  std::vector<std::vector<double>> smem(csize), wbuf(csize);
  std::vector<std::vector<int>> wpar(csize);
  std::vector<std::vector<unsigned>> cgen(csize);
  ClusterState cstate;
  for (int r = 0; r < csize; ++r) {
    Cta& cta = ctas[r];
    cta.nthreads = nthreads;
    cta.bid = bid0 + r;
    cta.nblocks = nblocks;
    cta.cur = 0;
    cta.crank = r;
    cta.csize = csize;
    cta.fiber_base = r * nthreads;
    cta.peers = ctas.data();
    cta.cluster = &cstate;
    std::memset(cta.arrived, 0, sizeof cta.arrived);
    std::memset(cta.gen, 0, sizeof cta.gen);
    smem[r].assign(smem_doubles + 2, fill);
    wbuf[r].assign(2 * nthreads, 0.0);
    wpar[r].assign(nthreads, 0);
    cgen[r].assign(nthreads, 0u);
    cta.smem = smem[r].data();
    cta.wbuf = wbuf[r].data();
    cta.wpar = wpar[r].data();
    cta.cgen = cgen[r].data();
  }
  Sched s;
  s.body = body;
  s.arg = arg;
  s.ctx.resize(nfib);
  s.done.assign(nfib, 0);
  s.stacks.resize(nfib);
  g_sched = &s;
  for (int t = 0; t < nfib; ++t) {
    s.stacks[t] = (char*)std::malloc(STACK);
    getcontext(&s.ctx[t]);
    s.ctx[t].uc_stack.ss_sp = s.stacks[t];
    s.ctx[t].uc_stack.ss_size = STACK;
    s.ctx[t].uc_link = &s.main_ctx;
    makecontext(&s.ctx[t], fiber_entry, 0);
  }
  // Scheduling order of the fibers between barriers.  A kernel without data races gives bit-identical results
  // under every order (tests/test_emu_parity.py::test_results_do_not_depend_on_the_thread_schedule):
  //   HMX_EMU_ORDER unset / "forward": 0, 1, 2, ...   "reverse": n-1, ..., 0   "shuffle:<seed>": a new random
  //   permutation in every round  (for a cluster the order runs over the fibers of all its CTAs)
  const char* mode = std::getenv("HMX_EMU_ORDER");
  const bool reverse = mode != nullptr && std::strcmp(mode, "reverse") == 0;
  const bool shuffle = mode != nullptr && std::strncmp(mode, "shuffle", 7) == 0;
  unsigned long long rng = 0x9E3779B97F4A7C15ull ^ (shuffle && mode[7] == ':' ? std::strtoull(mode + 8, nullptr, 10) : 0ull) ^
                           ((unsigned long long)bid0 << 32);
  std::vector<int> order(nfib);
  for (int t = 0; t < nfib; ++t) order[t] = reverse ? nfib - 1 - t : t;
  bool alive = true;
  while (alive) {
    alive = false;
    if (shuffle)
      for (int t = nfib - 1; t > 0; --t) {
        rng = rng * 6364136223846793005ull + 1442695040888963407ull;
        const int r = (int)((rng >> 33) % (unsigned long long)(t + 1));
        const int tmp = order[t];
        order[t] = order[r];
        order[r] = tmp;
      }
    for (int k = 0; k < nfib; ++k) {
      const int t = order[k];
      if (s.done[t]) continue;
      g_cta = &ctas[t / nthreads];
      g_cta->cur = t % nthreads;
      swapcontext(&s.main_ctx, &s.ctx[t]);
      if (!s.done[t]) alive = true;
    }
  }
  for (int t = 0; t < nfib; ++t) std::free(s.stacks[t]);
  g_cta = nullptr;
  g_sched = nullptr;
}

// run a grid (of clusters of `csize` CTAs; grid is rounded down to whole clusters), clusters spread over host threads
inline void run_grid(int grid, int nthreads, size_t smem_doubles, void (*body)(void*), void* arg, int host_threads, int csize = 1) {
  const int ncl = grid / csize > 0 ? grid / csize : 1;
  grid = ncl * csize;
  if (host_threads < 1) host_threads = 1;
  if (host_threads > ncl) host_threads = ncl;
  std::vector<std::thread> pool;
  for (int w = 0; w < host_threads; ++w)
    pool.emplace_back([=]() {
      for (int b = w; b < ncl; b += host_threads) run_cluster(csize, nthreads, b * csize, grid, smem_doubles, body, arg);
    });
  for (auto& th : pool) th.join();
}
}  // namespace emu
}  // namespace hmx
