// Micro cell kernel for the linear-elasticity HMM classes, ASSEMBLED + TMA-STAGED variant (opt-in).
//
// Same operator representation as hmx_cell_elasticity_asm.cuh (diagonal + positive-direction blocks of
// the periodic block stencil in an L2-resident per-CTA buffer, one thread per node, all right-hand
// sides per thread), but the blocks reach shared memory through a multi-stage ring filled by the TMA
// unit: one bulk copy (cp.async.bulk, 9 x N doubles = one stencil direction) per stage, issued S-1
// stages ahead by lane 0 of the warps in turn (a dedicated producer warp would cap every thread at 96
// registers: 17 warps put 5 on one scheduler) and signalled on an mbarrier; the warps wait on "full", do
//   y_i += B_d[i] p_{i+d} + B_d[i-d]^T p_{i-d}
// from the stage and release it on "empty".  There is no CTA-wide barrier inside the apply, so the
// warps drift apart and shared-memory reads, FP64 math and the L2 stream of different stages overlap
// (the barrier-per-stage version keeps all warps in the same phase: 19.6 us per iteration).
// p lives in shared memory; r, x, b in the L2 scratch (streamed once per iteration by their owners).
#pragma once
#include "hmx_cell_common.cuh"
#include "hmx_cell_elasticity_asm.cuh"  // asm_assemble_row, sym_inverse

namespace hmx {

template <class CO, int NM, int NT>
struct ElasticityTmaLayout {
  static constexpr int D = CO::DIM;
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int N = Grid<D, NM>::N;  // consumer threads = nodes
  static constexpr int NRHS = D * (D + 1) / 2;
  static constexpr int NV = NRHS;
  static constexpr int NVEC = NRHS * D;
  static constexpr int NH = (1 << D) - 1;
  static constexpr int NDIR = NH + 1;  // stored directions = stages per iteration
  static constexpr int NB = D * D;
  static constexpr int NWC = N / 32;  // consumer warps
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  static constexpr int NSYM = D * (D + 1) / 2;
  static constexpr int NRC = AtomIdx<D, NM, CO::YDEP>::NRC;
  static constexpr int NREDV = 2 * NRHS > NA1 ? 2 * NRHS : NA1;
  static constexpr int SMAX = 4;
  static constexpr int STAGE = NB * N;  // doubles per stage
  static constexpr int o_bar = 0;       // full[SMAX], empty[SMAX]
  static constexpr int o_red = (2 * SMAX * HMX_MBAR_BYTES + 64) / 8;  // 2 buffers [NWC][NREDV]
  static constexpr int o_atoms = o_red + 2 * NWC * NREDV;              // [NA][T][NRC]
  static constexpr int o_p = ((o_atoms + NA1 * T * NRC + 1) / 2) * 2;  // [N][NVEC]
  static constexpr int o_ring = ((o_p + N * NVEC + 15) / 16) * 16;     // S x [NB][N], 128-byte aligned
  static constexpr int S_FIT = (227 * 1024 / 8 - o_ring) / STAGE;
  static constexpr int S = S_FIT < SMAX ? S_FIT : SMAX;
  static constexpr int total = o_ring + S * STAGE;
  static constexpr int KDOUBLES = NDIR * NB * N;
  static constexpr int scratch_doubles = KDOUBLES + 3 * N * NVEC;  // matrix, b, x, r per CTA
  static_assert(N % 32 == 0 && NT == N, "one thread per node");
  static_assert(S >= 2, "the ring needs at least two stages");
  static_assert((STAGE * 8) % 16 == 0, "bulk copies move multiples of 16 bytes");
};

// block_sum over the consumer threads only (named barrier 1)
template <int NV, int NW>
HMX_DEV void consumer_sum(double (&v)[NV], double* buf, int nthreads) {
  const int lane = tid() & 31, warp = tid() >> 5;
  HMX_UNROLL
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
    HMX_UNROLL
    for (int k = 0; k < NV; ++k) buf[warp * NV + k] = v[k];
  }
  group_sync(1, nthreads);
  HMX_UNROLL
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int w = 0; w < NW; ++w) s += buf[w * NV + k];
    v[k] = s;
  }
}

template <class CO, int NM, int NT>
HMX_DEV void elasticity_tma_cell_body(const CellParams& P) {
  using L = ElasticityTmaLayout<CO, NM, NT>;
  using G = Grid<CO::DIM, NM>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP>;
  constexpr int D = L::D, T = L::T, N = L::N, NRHS = L::NRHS, NV = L::NV, NVEC = L::NVEC, NH = L::NH, NDIR = L::NDIR;
  constexpr int NB = L::NB, NWC = L::NWC, NA = L::NA, NA1 = L::NA1, NSYM = L::NSYM, NRC = L::NRC, S = L::S, STAGE = L::STAGE;
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;
  constexpr unsigned STAGE_BYTES = STAGE * 8;

  double* sm = dyn_smem();
  MBar* full = reinterpret_cast<MBar*>(reinterpret_cast<char*>(sm));
  MBar* empty = reinterpret_cast<MBar*>(reinterpret_cast<char*>(sm) + L::SMAX * HMX_MBAR_BYTES);
#define HMX_FULL(b_) reinterpret_cast<MBar*>(reinterpret_cast<char*>(full) + (b_)*HMX_MBAR_BYTES)
#define HMX_EMPTY(b_) reinterpret_cast<MBar*>(reinterpret_cast<char*>(empty) + (b_)*HMX_MBAR_BYTES)
  double* s_red = sm + L::o_red;
  double* s_atoms = sm + L::o_atoms;
  double* s_p = sm + L::o_p;
  double* s_ring = sm + L::o_ring;
  double* g_K = P.scratch + (size_t)bid() * L::scratch_doubles;  // [NDIR][NB][N]
  double* g_b = g_K + L::KDOUBLES;                               // [NVEC][N]
  double* g_x = g_b + N * NVEC;
  double* g_r = g_x + N * NVEC;

  const int t_id = tid();
  const bool consumer = t_id < N;
  const int i = consumer ? t_id : 0;
  const int lane = t_id & 31;
  const double h = 1.0 / (double)NM;
  const double vol = (D == 2 ? 0.5 * h * h : h * h * h / 6.0);
  int red_flip = 0;
  int c[3];
  G::decode(i, c);
  int jp[NDIR], jm[NDIR];
  HMX_UNROLL
  for (int d = 1; d <= NH; ++d) {
    jp[d] = G::template shifted<1>(c, d);
    jm[d] = G::template shifted<-1>(c, d);
  }
  jp[0] = jm[0] = i;

  for (long long pt = bid(); pt < P.n_pts; pt += nblocks()) {
    double xm[3], verts[(D + 1) * 3];
    macro_point<D>(P, pt, xm, verts);
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double Mn[D * D];
    CO::dtheta(xm, Mn);
    HMX_UNROLL
    for (int k = 0; k < D * D; ++k) Mn[k] *= (double)NM;

    // ---- 1. atoms, barriers ----
    if (t_id == 0) {
      for (int b = 0; b < S; ++b) {
        mbar_init(HMX_FULL(b), 1);        // the producer's arrive.expect_tx (+ the copy's transaction bytes)
        mbar_init(HMX_EMPTY(b), NWC);     // one arrival per consumer warp
      }
      mbar_fence_init();
    }
    if (NA > 0) {
      for (int idx = t_id; idx < T * NRC; idx += NT) {
        const int t = idx / NRC, rc = idx - t * NRC;
        int cc[3];
        AI::rdecode(rc, cc);
        double acc[NA1];
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) acc[k] = 0.0;
        for (int qq = 0; qq < P.nq; ++qq) {
          double y[D], s[NA1];
          HMX_UNROLL
          for (int a = 0; a < D; ++a) y[a] = ((double)cc[a] + P.qp[(t * P.nq + qq) * D + a]) * h;
          CO::atoms(pc, y, s);
          const double wq = P.qw[qq];
          HMX_UNROLL
          for (int k = 0; k < NA1; ++k) acc[k] += wq * s[k];
        }
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) s_atoms[(k * T + t) * NRC + rc] = acc[k];
      }
    }
    sync();  // #1 (all threads)

    // ---- 2. consumers assemble their rows, producer waits ----
    double smean[NA1];
    HMX_UNROLL
    for (int k = 0; k < NA1; ++k) smean[k] = 0.0;
    double di[NSYM], r[NVEC];
    HMX_UNROLL
    for (int k = 0; k < NSYM; ++k) di[k] = 0.0;
    HMX_UNROLL
    for (int k = 0; k < NVEC; ++k) r[k] = 0.0;
    if (consumer) {
      if (NA > 0) {
        for (int idx = t_id; idx < T * NRC; idx += N) {
          HMX_UNROLL
          for (int k = 0; k < NA; ++k) smean[k] += s_atoms[k * T * NRC + idx];
        }
        consumer_sum<NA1, NWC>(smean, s_red + (red_flip ^= 1) * NWC * L::NREDV, N);
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) smean[k] *= 1.0 / (double)(T * NRC);
      }
      asm_assemble_row<CO, NM>(c, i, pc, Mn, s_atoms, vol, g_K, di, r);
      HMX_UNROLL
      for (int k = 0; k < NVEC; ++k) {
        g_b[k * N + i] = r[k];
        g_r[k * N + i] = r[k];
        g_x[k * N + i] = 0.0;
      }
      fence_async_proxy();  // the matrix rows were written by ordinary stores; the TMA unit reads them next
    }
    sync();  // #2: the matrix is complete

    {
      // stage g (global count) = stencil direction g % NDIR in ring buffer g % S.  It is issued while stage
      // g - (S-1) is processed, by lane 0 of warp g % NWC, after every warp has released the buffer.
      auto issue = [&](long long g) {
        if (lane == 0 && (t_id >> 5) == (int)(g % NWC)) {
          const int b = (int)(g % S);
          mbar_wait(HMX_EMPTY(b), (unsigned)((g / S) & 1) ^ 1u);  // first lap: passes at once
          mbar_arrive_expect_tx(HMX_FULL(b), STAGE_BYTES);
          bulk_g2s(s_ring + (size_t)b * STAGE, g_K + (size_t)(g % NDIR) * STAGE, STAGE_BYTES, HMX_FULL(b));
        }
      };
      for (long long g = 0; g < S - 1; ++g) issue(g);
      // ================= consumers: PCG on all right-hand sides =================
      double rz[NRHS], rz0[NRHS];
      bool active[NRHS];
      {
        double part[NRHS];
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          part[q] = 0.0;
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            double z = 0.0;
            HMX_UNROLL
            for (int j2 = 0; j2 < D; ++j2) z += di[sym_index(D, j, j2)] * r[q * D + j2];
            part[q] += r[q * D + j] * z;
            s_p[i * NVEC + q * D + j] = z;
          }
        }
        consumer_sum<NRHS, NWC>(part, s_red + (red_flip ^= 1) * NWC * L::NREDV, N);  // publishes p
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          rz[q] = rz0[q] = part[q];
          active[q] = part[q] > P.atol * P.atol;
        }
      }
      int it = 0, its[NRHS];
      bool any = false;
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        its[q] = 0;
        any = any || active[q];
      }
      long long gc = 0;  // stages consumed so far
      while (any && it < P.max_it) {
        ++it;
        double y[NVEC];
        HMX_UNROLL
        for (int k = 0; k < NVEC; ++k) y[k] = 0.0;
        for (int s = 0; s < NDIR; ++s, ++gc) {
          issue(gc + S - 1);  // the matrix does not change: prefetching runs across iteration boundaries
          const int b = (int)(gc % S);
          mbar_wait(HMX_FULL(b), (unsigned)((gc / S) & 1));
          const double* kb = s_ring + (size_t)b * STAGE;
          {
            double kk[NB], pj[NVEC];
            HMX_UNROLL
            for (int k = 0; k < NB; ++k) kk[k] = kb[k * N + i];
            const double* pn = s_p + jp[s] * NVEC;
            HMX_UNROLL
            for (int k = 0; k < NVEC; k += 2) ld_pair(pn + k, pj[k], pj[k + 1]);
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q)
              HMX_UNROLL
              for (int j = 0; j < D; ++j)
                HMX_UNROLL
                for (int j2 = 0; j2 < D; ++j2) y[q * D + j] += kk[j * D + j2] * pj[q * D + j2];
          }
          if (s > 0) {
            double kk[NB], pj[NVEC];
            HMX_UNROLL
            for (int k = 0; k < NB; ++k) kk[k] = kb[k * N + jm[s]];
            const double* pn = s_p + jm[s] * NVEC;
            HMX_UNROLL
            for (int k = 0; k < NVEC; k += 2) ld_pair(pn + k, pj[k], pj[k + 1]);
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q)
              HMX_UNROLL
              for (int j = 0; j < D; ++j)
                HMX_UNROLL
                for (int j2 = 0; j2 < D; ++j2) y[q * D + j] += kk[j2 * D + j] * pj[q * D + j2];
          }
          warp_sync();
          if (lane == 0) mbar_arrive(HMX_EMPTY(b));  // this warp is done with the stage
        }
        double pAp[NRHS];
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          pAp[q] = 0.0;
          HMX_UNROLL
          for (int j = 0; j < D; ++j) pAp[q] += s_p[i * NVEC + q * D + j] * y[q * D + j];
        }
        HMX_UNROLL
        for (int k = 0; k < NVEC; ++k) r[k] = g_r[k * N + i];  // in flight during the reduction
        consumer_sum<NRHS, NWC>(pAp, s_red + (red_flip ^= 1) * NWC * L::NREDV, N);
        double alpha[NRHS], part[NRHS], z[NVEC];
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          alpha[q] = (active[q] && pAp[q] > 0.0) ? rz[q] / pAp[q] : 0.0;
          part[q] = 0.0;
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            if (active[q]) g_x[(q * D + j) * N + i] += alpha[q] * s_p[i * NVEC + q * D + j];
            r[q * D + j] -= alpha[q] * y[q * D + j];
            g_r[(q * D + j) * N + i] = r[q * D + j];
          }
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            z[q * D + j] = 0.0;
            HMX_UNROLL
            for (int j2 = 0; j2 < D; ++j2) z[q * D + j] += di[sym_index(D, j, j2)] * r[q * D + j2];
            part[q] += r[q * D + j] * z[q * D + j];
          }
        }
        consumer_sum<NRHS, NWC>(part, s_red + (red_flip ^= 1) * NWC * L::NREDV, N);
        any = false;
        double beta[NRHS];
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          beta[q] = 0.0;
          if (active[q]) {
            beta[q] = part[q] / rz[q];
            rz[q] = part[q];
            const double tol = fmax(P.rtol * P.rtol * rz0[q], P.atol * P.atol);
            if (!(part[q] > tol)) active[q] = false;
            its[q] = it;
          }
          any = any || active[q];
        }
        if (any) {
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q)
            if (active[q]) {
              HMX_UNROLL
              for (int j = 0; j < D; ++j) s_p[i * NVEC + q * D + j] = z[q * D + j] + beta[q] * s_p[i * NVEC + q * D + j];
            }
          group_sync(1, N);  // publish p to the other consumers
        }
      }
      // the S-1 stages issued ahead of the last iteration must land before the ring is reused
      for (long long g = gc; g < gc + S - 1; ++g)
        if (lane == 0 && (t_id >> 5) == (int)(g % NWC)) mbar_wait(HMX_FULL((int)(g % S)), (unsigned)((g / S) & 1));

      // ---- epilogue: A_hom[p][q] = <C>[p][q] - b_p.x_q - x_p.r_q ----
      double Ah[NRHS * NRHS];
      for (int p = 0; p < NRHS; ++p) {
        double zz[2 * NRHS];
        HMX_UNROLL
        for (int k = 0; k < 2 * NRHS; ++k) zz[k] = 0.0;
        HMX_UNROLL
        for (int j = 0; j < D; ++j) {
          const double bp = g_b[(p * D + j) * N + i], xp = g_x[(p * D + j) * N + i];
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            zz[q] += bp * g_x[(q * D + j) * N + i];
            zz[NRHS + q] += xp * g_r[(q * D + j) * N + i];
          }
        }
        consumer_sum<2 * NRHS, NWC>(zz, s_red + (red_flip ^= 1) * NWC * L::NREDV, N);
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) Ah[p * NRHS + q] = -zz[q] - zz[NRHS + q];
      }
      if (P.chi != nullptr) {
        for (int k = 0; k < NVEC; ++k) P.chi[((size_t)pt * NVEC + k) * N + i] = g_x[k * N + i];
      }
      if (t_id == 0) {
        for (int q = 0; q < NRHS; ++q) {
          double e[NV], sg[NV];
          HMX_UNROLL
          for (int v = 0; v < NV; ++v) e[v] = (v == q) ? 1.0 : 0.0;
          CO::stress(pc, smean, e, sg);
          for (int p = 0; p < NRHS; ++p) Ah[p * NRHS + q] += sg[p];
        }
        if (P.A_hom != nullptr)
          for (int k = 0; k < NRHS * NRHS; ++k) P.A_hom[pt * NRHS * NRHS + k] = Ah[k];
        if (P.S_loc != nullptr) macro_element_matrix<D, 1>(verts, Ah, P.S_loc + pt * (D + 1) * D * (D + 1) * D);
        if (P.iters != nullptr) P.iters[pt] = it;
        double worst = 0.0;
        unsigned long long tot = 0;
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          if (rz0[q] > P.atol * P.atol) worst = fmax(worst, sqrt(rz[q] / rz0[q]));
          tot += (unsigned long long)its[q];
        }
        if (P.resid != nullptr) P.resid[pt] = worst;
        if (P.work != nullptr) atomic_add_u64(P.work, tot);
      }
    }
    sync();  // #3: producer drained, consumers done: barriers and shared memory can be reused
    if (t_id == 0) {
      for (int b = 0; b < S; ++b) {
        mbar_inval(HMX_FULL(b));
        mbar_inval(HMX_EMPTY(b));
      }
    }
    sync();  // #4
  }
#undef HMX_FULL
#undef HMX_EMPTY
}

}  // namespace hmx
