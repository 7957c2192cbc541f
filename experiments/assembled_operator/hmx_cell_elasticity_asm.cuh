// Micro cell kernel for the linear-elasticity HMM classes, ASSEMBLED variant: one CTA per macro
// point, the periodic P1 stiffness matrix of the point is assembled once into an L2-resident
// per-CTA buffer and every PCG iteration streams it back.
//
// Why: the matrix-free element kernel (hmx_cell_elasticity.cuh) executes ~4.5x the FLOPs of the
// assembled stencil (612 vs 135 FMA per node and right-hand side in 3-D), which caps its
// algorithmic FP64 fraction near 25 %.  The assembled rows (D x D blocks on the 7-/15-point
// stencil: 135 doubles per node, 553 KB for an 8^3 cell) do not fit in shared memory -- but 148
// of them fit in B200's 126 MB L2, and L2 delivers 553 KB per SM in 6.1 us when every SM streams
// its own buffer (scripts/micro/l2bw.cu: 13.7 TB/s aggregate), against 16.5 us per iteration of the
// matrix-free kernel.  Each block is loaded once per iteration and reused by all D(D+1)/2
// right-hand sides from registers (54 FMA per 9 loaded doubles).
//
// Thread i owns node i and all right-hand sides: y = K p in registers (18 doubles), p and r in
// shared memory (node-major, 16-byte loads of the neighbours' p), x, b and the matrix in the
// L2-resident scratch.  Only the diagonal block and the blocks of the 2^D-1 POSITIVE stencil
// directions are stored (K_{i,i-d} = K_{i-d,i}^T): 72 doubles per node, 295 KB per 8^3 cell, 44 MB for
// 148 CTAs -- full rows (82 MB) did not stay in the two-partition L2 (31 % of the matrix reads went
// to DRAM, profiles/r01_c4_v4a_raw.txt).
//
// The apply is STAGED through shared memory, one stencil direction d (9 x N doubles, 36 KB) at a
// time in a two-buffer ring: while the CTA computes  y_i += B_d[i] p_{i+d} + B_d[i-d]^T p_{i-d}  from
// the current buffer -- so every block fetched from L2 is used twice --, each thread already holds
// its 9 doubles of direction d+1 in flight (coalesced LDG issued before the math) and stores them
// into the other buffer afterwards; one BAR.SYNC per stage.  Loading straight into registers
// (first version) left the L2 latency exposed: 16 warps at 128 registers cannot keep enough
// loads in flight (long_scoreboard 5.1 stall cycles per issue, profiles/r01_c4_v4b_raw.txt).
// No colouring, no scatter: the apply is a pure gather.
#pragma once
#include "hmx_cell_common.cuh"
#include "hmx_cell_elasticity.cuh"  // sym_inverse

namespace hmx {

template <class CO, int NM, int NT>
struct ElasticityAsmLayout {
  static constexpr int D = CO::DIM;
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int N = Grid<D, NM>::N;
  static constexpr int NRHS = D * (D + 1) / 2;
  static constexpr int NV = NRHS;
  static constexpr int NVEC = NRHS * D;           // values per node: all right-hand sides x components
  static constexpr int NH = (1 << D) - 1;         // positive stencil directions
  static constexpr int NSTEN = 2 * NH + 1;        // 7 / 15
  static constexpr int NB = D * D;                // block entries
  static constexpr int NW = NT / 32;
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  static constexpr int NSYM = D * (D + 1) / 2;
  static constexpr int NRC = AtomIdx<D, NM, CO::YDEP>::NRC;
  static constexpr int NREDV = 2 * NRHS > NA1 ? 2 * NRHS : NA1;
  static constexpr int o_red = 0;                          // 2 buffers [NW][NREDV]
  static constexpr int NBQ = ((NB + 1) / 2) * 2;                          // block entries padded to 16-byte pairs
  static constexpr int o_ring = ((o_red + 2 * NW * NREDV + 1) / 2) * 2;   // 2 x [N][NBQ] staged matrix blocks
  static constexpr int o_atoms = o_ring;                                  // [NA][T][NRC]: dead once the matrix exists
  static constexpr int RING = 2 * NBQ * N > NA1 * T * NRC ? 2 * NBQ * N : NA1 * T * NRC;
  static constexpr int o_p = o_ring + RING;                               // [N][NVEC]
  static constexpr int o_r = o_p + N * NVEC;                              // [N][NVEC]
  static constexpr int total = o_r + N * NVEC;
  static constexpr int KDOUBLES = (1 + NH) * NB * N;        // [direction][entry][node]: diagonal + positive directions
  static constexpr int scratch_doubles = KDOUBLES + 2 * N * NVEC;  // matrix, load vectors, correctors; per CTA
  static_assert(NT >= N && NT % 32 == 0, "one thread per node");
  static_assert(NVEC % 2 == 0, "the per-node vectors are moved with 16-byte accesses");
};

// stencil slot of the edge from local vertex a to local vertex b of a type-t simplex
template <int D>
HMX_HOSTDEV constexpr int sten_slot(int t, int a, int b) {
  return a == b ? 0
                : (b > a ? (kuhn_pmask<D>(t, b) & ~kuhn_pmask<D>(t, a))
                         : ((1 << D) - 1) + (kuhn_pmask<D>(t, a) & ~kuhn_pmask<D>(t, b)));
}

// engineering-Voigt strain of the vector basis function phi e_j whose mapped gradient is m
template <int D>
HMX_DEV void basis_strain(const double (&m)[D], int j, double (&e)[D * (D + 1) / 2]) {
  HMX_UNROLL
  for (int v = 0; v < D; ++v) e[v] = (v == j) ? m[v] : 0.0;
  int v = D;
  HMX_UNROLL
  for (int r = 0; r < D; ++r)
    HMX_UNROLL
    for (int c = r + 1; c < D; ++c) {
      e[v] = ((c == j) ? m[r] : 0.0) + ((r == j) ? m[c] : 0.0);
      ++v;
    }
}

// Row of node i of the periodic stiffness matrix, one stencil direction at a time (blocks in registers):
// diagonal + positive directions go to g_K as [direction][entry][node]; also returns the inverse diagonal
// block `di` and the load vectors `r` (b_q[i] for every right-hand side).
template <class CO, int NM>
HMX_DEV void asm_assemble_row(const int (&c)[3], int i, const double* pc, const double (&Mn)[CO::DIM * CO::DIM],
                              const double* s_atoms, double vol, double* g_K, double (&di)[CO::DIM * (CO::DIM + 1) / 2],
                              double (&r)[CO::DIM * CO::DIM * (CO::DIM + 1) / 2]) {
  using G = Grid<CO::DIM, NM>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP>;
  constexpr int D = CO::DIM, T = kuhn_ntypes<D>(), N = G::N, NRHS = D * (D + 1) / 2, NV = NRHS, NH = (1 << D) - 1;
  constexpr int NB = D * D, NA = CO::NATOMS, NA1 = NA > 0 ? NA : 1, NSYM = D * (D + 1) / 2, NRC = AI::NRC;
  HMX_UNROLL
  for (int d = 0; d <= NH; ++d) {  // diagonal and positive directions; the rest are transposes
    double blk[NB];
    HMX_UNROLL
    for (int k = 0; k < NB; ++k) blk[k] = 0.0;
    HMX_UNROLL
    for (int t = 0; t < T; ++t) {
      HMX_UNROLL
      for (int a = 0; a <= D; ++a) {
        // does the type-t simplex in which node i is vertex a have an edge in direction d ?
        bool any_b = false;
        HMX_UNROLL
        for (int b = 0; b <= D; ++b) any_b = any_b || sten_slot<D>(t, a, b) == d;
        if (!any_b) continue;
        int o[3];
        G::template shift_coords<-1>(c, kuhn_pmask<D>(t, a), o);
        const int ro = AI::ridx(o);
        double sa[NA1];
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) sa[k] = NA > 0 ? s_atoms[(k * T + t) * NRC + ro] : 0.0;
        double ma[D];
        HMX_UNROLL
        for (int p = 0; p < D; ++p) {
          ma[p] = 0.0;
          if (a >= 1) ma[p] += Mn[p * D + kuhn_axis<D>(t, a >= 1 ? a - 1 : 0)];
          if (a < D) ma[p] -= Mn[p * D + kuhn_axis<D>(t, a < D ? a : 0)];
        }
        double sg[D][NV];
        HMX_UNROLL
        for (int j = 0; j < D; ++j) {
          double ea[NV];
          basis_strain<D>(ma, j, ea);
          CO::stress(pc, sa, ea, sg[j]);
          if (d == 0) {
            // load vectors ride along with the diagonal block: b_q[i][j] -= |e| (C E_q) : e(phi_a e_j)
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) r[q * D + j] -= vol * sg[j][q];  // (C ea)[q] = ea : C : E_q
          }
        }
        HMX_UNROLL
        for (int b = 0; b <= D; ++b) {
          if (sten_slot<D>(t, a, b) != d) continue;
          double mb[D];
          HMX_UNROLL
          for (int p = 0; p < D; ++p) {
            mb[p] = 0.0;
            if (b >= 1) mb[p] += Mn[p * D + kuhn_axis<D>(t, b >= 1 ? b - 1 : 0)];
            if (b < D) mb[p] -= Mn[p * D + kuhn_axis<D>(t, b < D ? b : 0)];
          }
          HMX_UNROLL
          for (int j2 = 0; j2 < D; ++j2) {
            double eb[NV];
            basis_strain<D>(mb, j2, eb);
            HMX_UNROLL
            for (int j = 0; j < D; ++j) {
              double s = 0.0;
              HMX_UNROLL
              for (int v = 0; v < NV; ++v) s += sg[j][v] * eb[v];
              blk[j * D + j2] += vol * s;
            }
          }
        }
      }
    }
    HMX_UNROLL
    for (int k = 0; k < NB; ++k) g_K[(size_t)(d * NB + k) * N + i] = blk[k];
    if (d == 0) {
      double sym[NSYM], inv[NSYM];
      HMX_UNROLL
      for (int j = 0; j < D; ++j)
        HMX_UNROLL
        for (int j2 = j; j2 < D; ++j2) sym[sym_index(D, j, j2)] = blk[j * D + j2];
      sym_inverse<D>(sym, inv);
      HMX_UNROLL
      for (int k = 0; k < NSYM; ++k) di[k] = inv[k];
    }
  }
}

template <class CO, int NM, int NT>
HMX_DEV void elasticity_asm_cell_body(const CellParams& P) {
  using L = ElasticityAsmLayout<CO, NM, NT>;
  using G = Grid<CO::DIM, NM>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP>;
  constexpr int D = L::D, T = L::T, N = L::N, NRHS = L::NRHS, NV = L::NV, NVEC = L::NVEC, NH = L::NH, NSTEN = L::NSTEN;
  constexpr int NB = L::NB, NW = L::NW, NA = L::NA, NA1 = L::NA1, NSYM = L::NSYM, NRC = L::NRC;
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;

  double* sm = dyn_smem();
  double* s_red = sm + L::o_red;
  double* s_atoms = sm + L::o_atoms;  // aliases the ring: only alive during the assembly
  double* s_ring = sm + L::o_ring;    // 2 x [N][NBQ]
  constexpr int NBQ = L::NBQ;
  double* s_p = sm + L::o_p;
  double* s_r = sm + L::o_r;
  double* g_K = P.scratch + (size_t)bid() * L::scratch_doubles;  // [1+NH][NB][N]
  double* g_b = g_K + L::KDOUBLES;                               // [NVEC][N]
  double* g_x = g_b + N * NVEC;                                  // [NVEC][N]

  const int i = tid();
  const bool own = i < N;
  const double h = 1.0 / (double)NM;
  const double vol = (D == 2 ? 0.5 * h * h : h * h * h / 6.0);
  int red_flip = 0;
  int c[3] = {0, 0, 0};
  if (own) G::decode(i, c);

  for (long long pt = bid(); pt < P.n_pts; pt += nblocks()) {
    double xm[3], verts[(D + 1) * 3];
    macro_point<D>(P, pt, xm, verts);
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double Mn[D * D];  // n M
    CO::dtheta(xm, Mn);
    HMX_UNROLL
    for (int k = 0; k < D * D; ++k) Mn[k] *= (double)NM;

    // ---- 1. atoms (natural reduced layout) ----
    if (NA > 0) {
      for (int idx = i; idx < T * NRC; idx += NT) {
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
    sync();
    double smean[NA1];
    HMX_UNROLL
    for (int k = 0; k < NA1; ++k) smean[k] = 0.0;
    if (NA > 0) {
      for (int idx = i; idx < T * NRC; idx += NT) {
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) smean[k] += s_atoms[k * T * NRC + idx];
      }
      block_sum<NA1, NW>(smean, s_red + (red_flip ^= 1) * NW * L::NREDV);
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) smean[k] *= 1.0 / (double)(T * NRC);
    }

    // ---- 2. assemble the row of node i (diagonal + positive directions), inverse diagonal block, loads ----
    double di[NSYM];  // inverse diagonal block of node i (thread-private -> registers)
    HMX_UNROLL
    for (int k = 0; k < NSYM; ++k) di[k] = 0.0;
    double r[NVEC];  // starts as the load vectors b_q[i]
    HMX_UNROLL
    for (int k = 0; k < NVEC; ++k) r[k] = 0.0;
    if (own) {
      asm_assemble_row<CO, NM>(c, i, pc, Mn, s_atoms, vol, g_K, di, r);
      HMX_UNROLL
      for (int k = 0; k < NVEC; ++k) g_b[k * N + i] = r[k];
    }
    sync();  // every thread is done with the atoms: their storage becomes the ring

    // ---- 3. PCG on all right-hand sides; y in registers, p, r in shared memory, x in the scratch ----
    double rz[NRHS], rz0[NRHS];
    bool active[NRHS];
    {
      double part[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) part[q] = 0.0;
      if (own) {
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q)
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            double z = 0.0;
            HMX_UNROLL
            for (int j2 = 0; j2 < D; ++j2) z += di[sym_index(D, j, j2)] * r[q * D + j2];
            part[q] += r[q * D + j] * z;
            s_p[i * NVEC + q * D + j] = z;
            s_r[i * NVEC + q * D + j] = r[q * D + j];
            g_x[(q * D + j) * N + i] = 0.0;
          }
        // stage 0 (the diagonal blocks) of the first iteration
        HMX_UNROLL
        for (int k = 0; k < NB; ++k) s_ring[i * NBQ + k] = g_K[(size_t)k * N + i];
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);  // publishes p and ring buffer 0
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
    // neighbours of node i: +mask and -mask for every positive direction
    int jp[NH + 1], jm[NH + 1];
    HMX_UNROLL
    for (int d = 1; d <= NH; ++d) {
      jp[d] = G::template shifted<1>(c, d);
      jm[d] = G::template shifted<-1>(c, d);
    }
    jp[0] = jm[0] = i;
    constexpr int NSTAGE = NH + 1;  // stage s holds direction s; buffer s & 1 (NSTAGE is even: 4 or 8)
    static_assert(NSTAGE % 2 == 0, "the ring parity must repeat from one iteration to the next");
    while (any && it < P.max_it) {
      ++it;
      double y[NVEC], pAp[NRHS];
      HMX_UNROLL
      for (int k = 0; k < NVEC; ++k) y[k] = 0.0;
      HMX_UNROLL
      for (int s = 0; s < NSTAGE; ++s) {
        // (a) the loads of the next stage (next iteration's diagonal after the last one) go in flight
        const int nxt = (s + 1) % NSTAGE;
        double kn[NB];
        if (own) {
          HMX_UNROLL
          for (int k = 0; k < NB; ++k) kn[k] = ld_stream(g_K + (size_t)(nxt * NB + k) * N + i);
        }
        // (b) math of this stage from ring buffer s & 1 (16-byte shared-memory loads)
        if (own) {
          const double* kb = s_ring + (s & 1) * NBQ * N;
          {
            double kk[NBQ], pj[NVEC];
            HMX_UNROLL
            for (int k = 0; k < NBQ; k += 2) ld_pair(kb + i * NBQ + k, kk[k], kk[k + 1]);
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
          if (s > 0) {  // the same direction seen from node i - d: the neighbour's block, transposed
            double kk[NBQ], pj[NVEC];
            HMX_UNROLL
            for (int k = 0; k < NBQ; k += 2) ld_pair(kb + jm[s] * NBQ + k, kk[k], kk[k + 1]);
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
          // (c) park the next stage in the other buffer (its readers finished a barrier ago)
          double* kw = s_ring + ((s + 1) & 1) * NBQ * N + i * NBQ;
          HMX_UNROLL
          for (int k = 0; k + 1 < NB; k += 2) st_pair(kw + k, kn[k], kn[k + 1]);
          if (NB % 2) kw[NB - 1] = kn[NB - 1];
        }
        if (s + 1 < NSTAGE) sync();  // after the last stage the reduction's barrier does the job
      }
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) pAp[q] = 0.0;
      if (own) {
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q)
          HMX_UNROLL
          for (int j = 0; j < D; ++j) pAp[q] += s_p[i * NVEC + q * D + j] * y[q * D + j];
      }
      block_sum<NRHS, NW>(pAp, s_red + (red_flip ^= 1) * NW * L::NREDV);
      double alpha[NRHS], part[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        alpha[q] = (active[q] && pAp[q] > 0.0) ? rz[q] / pAp[q] : 0.0;
        part[q] = 0.0;
      }
      double z[NVEC];
      HMX_UNROLL
      for (int k = 0; k < NVEC; ++k) z[k] = 0.0;
      if (own) {
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            if (active[q]) g_x[(q * D + j) * N + i] += alpha[q] * s_p[i * NVEC + q * D + j];
            r[q * D + j] = s_r[i * NVEC + q * D + j] - alpha[q] * y[q * D + j];
            s_r[i * NVEC + q * D + j] = r[q * D + j];
          }
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            HMX_UNROLL
            for (int j2 = 0; j2 < D; ++j2) z[q * D + j] += di[sym_index(D, j, j2)] * r[q * D + j2];
            part[q] += r[q * D + j] * z[q * D + j];
          }
        }
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);
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
        if (own) {
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q)
            if (active[q]) {
              HMX_UNROLL
              for (int j = 0; j < D; ++j) s_p[i * NVEC + q * D + j] = z[q * D + j] + beta[q] * s_p[i * NVEC + q * D + j];
            }
        }
        sync();
      }
    }

    // ---- 4. epilogue: A_hom[p][q] = <C>[p][q] - b_p.x_q - x_p.r_q ----
    {
      double Ah[NRHS * NRHS];
      for (int p = 0; p < NRHS; ++p) {
        double zz[2 * NRHS];
        HMX_UNROLL
        for (int k = 0; k < 2 * NRHS; ++k) zz[k] = 0.0;
        if (own) {
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            const double bp = g_b[(p * D + j) * N + i], xp = g_x[(p * D + j) * N + i];
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) {
              zz[q] += bp * g_x[(q * D + j) * N + i];
              zz[NRHS + q] += xp * s_r[i * NVEC + q * D + j];
            }
          }
        }
        block_sum<2 * NRHS, NW>(zz, s_red + (red_flip ^= 1) * NW * L::NREDV);
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) Ah[p * NRHS + q] = -zz[q] - zz[NRHS + q];
      }
      if (P.chi != nullptr && own) {
        for (int k = 0; k < NVEC; ++k) P.chi[((size_t)pt * NVEC + k) * N + i] = g_x[k * N + i];
      }
      if (i == 0) {
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
    sync();
  }
}

}  // namespace hmx
