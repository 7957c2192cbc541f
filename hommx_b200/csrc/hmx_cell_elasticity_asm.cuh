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
// Thread i owns node i and all right-hand sides: y = K p in registers, p and r in shared memory
// (node-major, 16-byte loads of the neighbours' p), x and the matrix in the L2-resident scratch.
// Only the diagonal block and the blocks of the 2^D-1 POSITIVE stencil directions are stored
// (K_{i,i-d} = K_{i-d,i}^T is read from the neighbour's row): 80 doubles per node, 328 KB per 8^3
// cell, 48 MB for 148 CTAs -- the full rows (82 MB) did not stay in the two-partition L2 (31 % of
// the matrix reads went to DRAM, profiles/r01_c4_v4a_raw.txt).  Layout [direction][pair][node][2]:
// a warp reads 512 consecutive bytes with one LDG.128 per thread.  No colouring, no scatter: the
// apply is a pure gather.  Preconditioner, stopping rule and epilogue are those of
// the matrix-free kernel (block Jacobi; A_hom = <C> - b_p.x_q - x_p.r_q).
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
  static constexpr int o_atoms = o_red + 2 * NW * NREDV;   // [NA][T][NRC]
  static constexpr int o_dinv = ((o_atoms + NA1 * T * NRC + 1) / 2) * 2;  // [NSYM][N]
  static constexpr int o_p = ((o_dinv + NSYM * N + 1) / 2) * 2;           // [N][NVEC]
  static constexpr int o_r = o_p + N * NVEC;                              // [N][NVEC]
  static constexpr int total = o_r + N * NVEC;
  static constexpr int NBP = (NB + 1) / 2;                  // block entries padded to pairs (LDG.128)
  static constexpr int KDOUBLES = (1 + NH) * NBP * 2 * N;   // diagonal + positive directions
  static constexpr int scratch_doubles = KDOUBLES + 2 * N * NVEC;  // matrix, load vectors, correctors; per CTA
  static_assert(NT >= N && NT % 32 == 0, "one thread per node");
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
  double* s_atoms = sm + L::o_atoms;
  double* s_dinv = sm + L::o_dinv;
  double* s_p = sm + L::o_p;
  double* s_r = sm + L::o_r;
  constexpr int NBP = L::NBP;
  double* g_K = P.scratch + (size_t)bid() * L::scratch_doubles;  // [1+NH][NBP][N][2]
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

    // ---- 2. assemble the row of node i: one stencil direction at a time, blocks in registers ----
    double r[NVEC];  // starts as the load vectors b_q[i]
    HMX_UNROLL
    for (int k = 0; k < NVEC; ++k) r[k] = 0.0;
    if (own) {
      HMX_UNROLL
      for (int d = 0; d <= NH; ++d) {  // diagonal and positive directions; the rest are transposes
        double blk[2 * NBP];
        HMX_UNROLL
        for (int k = 0; k < 2 * NBP; ++k) blk[k] = 0.0;
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
        for (int k = 0; k < NBP; ++k) st_pair(g_K + ((size_t)(d * NBP + k) * N + i) * 2, blk[2 * k], blk[2 * k + 1]);
        if (d == 0) {
          double sym[NSYM], inv[NSYM];
          HMX_UNROLL
          for (int j = 0; j < D; ++j)
            HMX_UNROLL
            for (int j2 = j; j2 < D; ++j2) sym[sym_index(D, j, j2)] = blk[j * D + j2];
          sym_inverse<D>(sym, inv);
          HMX_UNROLL
          for (int k = 0; k < NSYM; ++k) s_dinv[k * N + i] = inv[k];
        }
      }
      HMX_UNROLL
      for (int k = 0; k < NVEC; ++k) g_b[k * N + i] = r[k];
    }

    // ---- 3. PCG on all right-hand sides; r, y in registers, p, x in shared memory ----
    double rz[NRHS], rz0[NRHS];
    bool active[NRHS];
    {
      double part[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) part[q] = 0.0;
      if (own) {
        double di[NSYM];
        HMX_UNROLL
        for (int k = 0; k < NSYM; ++k) di[k] = s_dinv[k * N + i];
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
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);  // publishes p and the matrix rows
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
    while (any && it < P.max_it) {
      ++it;
      double y[NVEC], pAp[NRHS];
      HMX_UNROLL
      for (int k = 0; k < NVEC; ++k) y[k] = 0.0;
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) pAp[q] = 0.0;
      if (own) {
        // software pipeline over the stencil directions: the 16-byte loads of direction d+2 are
        // issued before the FMAs of direction d (L2 latency ~ 2 directions of math per warp)
        // d = 0 diagonal, 1..NH node i + mask (own block), NH+1..2NH node i - mask (neighbour's block, transposed)
        constexpr int DEPTH = 3;
        double kb[DEPTH][2 * NBP];
        int jnb[NSTEN];
        HMX_UNROLL
        for (int d = 0; d < NSTEN; ++d)
          jnb[d] = d == 0 ? i : (d > NH ? G::template shifted<-1>(c, d - NH) : G::template shifted<1>(c, d));
#define HMX_LOADK(d_)                                                                                            \
  {                                                                                                              \
    const int dd_ = (d_) > NH ? (d_)-NH : (d_);                                                                  \
    const int row_ = (d_) > NH ? jnb[d_] : i;                                                                    \
    HMX_UNROLL                                                                                                   \
    for (int k = 0; k < NBP; ++k)                                                                                \
      ld_stream_pair(g_K + ((size_t)(dd_ * NBP + k) * N + row_) * 2, kb[(d_) % DEPTH][2 * k], kb[(d_) % DEPTH][2 * k + 1]); \
  }
        HMX_UNROLL
        for (int d = 0; d < DEPTH - 1; ++d) HMX_LOADK(d);
        HMX_UNROLL
        for (int d = 0; d < NSTEN; ++d) {
          if (d + DEPTH - 1 < NSTEN) HMX_LOADK(d + DEPTH - 1);
          const bool tr = d > NH;
          const double* pn = s_p + jnb[d] * NVEC;
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            double pj[D];
            HMX_UNROLL
            for (int j2 = 0; j2 < D; ++j2) pj[j2] = pn[q * D + j2];
            HMX_UNROLL
            for (int j = 0; j < D; ++j)
              HMX_UNROLL
              for (int j2 = 0; j2 < D; ++j2)
                y[q * D + j] += (tr ? kb[d % DEPTH][j2 * D + j] : kb[d % DEPTH][j * D + j2]) * pj[j2];
          }
        }
#undef HMX_LOADK
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
        double di[NSYM];
        HMX_UNROLL
        for (int k = 0; k < NSYM; ++k) di[k] = s_dinv[k * N + i];
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
