// Micro cell kernel for the scalar (Poisson) HMM classes: one CTA per macro quadrature point.
//
// Replaces, for PoissonHMM / PoissonStratifiedHMM, everything the reference does inside
// BaseHMM._compute_local_stiffness (hmm.py:334-369): coefficient evaluation, assembly of
// the periodic micro stiffness matrix (cell_problem.py:367-369), the corrector solves
// (cell_problem.py:384) and the n_b^2 assemble_scalar passes (hmm.py:361-364), in the
// equivalent d-RHS / A_hom formulation (hmm.py:1219-1245).
//
// Per macro point (all data stays on chip):
//   1. atoms      : y-dependent scalars of A(x, y) averaged per micro element with the
//                   quadrature rule handed in by the host (no coefficient array in HBM).
//   2. operator   : the periodic P1 stiffness matrix has the 7-/15-point stencil of the
//                   Kuhn mesh; its symmetric half (+ the diagonal) is built in shared
//                   memory from the atoms and a small per-point table
//                   kap[k][t][a<b] = |e| g_a^T (M^T C_k M) g_b.
//   3. PCG        : Jacobi-preconditioned CG in FP64 on all D right-hand sides at once.
//                   A thread owns NPT nodes: x, r, b live in its registers, only the search
//                   directions p sit in shared memory; dot products are warp-shuffle +
//                   one-barrier block reductions.
//   4. epilogue   : A_hom[p][q] = <A>[p][q] - b_p.x_q - x_p.r_q   (error quadratic in the
//                   residual), optionally S_loc = |T| G^T A_hom G for the macro cell.
#pragma once
#include "hmx_cell_common.cuh"

namespace hmx {

// AGLOB = 1: the per-element atoms live in the L2 scratch instead of shared memory (cells with many atoms
// on a fully y-dependent coefficient, e.g. 4 atoms at n >= 10 in 3-D, exceed 227 KB otherwise).
template <class CO, int NM, int NT, int COLL = 0, int AGLOB = 0>
struct PoissonLayout {
  static constexpr int D = CO::DIM;
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int N = Grid<D, NM, COLL>::N;
  static constexpr int NRHS = D;
  static constexpr int NH = (1 << D) - 1;  // stored (positive) stencil directions
  static constexpr int NW = NT / 32;
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  static constexpr int NPAIR = D * (D + 1) / 2;
  static constexpr int NSYM = D * (D + 1) / 2;
  static constexpr int NRC = AtomIdx<D, NM, CO::YDEP>::NRC;
  static constexpr int NRED = 2 * NRHS * NRHS > NA1 ? 2 * NRHS * NRHS : NA1;
  // offsets in doubles
  static constexpr int o_red = 0;                                    // 2 buffers
  static constexpr int o_vt = o_red + 2 * NW * NRED;                 // [(D+1)][3] macro cell vertices (for the epilogue)
  static constexpr int o_bk = o_vt + (D + 1) * 3;                    // [(1+NA)][NSYM]   M^T C_k M
  static constexpr int o_rk = o_bk + (1 + NA) * NSYM;                // [(1+NA)][D][D]   M^T C_k
  static constexpr int o_ck = o_rk + (1 + NA) * D * D;               // [(1+NA)][NSYM]   C_k
  static constexpr int o_kap = o_ck + (1 + NA) * NSYM;               // [(1+NA)][T][NPAIR]
  static constexpr int o_beta = o_kap + (1 + NA) * T * NPAIR;        // [NA][T][D+1][NRHS]
  static constexpr int o_K = ((o_beta + NA1 * T * (D + 1) * NRHS + 1) / 2) * 2;  // [NH][N] half stencil
  static constexpr int o_p = o_K + NH * N;                           // [NRHS][N] search directions
  static constexpr int o_atoms = o_p;                                // [NA][T][NRC]: dead once the stencil and
                                                                     // load vectors exist, so it shares p's storage
  static constexpr int total = o_p + ((AGLOB || NRHS * N > NA1 * T * NRC) ? NRHS * N : NA1 * T * NRC);
  static constexpr int scratch_doubles = AGLOB ? NA1 * T * NRC : 0;
};

template <class CO, int NM, int NT, int COLL = 0, int AGLOB = 0>
HMX_DEV void poisson_cell_body(const CellParams& P) {
  static_assert((COLL & CO::YDEP) == 0, "only axes the coefficient does not depend on can be collapsed");
  using L = PoissonLayout<CO, NM, NT, COLL, AGLOB>;
  using G = Grid<CO::DIM, NM, COLL>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP>;
  constexpr int D = L::D, T = L::T, N = L::N, NRHS = L::NRHS, NH = L::NH, NW = L::NW;
  constexpr int NA = L::NA, NA1 = L::NA1, NPAIR = L::NPAIR, NSYM = L::NSYM, NRC = L::NRC;
  constexpr int NPT = (N + NT - 1) / NT;
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;
  static_assert(NT % 32 == 0, "block size must be a multiple of the warp size");

  double* sm = dyn_smem();
  double* s_red = sm + L::o_red;
  double* s_vt = sm + L::o_vt;
  double* s_bk = sm + L::o_bk;
  double* s_rk = sm + L::o_rk;
  double* s_ck = sm + L::o_ck;
  double* s_kap = sm + L::o_kap;
  double* s_beta = sm + L::o_beta;
  double* s_atoms = AGLOB ? P.scratch + (size_t)bid() * L::scratch_doubles : sm + L::o_atoms;
  double* s_K = sm + L::o_K;
  double* s_p = sm + L::o_p;

  const int t_id = tid();
  const double h = 1.0 / (double)NM;
  // |e| times the number of identical layers a collapsed grid stands for
  const double vol = (D == 2 ? 0.5 : 1.0 / 6.0) * (D == 2 ? h * h : h * h * h) * (double)G::NLAYERS;
  int red_flip = 0;

  // Node ownership.  TILED: a thread owns NPT CONSECUTIVE nodes along the last axis (a piece of a grid line; a warp
  // holds 32 neighbouring lines, so every access is to consecutive doubles).  In K p a neighbouring line then serves
  // two stencil directions of all NPT nodes from NPT + 1 loads (instead of 2 NPT), the own line needs two halo loads:
  // 8 instead of 14 (3-D), 3 instead of 6 (2-D) loads of p per node, right-hand side and iteration -- the kernel is
  // bound by shared-memory bandwidth (74 % of the pipe, profiles/r02_c3_poisson_v2_raw.txt).  Otherwise thread t owns
  // nodes t, t + NT, ...
  constexpr int LAST = 1 << (D - 1);  // stencil mask bit of the last axis
  constexpr int NC = N / NM;          // nodes per layer of the last axis = number of grid lines (COLL = 0)
#ifndef HMX_POISSON_TILED
#define HMX_POISSON_TILED 1
#endif
  constexpr int NACT = NPT > 0 && NM % NPT == 0 ? NC * (NM / NPT) : NT + 1;  // threads that own a line piece
  constexpr bool TILED = HMX_POISSON_TILED && COLL == 0 && NPT >= 2 && NM % NPT == 0 && NACT <= NT;
  const bool own = !TILED || t_id < NACT;  // (NT is NACT rounded up to whole warps: the last threads own nothing)
  const int line = TILED ? (own ? t_id : 0) % NC : 0, z0 = TILED ? ((own ? t_id : 0) / NC) * NPT : 0;
  auto node = [&](int j) { return TILED ? (own ? line + NC * (z0 + j) : N) : t_id + j * NT; };
  int lnp[TILED ? LAST : 1], lnm[TILED ? LAST : 1], zoff[TILED ? NPT + 2 : 1];  // neighbour lines (layer 0), layers z0-1 .. z0+NPT
  if (TILED) {
    int c[3];
    G::decode(line, c);
    lnp[0] = lnm[0] = line;
    HMX_UNROLL
    for (int m = 1; m < LAST; ++m) {
      lnp[TILED ? m : 0] = G::template shifted<1>(c, m);
      lnm[TILED ? m : 0] = G::template shifted<-1>(c, m);
    }
    HMX_UNROLL
    for (int k = 0; k < NPT + 2; ++k) zoff[TILED ? k : 0] = NC * ((z0 + k - 1 + NM) % NM);
  }
  // neighbours of the owned nodes: they depend on the thread only, not on the macro point -- computing the
  // periodic wrap inside the PCG loop cost ~35 % of all issued instructions (profiles/r01_p2_inclusion16_raw.txt)
  constexpr bool NBREG = !TILED && NPT * 2 * NH <= 32;
  int nbp[NBREG ? NPT : 1][NBREG ? NH : 1], nbm[NBREG ? NPT : 1][NBREG ? NH : 1];
  if (NBREG) {
    HMX_UNROLL
    for (int j = 0; j < NPT; ++j) {
      const int i = node(j);
      int c[3];
      G::decode(i < N ? i : 0, c);
      HMX_UNROLL
      for (int s = 0; s < NH; ++s) {
        nbp[NBREG ? j : 0][NBREG ? s : 0] = G::template shifted<1>(c, s + 1);
        nbm[NBREG ? j : 0][NBREG ? s : 0] = G::template shifted<-1>(c, s + 1);
      }
    }
  }

  for (long long pt = bid(); pt < P.n_pts; pt += nblocks()) {
    // ---- 0. macro point, per-point constants, stratification Jacobian (registers) ----
    double xm[3];
    {  // the vertices are needed again in the epilogue only: parked in shared memory, not in 24 registers
      double verts[(D + 1) * 3];
      macro_point<D>(P, pt, xm, verts);
      if (t_id == NT - 1) {
        HMX_UNROLL
        for (int k = 0; k < (D + 1) * 3; ++k) s_vt[k] = verts[k];
      }
    }
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double M[D * D];  // M[p*D+i] = d theta_i / d x_p  (hmm.py:756-757)
    CO::dtheta(xm, M);

    // ---- 1. atoms: per-element quadrature means on the reduced cube set ----
    if (NA > 0) {
      for (int idx = t_id; idx < T * NRC; idx += NT) {
        const int t = idx / NRC, rc = idx - t * NRC;
        int c[3];
        AI::rdecode(rc, c);
        double acc[NA1];
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) acc[k] = 0.0;
        for (int q = 0; q < P.nq; ++q) {
          double y[D], s[NA1];
          HMX_UNROLL
          for (int a = 0; a < D; ++a) y[a] = ((double)c[a] + P.qp[(t * P.nq + q) * D + a]) * h;
          CO::atoms(pc, y, s);
          const double w = P.qw[q];
          HMX_UNROLL
          for (int k = 0; k < NA1; ++k) acc[k] += w * s[k];
        }
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) s_atoms[(k * T + t) * NRC + rc] = acc[k];
      }
    }
    // ---- 2a. C_k (A = C_0 + sum_k s_k C_k), B_k = M^T C_k M, R_k = M^T C_k ----
    if (t_id <= NA) {
      double Call[(1 + NA) * NSYM];
      CO::tensor_affine(pc, Call);
      double C[D][D];
      HMX_UNROLL
      for (int kk = 0; kk <= NA; ++kk)
        if (kk == t_id) {
          HMX_UNROLL
          for (int i = 0; i < D; ++i)
            HMX_UNROLL
            for (int j = 0; j < D; ++j) C[i][j] = Call[kk * NSYM + sym_index(D, i, j)];
        }
      double R[D][D];  // R[i][q] = sum_p M[p][i] C[p][q]
      HMX_UNROLL
      for (int i = 0; i < D; ++i)
        HMX_UNROLL
        for (int q = 0; q < D; ++q) {
          double s = 0.0;
          HMX_UNROLL
          for (int p = 0; p < D; ++p) s += M[p * D + i] * C[p][q];
          R[i][q] = s;
          s_rk[(t_id * D + i) * D + q] = s;
        }
      HMX_UNROLL
      for (int i = 0; i < D; ++i)
        HMX_UNROLL
        for (int j = i; j < D; ++j) {
          double s = 0.0;
          HMX_UNROLL
          for (int q = 0; q < D; ++q) s += R[i][q] * M[q * D + j];
          s_bk[t_id * NSYM + sym_index(D, i, j)] = s;
          s_ck[t_id * NSYM + sym_index(D, i, j)] = C[i][j];
        }
    }
    sync();
    // ---- 2b. element tables: kap (stiffness, pairs a<b) and beta (load vectors) ----
    {
      const double w = vol * (double)NM * (double)NM;
      for (int e = t_id; e < (1 + NA) * T * NPAIR; e += NT) {
        const int k = e / (T * NPAIR), t = (e / NPAIR) % T, pr = e % NPAIR;
        int a = 0, b = 1;
        for (int cnt = 0, aa = 0; aa < D; ++aa)
          for (int bb = aa + 1; bb <= D; ++bb, ++cnt)
            if (cnt == pr) {
              a = aa;
              b = bb;
            }
        // K_e[a][b] = w (B[a-1][b-1] - B[a-1][b] - B[a][b-1] + B[a][b]) in walk order
        double v = 0.0;
        for (int da = 0; da < 2; ++da)
          for (int db = 0; db < 2; ++db) {
            const int i = a - 1 + da, j = b - 1 + db;
            if (i < 0 || i >= D || j < 0 || j >= D) continue;
            const double bij = s_bk[k * NSYM + sym_index(D, kuhn_axis<D>(t, i), kuhn_axis<D>(t, j))];
            v += (da == db) ? bij : -bij;
          }
        s_kap[e] = w * v;
      }
      const double wb = -vol * (double)NM;
      for (int e = t_id; e < NA * T * (D + 1) * NRHS; e += NT) {
        const int q = e % NRHS, a = (e / NRHS) % (D + 1), t = (e / (NRHS * (D + 1))) % T, k = e / (NRHS * (D + 1) * T);
        double v = 0.0;
        if (a >= 1) v += s_rk[((1 + k) * D + kuhn_axis<D>(t, a - 1)) * D + q];
        if (a < D) v -= s_rk[((1 + k) * D + kuhn_axis<D>(t, a)) * D + q];
        s_beta[e] = wb * v;
      }
    }
    sync();

    // ---- 2c. half stencil and load vectors of the owned nodes ----
    // The tables kap / beta are the same for every node (structured mesh): each entry is loaded ONCE per thread and
    // used for (up to) JB owned nodes, whose accumulators live in registers together -- the per-node order of the
    // sums is (t, a, b) as in the reference assembly.  (One node at a time, this phase issued more shared-memory
    // loads than all PCG iterations of a C3 cell: 168 per node.)
    double bq[NPT][NRHS], kown[NPT];  // kown: sum of the node's own half-stencil entries (for the diagonal)
    constexpr int JB = NPT < 4 ? NPT : 4;
    HMX_UNROLL
    for (int j0 = 0; j0 < NPT; j0 += JB) {
      double acc[JB][NH];
      int cj[JB][3];
      bool valid[JB];
      HMX_UNROLL
      for (int jj = 0; jj < JB; ++jj) {
        const int j = j0 + jj < NPT ? j0 + jj : NPT - 1;
        const int i = node(j);
        valid[jj] = j0 + jj < NPT && i < N;
        G::decode(valid[jj] ? i : 0, cj[jj]);
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) bq[j][q] = 0.0;
        HMX_UNROLL
        for (int s = 0; s < NH; ++s) acc[jj][s] = 0.0;
      }
      HMX_UNROLL
      for (int t = 0; t < T; ++t) {
        HMX_UNROLL
        for (int a = 0; a <= D; ++a) {
          double be[NA1][NRHS], sa[JB][NA1];
          HMX_UNROLL
          for (int k = 0; k < NA; ++k)
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) be[k][q] = s_beta[((k * T + t) * (D + 1) + a) * NRHS + q];
          HMX_UNROLL
          for (int jj = 0; jj < JB; ++jj) {
            // node i is vertex a of the type-t element of cube o = c - P(t,a)
            int o[3];
            G::template shift_coords<-1>(cj[jj], kuhn_pmask<D>(t, a), o);
            const int ro = AI::ridx(o);
            HMX_UNROLL
            for (int k = 0; k < NA; ++k) sa[jj][k] = s_atoms[(k * T + t) * NRC + ro];
            if (j0 + jj < NPT) {
              HMX_UNROLL
              for (int k = 0; k < NA; ++k)
                HMX_UNROLL
                for (int q = 0; q < NRHS; ++q) bq[j0 + jj < NPT ? j0 + jj : 0][q] += be[k][q] * sa[jj][k];
            }
          }
          HMX_UNROLL
          for (int b = a + 1; b <= D; ++b) {
            // pair index of (a,b) in the a<b enumeration
            const int pr = a * D - a * (a - 1) / 2 + (b - a - 1);
            const int slot = (kuhn_pmask<D>(t, b) & ~kuhn_pmask<D>(t, a)) - 1;
            const double k0 = s_kap[(0 * T + t) * NPAIR + pr];
            double kk[NA1];
            HMX_UNROLL
            for (int k = 0; k < NA; ++k) kk[k] = s_kap[((1 + k) * T + t) * NPAIR + pr];
            HMX_UNROLL
            for (int jj = 0; jj < JB; ++jj) {
              double v = k0;
              HMX_UNROLL
              for (int k = 0; k < NA; ++k) v += kk[k] * sa[jj][k];
              acc[jj][slot] += v;
            }
          }
        }
      }
      HMX_UNROLL
      for (int jj = 0; jj < JB; ++jj)
        if (valid[jj]) {
          const int i = node(j0 + jj);
          double ks = 0.0;
          HMX_UNROLL
          for (int s = 0; s < NH; ++s) {
            s_K[s * N + i] = acc[jj][s];
            ks += acc[jj][s];
          }
          kown[j0 + jj < NPT ? j0 + jj : 0] = ks;
        } else if (j0 + jj < NPT) {
          kown[j0 + jj < NPT ? j0 + jj : 0] = 0.0;
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) bq[j0 + jj < NPT ? j0 + jj : 0][q] = 0.0;  // (a slot past the last node)
        }
    }
    // atom means for <A> (every thread gets them)
    double smean[NA1];
    {
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) smean[k] = 0.0;
      if (NA > 0) {
        for (int idx = t_id; idx < T * NRC; idx += NT) {
          HMX_UNROLL
          for (int k = 0; k < NA; ++k) smean[k] += s_atoms[k * T * NRC + idx];
        }
        block_sum<NA1, NW>(smean, s_red + (red_flip ^= 1) * NW * L::NRED);  // also publishes s_K
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) smean[k] *= 1.0 / (double)(T * NRC);
      } else {
        sync();
      }
    }
    // ---- 2d. diagonal from the zero row sums (constants are in the kernel of K) ----
    double dinv[NPT], kdiag[NPT];
    HMX_UNROLL
    for (int j = 0; j < NPT; ++j) {
      const int i = node(j);
      dinv[j] = kdiag[j] = 0.0;
      if (i < N) {
        int c[3];
        G::decode(i, c);
        double d = -kown[j];  // own entries from registers, the neighbours' entries towards this node from memory
        HMX_UNROLL
        for (int s = 0; s < NH; ++s) {
          // neighbour i - (s + 1): with line pieces, line lnm[mask without the last axis], layer z or z - 1
          const int im = TILED ? lnm[TILED ? ((s + 1) & (LAST - 1)) : 0] + zoff[TILED ? (((s + 1) & LAST) ? j : j + 1) : 0]
                               : G::template shifted<-1>(c, s + 1);
          d -= s_K[s * N + im];
        }
        kdiag[j] = d;
        dinv[j] = d != 0.0 ? 1.0 / d : 0.0;
      }
    }

    // ---- 3. PCG on all right-hand sides ----
    double xq[NPT][NRHS], rq[NPT][NRHS];
    double rz[NRHS], rz0[NRHS];
    bool active[NRHS];
    {
      double part[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) part[q] = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = node(j);
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          xq[j][q] = 0.0;
          rq[j][q] = bq[j][q];
          const double z = dinv[j] * rq[j][q];
          part[q] += rq[j][q] * z;
          if (i < N) s_p[q * N + i] = z;
        }
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NRED);  // also publishes p
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        rz[q] = rz0[q] = part[q];
        active[q] = part[q] > P.atol * P.atol;
      }
    }
    int it = 0, its[NRHS];
    HMX_UNROLL
    for (int q = 0; q < NRHS; ++q) its[q] = 0;
    bool any = false;
    HMX_UNROLL
    for (int q = 0; q < NRHS; ++q) any = any || active[q];
    while (any && it < P.max_it) {
      ++it;
      double Ap[NPT][NRHS], pown[NPT][NRHS], pAp[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) pAp[q] = 0.0;
      if (TILED) {
        const int i0 = line + zoff[TILED ? 1 : 0];  // first own node
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j)
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            pown[j][q] = s_p[q * N + i0 + NC * j];
            Ap[j][q] = kdiag[j] * pown[j][q];
          }
        {  // own line, direction of the last axis: K[s][i] couples i and i + e, so node j uses k[j + 1] up, k[j] down
          double kl[NPT + 1];
          HMX_UNROLL
          for (int k = 0; k <= NPT; ++k) kl[k] = s_K[(LAST - 1) * N + line + zoff[TILED ? k : 0]];
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            const double below = s_p[q * N + line + zoff[0]], above = s_p[q * N + line + zoff[TILED ? NPT + 1 : 0]];
            HMX_UNROLL
            for (int j = 0; j < NPT; ++j)
              Ap[j][q] += kl[j + 1] * (j + 1 < NPT ? pown[j + 1 < NPT ? j + 1 : 0][q] : above) + kl[j] * (j > 0 ? pown[j > 0 ? j - 1 : 0][q] : below);
          }
        }
        HMX_UNROLL
        for (int m = 1; m < LAST; ++m) {
          const int s0 = m - 1, s1 = (m | LAST) - 1;  // directions m (same layer) and m + e_last (next layer)
          {  // neighbours i + m, i + m + e_last: entries K[s][i] of the own nodes, line lnp[m], layers z .. z + 1
            const int ln = lnp[TILED ? m : 0];
            double k0[NPT], k1[NPT];
            HMX_UNROLL
            for (int j = 0; j < NPT; ++j) {
              k0[j] = s_K[s0 * N + i0 + NC * j];
              k1[j] = s_K[s1 * N + i0 + NC * j];
            }
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) {
              double v[NPT + 1];
              HMX_UNROLL
              for (int k = 0; k <= NPT; ++k) v[k] = s_p[q * N + ln + zoff[TILED ? k + 1 : 0]];
              HMX_UNROLL
              for (int j = 0; j < NPT; ++j) Ap[j][q] += k0[j] * v[j] + k1[j] * v[j + 1];
            }
          }
          {  // neighbours i - m, i - m - e_last: entries K[s][neighbour], line lnm[m], layers z .. z - 1
            const int ln = lnm[TILED ? m : 0];
            double k0[NPT], k1[NPT];
            HMX_UNROLL
            for (int j = 0; j < NPT; ++j) {
              k0[j] = s_K[s0 * N + ln + zoff[TILED ? j + 1 : 0]];
              k1[j] = s_K[s1 * N + ln + zoff[TILED ? j : 0]];
            }
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) {
              double v[NPT + 1];
              HMX_UNROLL
              for (int k = 0; k <= NPT; ++k) v[k] = s_p[q * N + ln + zoff[TILED ? k : 0]];
              HMX_UNROLL
              for (int j = 0; j < NPT; ++j) Ap[j][q] += k0[j] * v[j + 1] + k1[j] * v[j];
            }
          }
        }
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j)
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            if (NACT < NT && !own) pown[j][q] = Ap[j][q] = 0.0;  // (a thread without nodes read thread 0's line)
            pAp[q] += pown[j][q] * Ap[j][q];
          }
      } else {
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = node(j);
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) Ap[j][q] = pown[j][q] = 0.0;
        if (i < N) {
          int c[3];
          G::decode(i, c);
          const double kd = kdiag[j];
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            pown[j][q] = s_p[q * N + i];
            Ap[j][q] = kd * pown[j][q];
          }
          HMX_UNROLL
          for (int s = 0; s < NH; ++s) {
            const int ip = NBREG ? nbp[NBREG ? j : 0][NBREG ? s : 0] : G::template shifted<1>(c, s + 1);
            const int im = NBREG ? nbm[NBREG ? j : 0][NBREG ? s : 0] : G::template shifted<-1>(c, s + 1);
            const double kp = s_K[s * N + i], km = s_K[s * N + im];
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) Ap[j][q] += kp * s_p[q * N + ip] + km * s_p[q * N + im];
          }
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) pAp[q] += pown[j][q] * Ap[j][q];
        }
      }
      }
      block_sum<NRHS, NW>(pAp, s_red + (red_flip ^= 1) * NW * L::NRED);
      double alpha[NRHS], part[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        alpha[q] = (active[q] && pAp[q] > 0.0) ? fast_div(rz[q], pAp[q]) : 0.0;
        part[q] = 0.0;
      }
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j)
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) {
          xq[j][q] += alpha[q] * pown[j][q];
          rq[j][q] -= alpha[q] * Ap[j][q];
          part[q] += rq[j][q] * rq[j][q] * dinv[j];
        }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NRED);
      any = false;
      double beta[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        beta[q] = 0.0;
        if (active[q]) {
          beta[q] = fast_div(part[q], rz[q]);
          rz[q] = part[q];
          const double tol = fmax(P.rtol * P.rtol * rz0[q], P.atol * P.atol);
          if (!(part[q] > tol)) active[q] = false;
          its[q] = it;
        }
        any = any || active[q];
      }
      if (any) {
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j) {
          const int i = node(j);
          if (i < N) {
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q)
              if (active[q]) s_p[q * N + i] = dinv[j] * rq[j][q] + beta[q] * pown[j][q];
          }
        }
        sync();
      }
    }

    // ---- 4. epilogue: A_hom = <A> - b_p.x_q - x_p.r_q ----
    {
      double z[2 * NRHS * NRHS];
      HMX_UNROLL
      for (int k = 0; k < 2 * NRHS * NRHS; ++k) z[k] = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j)
        HMX_UNROLL
        for (int p = 0; p < NRHS; ++p)
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            z[p * NRHS + q] += bq[j][p] * xq[j][q];
            z[NRHS * NRHS + p * NRHS + q] += xq[j][p] * rq[j][q];
          }
      block_partials<2 * NRHS * NRHS, NW>(z, s_red + (red_flip ^= 1) * NW * L::NRED);  // warp partials only
      if (P.chi != nullptr) {  // correctors (BasePeriodicHMM.correctors, hmm.py:1211-1213), natural node order
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j) {
          const int i = node(j);
          if (i < N) {
            HMX_UNROLL
            for (int q = 0; q < NRHS; ++q) P.chi[((size_t)pt * NRHS + q) * N + i] = xq[j][q];
          }
        }
      }
      // one thread per entry of A_hom, another warp for the macro gradients meanwhile, then one thread per entry of
      // the macro element matrix (no serial chain on thread 0 while the CTA waits); the reduction buffer that the last
      // block_sum did NOT use is free
      constexpr int NBM = D + 1;
      double* s_ah = s_red + (red_flip ^ 1) * NW * L::NRED;  // [D][D]
      double* s_cm = s_ah + D * D;                           // [D][NBM] macro gradients, then |T|
      static_assert(D * D + D * NBM + 1 <= NW * L::NRED, "epilogue scratch");
      if (t_id < D * D) {
        const int p = t_id / D, q = t_id - p * D;
        double a = s_ck[sym_index(D, p, q)];
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) a += s_ck[(1 + k) * NSYM + sym_index(D, p, q)] * smean[k];
        // (z is indexed at run time: through the buffer the block_sum left the warp partials in)
        const double* zb = s_red + red_flip * NW * L::NRED;
        double z1 = 0.0, z2 = 0.0;
        for (int w = 0; w < NW; ++w) {
          z1 += zb[w * 2 * NRHS * NRHS + p * NRHS + q];
          z2 += zb[w * 2 * NRHS * NRHS + NRHS * NRHS + p * NRHS + q];
        }
        s_ah[t_id] = a - z1 - z2;
      }
      if (t_id == NT - 1) {  // (the last warp: the first holds the A_hom threads)
        if (P.S_loc != nullptr) {
          double Cm[D][NBM];
          s_cm[D * NBM] = macro_strain_matrix<D, 0>(s_vt, Cm);  // (written by this thread)
          HMX_UNROLL
          for (int p = 0; p < D; ++p)
            HMX_UNROLL
            for (int i = 0; i < NBM; ++i) s_cm[p * NBM + i] = Cm[p][i];
        }
        if (P.iters != nullptr) P.iters[pt] = it;
        if (P.work != nullptr) {
          unsigned long long tot = 0;
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) tot += (unsigned long long)its[q];
          atomic_add_u64(P.work, tot);
        }
        if (P.resid != nullptr) {
          double worst = 0.0;
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q)
            if (rz0[q] > P.atol * P.atol) worst = fmax(worst, sqrt(rz[q] / rz0[q]));
          P.resid[pt] = worst;
        }
      }
      sync();
      if (P.A_hom != nullptr && t_id < D * D) P.A_hom[pt * D * D + t_id] = s_ah[t_id];
      if (P.S_loc != nullptr && t_id < NBM * NBM) {
        const int i = t_id / NBM, j = t_id - i * NBM;
        double acc = 0.0;  // same summation order as macro_element_matrix
        HMX_UNROLL
        for (int p = 0; p < D; ++p)
          HMX_UNROLL
          for (int q2 = 0; q2 < D; ++q2) acc += s_cm[p * NBM + j] * s_ah[p * D + q2] * s_cm[q2 * NBM + i];
        P.S_loc[pt * NBM * NBM + t_id] = s_cm[D * NBM] * acc;
      }
    }
    sync();  // shared memory is reused by the next macro point
  }
}

}  // namespace hmx
