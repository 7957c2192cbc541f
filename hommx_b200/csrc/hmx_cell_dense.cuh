// K5 -- small elasticity cells solved DIRECTLY: dense FP64 Cholesky of the periodic micro stiffness
// matrix, one CTA per macro point (SURVEY.md 8a "K5", north star (2): "a batched dense FP64 Cholesky
// ... kept only where it beats CG for small cells").  variant = native.DENSE.
//
// Where it pays: cells with N_dof <= 192 whose PCG needs hundreds of iterations -- above all the
// axis-collapsed 8^3 fibre cell of BASELINE config 4 (8 x 8 x 1 nodes x 3 components = 192 unknowns,
// 6 right-hand sides, ~220 block-Jacobi PCG iterations = 41 MFLOP of operator applications per macro
// point; the factorisation is 2.4 MFLOP).
//
// Per macro point:
//   1. atoms (per-element quadrature means of the y-dependent scalars), as in the matrix-free kernel;
//   2. the stiffness matrix K (lower triangle, packed by columns, shared memory) and the d(d+1)/2 load
//      vectors b_q are assembled element by element: the cubes are taken colour by colour, one thread
//      per (cube, matrix column) or (cube, right-hand side), so every entry has one writer per colour
//      (no atomics, fixed order -> bitwise reproducible).  hmm.py:887-903 (forms), 397-432 (problem);
//   3. the constant null space (cell_problem.py:355-361) is removed by pinning the node-0 unknowns
//      (identity rows): the loads are orthogonal to the translations, so A_hom does not see the choice;
//   4. right-looking Cholesky with the matrix in REGISTERS: 16 x 16 threads own the entries
//      (i, j) = (16 a + ti, 16 b + tj) (block-cyclic, <= 78 doubles per thread); each step broadcasts one
//      column through shared memory (double-buffered -> ONE barrier per step); the forward substitution of
//      the load vectors rides along (thread i owns row i of all b_q);
//   5. A_hom[p][q] = <C>[p][q] - y_p . y_q  with  y_q = L^-1 b_q  -- the Schur complement of the bordered
//      matrix [[K, B], [B^T, <C>]] (hmm.py:1199-1245 restated): no back substitution at all;
//      the Gram matrix y^T y is accumulated step by step by 36 threads (no reduction).  Cholesky is
//      backward stable, so x^T dK x ~ eps x^T |K| x: A_hom carries no condition-number amplification;
//   6. correctors (P.chi, rare path): L is kept in shared memory and L^T x = y is solved column by column.
#pragma once
#include "hmx_cell_common.cuh"

#ifndef HMX_DENSE_LOOKAHEAD
#define HMX_DENSE_LOOKAHEAD 1  // 1: publish column k + 1 early in the trailing update of step k; 0: after it (round-1 order);
#endif                         // 2: the miscompiled form of 1 (repro only)

namespace hmx {

template <class CO, int NM, int NT, int COLL = 0>
struct DenseLayout {
  static constexpr int D = CO::DIM;
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int N = Grid<D, NM, COLL>::N;  // nodes (natural order)
  static constexpr int NRHS = D * (D + 1) / 2;
  static constexpr int NV = NRHS;
  static constexpr int NDOF = N * D;
  static constexpr int TG = 16;                      // threads own entries on a TG x TG cyclic grid
  static constexpr int R = (NDOF + TG - 1) / TG;     // register tile is R x R (lower half used)
  static constexpr int NPAD = R * TG;
  static constexpr int NW = NT / 32;
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  static constexpr int NRC = AtomIdx<D, NM, CO::YDEP, true>::NRC;
  static constexpr int NTRI = NDOF * (NDOF + 1) / 2;
  static constexpr int o_red = 0;                          // 2 x [NW][NA1]
  static constexpr int o_atoms = o_red + 2 * NW * NA1;     // [NA][T][NRC]
  static constexpr int o_col = o_atoms + NA1 * T * NRC;    // 2 x [NPAD] current / next column
  static constexpr int o_piv = o_col + 2 * NPAD;           // 2 x 1 / sqrt(pivot)
  static constexpr int o_yb = o_piv + 2;                   // 2 x [NRHS] row k of the loads
  static constexpr int o_b = o_yb + 2 * NRHS;              // [NRHS][NPAD] loads, later y = L^-1 b
  static constexpr int o_gram = o_b + NRHS * NPAD;         // [NRHS][NRHS]
  static constexpr int o_dsave = o_gram + NRHS * NRHS;     // [NPAD] 1 / L_kk (corrector path)
  static constexpr int o_epi = o_dsave + NPAD;             // A_hom [NRHS][NRHS], macro strain matrix [NV][(D+1) D], |T|
  static constexpr int o_tri = o_epi + NRHS * NRHS + NV * (D + 1) * D + 2;  // packed lower triangle, by columns
  static constexpr int total = o_tri + NTRI;
  static constexpr int scratch_doubles = 8;                // none needed
  static_assert(NT == TG * TG, "the dense kernel runs 256 threads");
  static_assert(R <= 12, "N_dof <= 192: the register tile holds R (R + 1) / 2 doubles per thread");
  static_assert(NDOF >= D + 1, "cell too small");
};

// packed position of entry (i, j), i >= j, of an n x n lower triangle stored by columns
HMX_DEV constexpr int tri_index(int n, int i, int j) { return j * n - (j * (j - 1)) / 2 + (i - j); }

template <class CO, int NM, int NT, int COLL = 0>
HMX_DEV void elasticity_dense_cell_body(const CellParams& P) {
  static_assert((COLL & CO::YDEP) == 0, "only axes the coefficient does not depend on can be collapsed");
  using L = DenseLayout<CO, NM, NT, COLL>;
  using G = Grid<CO::DIM, NM, COLL>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP, true>;
  constexpr int D = L::D, T = L::T, NRHS = L::NRHS, NV = L::NV, NDOF = L::NDOF, NPAD = L::NPAD, R = L::R, TG = L::TG;
  constexpr int NW = L::NW, NA = L::NA, NA1 = L::NA1, NRC = L::NRC;
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;
  constexpr int NC = 1 << D;
  constexpr int CM = COLL & (NC - 1);                     // collapsed corner bits
  constexpr int NCAN = 1 << (D - popcount3(CM));          // distinct (canonical) corners of a cube
#define HMX_NCOL(a_) (G::ext(a_) == 1 ? 1 : (G::ext(a_) % 2 == 0 ? 2 : 3))
#define HMX_HALF(a_) (G::ext(a_) / 2)
  constexpr int NCOLT = HMX_NCOL(0) * HMX_NCOL(1) * HMX_NCOL(2);

  double* sm = dyn_smem();
  double* s_red = sm + L::o_red;
  double* s_atoms = sm + L::o_atoms;
  double* s_col = sm + L::o_col;
  double* s_piv = sm + L::o_piv;
  double* s_yb = sm + L::o_yb;
  double* s_b = sm + L::o_b;
  double* s_gram = sm + L::o_gram;
  double* s_dsave = sm + L::o_dsave;
  double* s_tri = sm + L::o_tri;

  const int t_id = tid();
  const int ti = t_id % TG, tj = t_id / TG;
  const double h = 1.0 / (double)NM;
  const double vol = (D == 2 ? 0.5 * h * h : h * h * h / 6.0) * (double)G::NLAYERS;
  const double sqrtw = sqrt(vol);
  int red_flip = 0;

  for (long long pt = bid(); pt < P.n_pts; pt += nblocks()) {
    double xm[3], verts[(D + 1) * 3];
    macro_point<D>(P, pt, xm, verts);
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double Ms[D * D];  // sqrt(|e|) n M;  M[p*D+i] = d theta_i / d x_p  (hmm.py:1015-1016)
    CO::dtheta(xm, Ms);
    HMX_UNROLL
    for (int k = 0; k < D * D; ++k) Ms[k] *= (double)NM * sqrtw;

    // ---- 1. atoms; clear K and b ----
    if (NA > 0) {
      for (int idx = t_id; idx < T * NRC; idx += NT) {
        const int t = idx / NRC, rc = idx - t * NRC;
        int c[3];
        const bool real_slot = AI::rdecode(rc, c);
        double acc[NA1];
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) acc[k] = 0.0;
        for (int qq = 0; real_slot && qq < P.nq; ++qq) {
          double y[D], s[NA1];
          HMX_UNROLL
          for (int a = 0; a < D; ++a) y[a] = ((double)c[a] + P.qp[(t * P.nq + qq) * D + a]) * h;
          CO::atoms(pc, y, s);
          const double wq = P.qw[qq];
          HMX_UNROLL
          for (int k = 0; k < NA1; ++k) acc[k] += wq * s[k];
        }
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) s_atoms[(k * T + t) * NRC + rc] = acc[k];
      }
    }
    for (int i = t_id; i < L::NTRI; i += NT) s_tri[i] = 0.0;
    for (int i = t_id; i < NRHS * NPAD; i += NT) s_b[i] = 0.0;
    sync();

    // mean atoms -> <C> (the first term of A_hom)
    double smean[NA1];
    HMX_UNROLL
    for (int k = 0; k < NA1; ++k) smean[k] = 0.0;
    if (NA > 0) {
      for (int idx = t_id; idx < T * NRC; idx += NT) {
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) smean[k] += s_atoms[k * T * NRC + idx];
      }
      block_sum<NA1, NW>(smean, s_red + (red_flip ^= 1) * NW * NA1);
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) smean[k] *= 1.0 / (double)(T * ipow(NM, AI::NDEP));
    }

    // ---- 2. assembly: thread r = (node v, component ci) builds row r of K and entry r of every load ----
    // (gather over the T (D + 1) simplices around the node: every entry has exactly one writer, no phases)
    if (t_id < NDOF) {
      const int v = t_id / D, ci = t_id - v * D;
      int c[3];
      G::decode(v, c);
      constexpr int NSLOT = D == 2 ? 9 : 27;  // neighbour offsets {-1,0,1}^D; only the ones a simplex spans are touched
      double acc[NSLOT][D];
      HMX_UNROLL
      for (int sl = 0; sl < NSLOT; ++sl)
        HMX_UNROLL
        for (int j = 0; j < D; ++j) acc[sl][j] = 0.0;
      double bl[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) bl[q] = 0.0;
      HMX_UNROLL
      for (int t = 0; t < T; ++t) {
        // gradients of the simplex's vertex functions (steps along collapsed axes join identified nodes: left out)
        double g[D + 1][D];
        HMX_UNROLL
        for (int a = 0; a <= D; ++a)
          HMX_UNROLL
          for (int p = 0; p < D; ++p) {
            g[a][p] = 0.0;
            if (a >= 1 && !((CM >> kuhn_axis<D>(t, a >= 1 ? a - 1 : 0)) & 1)) g[a][p] += Ms[p * D + kuhn_axis<D>(t, a >= 1 ? a - 1 : 0)];
            if (a < D && !((CM >> kuhn_axis<D>(t, a < D ? a : 0)) & 1)) g[a][p] -= Ms[p * D + kuhn_axis<D>(t, a < D ? a : 0)];
          }
        HMX_UNROLL
        for (int a = 0; a <= D; ++a) {  // the simplex of type t that has the node as its vertex a
          int o[3];
          G::template shift_coords<-1>(c, kuhn_pmask<D>(t, a), o);
          const int ro = AI::ridx(o);
          double sa[NA1], e[NV], sig[NV];
          HMX_UNROLL
          for (int k2 = 0; k2 < NA1; ++k2) sa[k2] = NA > 0 ? s_atoms[(k2 * T + t) * NRC + ro] : 0.0;
          // engineering-Voigt strain of (vertex function a) e_ci
          HMX_UNROLL
          for (int vv = 0; vv < D; ++vv) e[vv] = (vv == ci) ? g[a][vv] : 0.0;
          {
            int vv = D;
            HMX_UNROLL
            for (int r = 0; r < D; ++r)
              HMX_UNROLL
              for (int c2 = r + 1; c2 < D; ++c2) {
                e[vv] = ((c2 == ci) ? g[a][r] : 0.0) + ((r == ci) ? g[a][c2] : 0.0);
                ++vv;
              }
          }
          CO::stress(pc, sa, e, sig);
          // loads: b_q[r] = -sqrt|e| E_q : sigma(row)  (the tensor is symmetric -- so is everything Cholesky needs)
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) bl[q] = fma(-sqrtw, sig[q], bl[q]);
          double S[D][D];
          HMX_UNROLL
          for (int vv = 0; vv < D; ++vv) S[vv][vv] = sig[vv];
          {
            int vv = D;
            HMX_UNROLL
            for (int r = 0; r < D; ++r)
              HMX_UNROLL
              for (int c2 = r + 1; c2 < D; ++c2) {
                S[r][c2] = S[c2][r] = sig[vv];
                ++vv;
              }
          }
          HMX_UNROLL
          for (int b = 0; b <= D; ++b) {
            // neighbour offset pmask(b) - pmask(a) per axis, collapsed axes dropped
            int sl = 0, w3 = 1;
            HMX_UNROLL
            for (int ax = 0; ax < D; ++ax) {
              const int off = ((CM >> ax) & 1) ? 0 : ((kuhn_pmask<D>(t, b) >> ax) & 1) - ((kuhn_pmask<D>(t, a) >> ax) & 1);
              sl += (off + 1) * w3;
              w3 *= 3;
            }
            HMX_UNROLL
            for (int j = 0; j < D; ++j) {
              double tv = 0.0;
              HMX_UNROLL
              for (int p = 0; p < D; ++p) tv = fma(S[j][p], g[b][p], tv);
              acc[sl][j] += tv;
            }
          }
        }
      }
      HMX_UNROLL
      for (int sl = 0; sl < NSLOT; ++sl) {
        int cw[3] = {c[0], c[1], c[2]};
        {
          int r = sl;
          HMX_UNROLL
          for (int ax = 0; ax < D; ++ax) {
            const int off = r % 3 - 1;
            r /= 3;
            cw[ax] = off > 0 ? G::up(c[ax], ax) : (off < 0 ? G::down(c[ax], ax) : c[ax]);
          }
        }
        const int wn = G::index(cw[0], cw[1], cw[2]) * D;
        HMX_UNROLL
        for (int j = 0; j < D; ++j)
          if (wn + j <= t_id) s_tri[tri_index(NDOF, t_id, wn + j)] += acc[sl][j];  // (tiny extents: offsets may coincide)
      }
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) s_b[q * NPAD + t_id] = bl[q];
    }
    sync();

    // ---- 3./4. pinned null space, matrix and loads into registers ----
    double A[R][R];  // A[a][b], b <= a: entry (TG a + ti, TG b + tj)
    HMX_UNROLL
    for (int a = 0; a < R; ++a)
      HMX_UNROLL
      for (int b = 0; b <= a; ++b) {
        const int i = TG * a + ti, j = TG * b + tj;
        double v = 0.0;
        if (i < NDOF && i >= j) v = (j < D) ? (i == j ? 1.0 : 0.0) : s_tri[tri_index(NDOF, i, j)];
        A[a][b] = v;
      }
    double brow[NRHS];  // row t_id of the loads
    HMX_UNROLL
    for (int p = 0; p < NRHS; ++p) brow[p] = (t_id < NDOF && t_id >= D) ? s_b[p * NPAD + t_id] : 0.0;
    double gram = 0.0;  // thread p * NRHS + q accumulates y_p . y_q
#ifdef HMX_DENSE_DEBUG
    int dbg_first = -1;  // first step whose pivot is not finite
#endif
    // column 0 / pivot 0 / load row 0 for the first step
    if (tj == 0) {
      HMX_UNROLL
      for (int a = 0; a < R; ++a) s_col[TG * a + ti] = A[a][0];
      if (ti == 0) s_piv[0] = fast_rsqrt(A[0][0]);
    }
    if (t_id == 0) {
      HMX_UNROLL
      for (int p = 0; p < NRHS; ++p) s_yb[p] = brow[p];
    }
    sync();
    HMX_UNROLL
    for (int kb = 0; kb < R; ++kb) {
      for (int kt = 0; kt < TG; ++kt) {
        const int k = TG * kb + kt;
        if (k >= NDOF) break;
        const int cur = k & 1, nxt = cur ^ 1;
        const double* colk = s_col + cur * NPAD;
        const double rinv = s_piv[cur], rinv2 = rinv * rinv;  // 1 / L_kk, 1 / pivot
#ifdef HMX_DENSE_DEBUG
        if (t_id == 0 && dbg_first < 0 && !(rinv == rinv && rinv < 1e300)) dbg_first = k;
#endif
        double lj[R], li[R];  // L_jk L_kk (raw column) and L_ik / L_kk: their product is L_ik L_jk
        HMX_UNROLL
        for (int b = kb; b < R; ++b) lj[b] = colk[TG * b + tj];
        HMX_UNROLL
        for (int a = kb; a < R; ++a) li[a] = colk[TG * a + ti] * rinv2;
        const int k1 = k + 1;
        const bool in_kb = kt < TG - 1;  // column k1 lies in block kb, else in kb + 1
#if HMX_DENSE_LOOKAHEAD == 1
        // LOOK-AHEAD: the trailing update starts with the two block columns that can hold column k + 1 (b = kb and
        // b = kb + 1: 2 R of the R (R + 1) / 2 FMA), then the owners of column k + 1 (tj == k1 mod TG) hand it, its pivot
        // and the load row on to the next step, and the rest of the update (20k FMA per step at full size) runs while
        // that hand-off (shared-memory stores, the rsqrt chain) is in flight: the dependent chain of a step is
        // barrier -> column load -> 2 R FMA -> store + rsqrt  instead of the whole trailing update.
        // Writing the `nxt` buffers this early is safe: they were `cur` of step k - 1, and every thread has passed the
        // barrier that ended that step.  Control flow stays uniform (only the stores are predicated): a first version
        // that updated column k + 1 inside `if (owner)` returned NaN pivots on the device at ptxas -O1 and above, finite
        // ones at -O0 and in the CPU emulation (round-1 DESIGN.md 9.2, reproduced in round 2 with scripts/probe_dense.py).
        HMX_UNROLL
        for (int a = kb; a < R; ++a) {
          if (tj > kt) A[a][kb] = fma(-li[a], lj[kb], A[a][kb]);
          if (kb + 1 < R && a >= kb + 1) A[a][kb + 1 < R ? kb + 1 : kb] = fma(-li[a], lj[kb + 1 < R ? kb + 1 : kb], A[a][kb + 1 < R ? kb + 1 : kb]);
        }
        if (k1 < NDOF && tj == (k1 % TG)) {
          double* coln = s_col + nxt * NPAD;
          HMX_UNROLL
          for (int a = kb; a < R; ++a) {
            const double v = in_kb ? A[a][kb] : ((a >= kb + 1 && kb + 1 < R) ? A[a][kb + 1 < R ? kb + 1 : kb] : 0.0);
            coln[TG * a + ti] = v;  // (entries above the diagonal are never used)
          }
          if (ti == tj) s_piv[nxt] = fast_rsqrt(in_kb ? A[kb][kb] : A[kb + 1 < R ? kb + 1 : kb][kb + 1 < R ? kb + 1 : kb]);
        }
        HMX_UNROLL
        for (int a = kb + 2; a < R; ++a)
          HMX_UNROLL
          for (int b = kb + 2; b <= a; ++b) A[a][b] = fma(-li[a], lj[b], A[a][b]);
#elif HMX_DENSE_LOOKAHEAD == 2
        // KNOWN-BAD variant, kept only as the repro of the round-1 NaN (scripts/repro_dense_lookahead.py): the same
        // look-ahead with column k + 1 updated INSIDE the owners' branch.  Every matrix entry receives exactly the same
        // sequence of FMAs as in the variants above, the CPU emulation and `-Xptxas -O0` give finite, oracle-exact
        // results -- at ptxas -O1 and above the device returns NaN pivots from a data-dependent, run-to-run
        // reproducible step on (192 unknowns; 48- and 72-unknown cells are fine).  Not a race: an extra CTA barrier
        // after the branch changes nothing.
        const bool owner1 = k1 < NDOF && tj == (k1 % TG);
        if (owner1) {
          double* coln = s_col + nxt * NPAD;
          if (in_kb) {
            HMX_UNROLL
            for (int a = kb; a < R; ++a) {
              A[a][kb] = fma(-li[a], lj[kb], A[a][kb]);
              coln[TG * a + ti] = A[a][kb];
            }
            if (ti == tj) s_piv[nxt] = fast_rsqrt(A[kb][kb]);
          } else if (kb + 1 < R) {
            HMX_UNROLL
            for (int a = kb + 1; a < R; ++a) {
              A[a][kb + 1 < R ? kb + 1 : kb] = fma(-li[a], lj[kb + 1 < R ? kb + 1 : kb], A[a][kb + 1 < R ? kb + 1 : kb]);
              coln[TG * a + ti] = A[a][kb + 1 < R ? kb + 1 : kb];
            }
            coln[TG * kb + ti] = 0.0;
            if (ti == tj) s_piv[nxt] = fast_rsqrt(A[kb + 1 < R ? kb + 1 : kb][kb + 1 < R ? kb + 1 : kb]);
          }
        }
        HMX_UNROLL
        for (int a = kb; a < R; ++a)
          HMX_UNROLL
          for (int b = kb; b <= a; ++b) {
            if (b == kb) {
              if (tj > kt && !(in_kb && owner1)) A[a][b] = fma(-li[a], lj[b], A[a][b]);
            } else if (b == kb + 1) {
              if (in_kb || !owner1) A[a][b] = fma(-li[a], lj[b], A[a][b]);
            } else {
              A[a][b] = fma(-li[a], lj[b], A[a][b]);
            }
          }
#else
        // trailing update of the owned entries (i >= j > k)
        HMX_UNROLL
        for (int a = kb; a < R; ++a)
          HMX_UNROLL
          for (int b = kb; b <= a; ++b) {
            if (b == kb) {
              if (tj > kt) A[a][b] = fma(-li[a], lj[b], A[a][b]);
            } else {
              A[a][b] = fma(-li[a], lj[b], A[a][b]);
            }
          }
#endif
        // forward substitution of the loads: y_p[k] = b_p[k] / L_kk, b_p[i] -= L_ik y_p[k]
        if (t_id > k && t_id < NDOF) {
          const double lrow = colk[t_id] * rinv2;
          HMX_UNROLL
          for (int p = 0; p < NRHS; ++p) brow[p] = fma(-lrow, s_yb[cur * NRHS + p], brow[p]);
        }
        if (t_id < NRHS * NRHS) {
          gram = fma(s_yb[cur * NRHS + t_id / NRHS] * rinv2, s_yb[cur * NRHS + t_id % NRHS], gram);
        }
        if (P.chi != nullptr) {  // keep L and y for the back substitution
          if (t_id >= k && t_id < NDOF) s_tri[tri_index(NDOF, t_id, k)] = colk[t_id] * rinv;
          if (t_id < NRHS) s_b[t_id * NPAD + k] = s_yb[cur * NRHS + t_id] * rinv;
          if (t_id == 0) s_dsave[k] = rinv;
        }
        // hand the next column, pivot and load row to the next step
        if (k1 < NDOF) {
#if HMX_DENSE_LOOKAHEAD == 0
          if (tj == (k1 % TG)) {
            double* coln = s_col + nxt * NPAD;
            HMX_UNROLL
            for (int a = kb; a < R; ++a) {
              const double v = in_kb ? A[a][kb] : ((a >= kb + 1 && kb + 1 < R) ? A[a][kb + 1 < R ? kb + 1 : kb] : 0.0);
              coln[TG * a + ti] = v;  // (entries above the diagonal are never used)
            }
            if (ti == tj) s_piv[nxt] = fast_rsqrt(in_kb ? A[kb][kb] : A[kb + 1 < R ? kb + 1 : kb][kb + 1 < R ? kb + 1 : kb]);
          }
#endif
          if (t_id == k1) {
            HMX_UNROLL
            for (int p = 0; p < NRHS; ++p) s_yb[nxt * NRHS + p] = brow[p];
          }
        }
        sync();
      }
    }
    if (t_id < NRHS * NRHS) s_gram[t_id] = gram;
    sync();

    // ---- 6. correctors (rare path): L^T x = y, one column per step ----
    if (P.chi != nullptr) {
      // s_b holds y; thread j keeps x_q[j] of all right-hand sides
      double xr[NRHS];
      HMX_UNROLL
      for (int p = 0; p < NRHS; ++p) xr[p] = t_id < NDOF ? s_b[p * NPAD + t_id] : 0.0;
      for (int k = NDOF - 1; k >= 0; --k) {
        if (t_id == k) {
          HMX_UNROLL
          for (int p = 0; p < NRHS; ++p) {
            xr[p] *= s_dsave[k];
            s_yb[p] = xr[p];
          }
        }
        sync();
        if (t_id < k) {
          const double lkj = s_tri[tri_index(NDOF, k, t_id)];
          HMX_UNROLL
          for (int p = 0; p < NRHS; ++p) xr[p] = fma(-lkj, s_yb[p], xr[p]);
        }
        sync();
      }
      // natural node order [q][component][node]; shift to zero mean per component (the pinned solution differs
      // from it by a translation)
      if (t_id < NDOF) {
        HMX_UNROLL
        for (int p = 0; p < NRHS; ++p) s_b[p * NPAD + t_id] = xr[p];
      }
      sync();
      for (int idx = t_id; idx < NRHS * D; idx += NT) {
        const int p = idx / D, c = idx - p * D;
        double m = 0.0;
        for (int nd = 0; nd < L::N; ++nd) m += s_b[p * NPAD + nd * D + c];
        m /= (double)L::N;
        for (int nd = 0; nd < L::N; ++nd)
          P.chi[((size_t)pt * NRHS * D + idx) * L::N + nd] = s_b[p * NPAD + nd * D + c] - m;
      }
    }

    // ---- 5. A_hom = <C> - y^T y, local macro matrix (one thread per entry: with one CTA per SM nothing else
    //         would hide a serial epilogue) ----
    constexpr int NB = (D + 1) * D;
    double* s_ah = sm + L::o_epi;            // [NRHS][NRHS]
    double* s_cm = s_ah + NRHS * NRHS;       // [NV][NB] macro strain matrix, then |T|
    if (t_id == 0) {
      for (int qq = 0; qq < NRHS; ++qq) {
        double e[NV], sg[NV];
        HMX_UNROLL
        for (int v = 0; v < NV; ++v) e[v] = (v == qq) ? 1.0 : 0.0;
        CO::stress(pc, smean, e, sg);
        for (int p = 0; p < NRHS; ++p) s_ah[p * NRHS + qq] = sg[p] - s_gram[p * NRHS + qq];
      }
      if (P.S_loc != nullptr) {
        double Cm[NV][NB];
        s_cm[NV * NB] = macro_strain_matrix<D, 1>(verts, Cm);
        HMX_UNROLL
        for (int p = 0; p < NV; ++p)
          HMX_UNROLL
          for (int i = 0; i < NB; ++i) s_cm[p * NB + i] = Cm[p][i];
      }
#ifdef HMX_DENSE_DEBUG
      if (P.iters != nullptr) P.iters[pt] = dbg_first;
#else
      if (P.iters != nullptr) P.iters[pt] = 0;  // direct solve
#endif
      if (P.resid != nullptr) P.resid[pt] = 0.0;
    }
    sync();
    if (P.A_hom != nullptr && t_id < NRHS * NRHS) P.A_hom[pt * NRHS * NRHS + t_id] = s_ah[t_id];
    if (P.S_loc != nullptr && t_id < NB * NB) {
      const int i = t_id / NB, j = t_id - i * NB;
      double acc = 0.0;  // same summation order as macro_element_matrix
      HMX_UNROLL
      for (int p = 0; p < NV; ++p)
        HMX_UNROLL
        for (int q2 = 0; q2 < NV; ++q2) acc += s_cm[p * NB + j] * s_ah[p * NRHS + q2] * s_cm[q2 * NB + i];
      P.S_loc[pt * NB * NB + t_id] = s_cm[NV * NB] * acc;
    }
    sync();  // shared memory is reused by the next macro point
  }
#undef HMX_NCOL
#undef HMX_HALF
}

}  // namespace hmx
