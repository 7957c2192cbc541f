// Two-level preconditioner of the matrix-free elasticity cell kernel (HMX_PRECOND = 1):
//   M^-1 = blockdiag(K)^-1  +  P E^-1 P^T,      E = P^T K P  (Galerkin, dense, factorised once per macro point).
//
// Why: the cell problem of BASELINE config 4 (stiff fibres mu = 100 in a nearly incompressible soft matrix,
// lambda / mu = 1000) needs ~245 block-Jacobi PCG iterations per right-hand side; the additive coarse correction
// brings that to ~120 (scripts/study_preconditioners.py, measured on the device in DESIGN.md) for ~6 % more work
// per iteration.  It replaces nothing of the reference's algorithm (the reference solves every corrector with
// GMRES + ILU, cell_problem.py:363-388): it only changes how fast the same discrete correctors are reached.
//
// Coarse space, two nested steps:
//   level 1: P1 on the Kuhn mesh of the (NM/2)^D grid.  The Kuhn triangulation of the fine grid refines it, so the
//            level-1 functions ARE fine-grid P1 functions: a fine node is a level-1 node (weight 1) or the midpoint
//            of a Kuhn edge (1/2, 1/2), and the Galerkin matrix E1 = P1^T K P1 is the ordinary P1 stiffness matrix
//            of the level-1 mesh with every coarse simplex carrying the MEAN of the coefficients (atoms) of the 2^D
//            fine simplices it contains -- it is assembled directly, no operator applications are needed;
//   level 2 (only when D * (NM/2)^D unknowns do not fit: NM = 8 in 3-D): along ONE axis SA the coefficient does not
//            depend on (CO::YDEP) the level-1 functions are summed up, i.e. the coarse functions are constant along
//            SA -- as the correctors are for such a coefficient -- E2 = P2^T E1 P2 (3-D, NM = 8: 1 x 4 x 4 nodes, 48
//            unknowns).  A coefficient that depends on every axis gets no such step: semi-coarsening along an axis the
//            coefficient varies on makes PCG SLOWER than block Jacobi (measured: 350 -> 570 iterations for a stiff
//            ball at 8^3, scripts/study_preconditioners.py), so its 8^3 cell keeps block Jacobi.
// E (+ a rank-D term gamma Z Z^T on the translations Z, so that the singular Galerkin matrix becomes definite
// without changing the solution for the consistent residuals PCG produces) is inverted explicitly (symmetric
// sweep operator, in place on the packed lower triangle); the apply is one dense symmetric matrix-vector product
// per right-hand side and iteration, FP64 (rounded to FP32 the coarse inverse loses a third of its effect).
//
// Layout trick: in the parity-major node layout (PGrid) the class-0 slots (all coordinates even) are the level-1
// grid in natural order, so restriction, coarse solution and prolongation all happen IN PLACE in the class-0
// slots of the right-hand side's y vector (which holds the residual while the operator is not being applied).
#pragma once
#include "hmx_cell_common.cuh"

#ifndef HMX_PRECOND
#define HMX_PRECOND 0  // 0 = block Jacobi, 1 = additive two-level where the coarse space fits
#endif

namespace hmx {

// Kuhn type of the level-1 simplex that contains the fine simplex of type t in sub-cube b (one bit per axis) of a
// level-1 cube: points of a type-t simplex have xi_pi(0) >= xi_pi(1) >= ..., so in the coarse cube (b + xi) / 2 the
// axes with b = 1 come first, each group in the order of pi (a stable partition of pi by b).
template <int D>
HMX_HOSTDEV constexpr int kuhn_parent_type(int t, int b) {
  int perm[3] = {0, 0, 0};
  int n = 0;
  for (int pass = 1; pass >= 0; --pass)
    for (int k = 0; k < D; ++k) {
      const int ax = kuhn_axis<D>(t, k);
      if (((b >> ax) & 1) == pass) perm[n++] = ax;
    }
  for (int s = 0; s < kuhn_ntypes<D>(); ++s) {
    bool ok = true;
    for (int k = 0; k < D; ++k) ok = ok && kuhn_axis<D>(s, k) == perm[k];
    if (ok) return s;
  }
  return -1;
}

// BASE: doubles of shared memory the kernel uses without the coarse space (the coarse arrays must fit on top).
template <class CO, int NM, int NT, int COLL, int VGLOB, int BASE = 0>
struct CoarseSpace {
  static constexpr int D = CO::DIM;
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int NRHS = D * (D + 1) / 2;
  static constexpr int TPR = NT / NRHS;
  static constexpr int NA1 = CO::NATOMS > 0 ? CO::NATOMS : 1;
  static constexpr int H = NM / 2;  // level-1 grid extent
  static constexpr int NC1 = ipow(H, D);
  static constexpr int MAXDOF = 96;  // 96 x 97 / 2 doubles = 37 KB
  static constexpr int SMEM_DOUBLES = 232448 / 8;  // 227 KB of dynamic shared memory per CTA
  static constexpr bool GEOM = HMX_PRECOND == 1 && COLL == 0 && VGLOB == 0 && NM % 2 == 0 && NM >= 4 && TPR >= 32;
  // first axis the coefficient does not depend on, -1 if it depends on all
  HMX_HOSTDEV static constexpr int invariant_axis() {
    for (int a = 0; a < D; ++a)
      if (!((CO::YDEP >> a) & 1)) return a;
    return -1;
  }
  static constexpr bool SEMI = GEOM && D * NC1 > MAXDOF && invariant_axis() >= 0;
  static constexpr int SA = invariant_axis() >= 0 ? invariant_axis() : 0;
  HMX_HOSTDEV static constexpr int m2(int a) { return a < D ? ((SEMI && a == SA) ? 1 : H) : 1; }
  HMX_HOSTDEV static constexpr int stride1(int a) { return ipow(H, a); }  // level-1 node c sits in class-0 slot sum c_a H^a
  static constexpr int NC2 = m2(0) * m2(1) * m2(2);
  static constexpr int NCD = D * NC2;  // coarse unknowns, index = node * D + component
  static constexpr int NTRI = NCD * (NCD + 1) / 2;
  static constexpr int NS = ipow(3, D);  // neighbour offsets {-1,0,1}^D of a level-1 row
  static constexpr int R1 = D * NC1;     // level-1 rows
  static constexpr int E1_DOUBLES = NS * D * R1;
  using CAI = AtomIdx<D, H, CO::YDEP, false>;  // level-1 cubes on the reduced (atom-dependent) axes, natural order
  static constexpr int NRC1 = CAI::NRC;
  static constexpr int setup_doubles = E1_DOUBLES + NA1 * T * NRC1;  // scratch of coarse_setup (the p / y area)
  // shared scratch: set-up = scale vector + two column buffers + two pivots of the inversion (rows / columns padded
  // to the 16 x (NT / 16) thread grid); solve = the restricted residual of every right-hand side
  HMX_HOSTDEV static constexpr int inv_pad() {
    const int ra = (NCD + 15) / 16 * 16, rb = (NCD + NT / 16 - 1) / (NT / 16) * (NT / 16);
    return ra > rb ? ra : rb;
  }
  static constexpr int CBUF = (3 * inv_pad() + 2 > NRHS * NCD ? 3 * inv_pad() + 2 : NRHS * NCD) + 2;
  static constexpr int NPAR = (ipow(2, D) * NC1 + 1) / 2;  // parent table (16-bit pairs), in doubles
  // on when the coarse unknowns fit, the set-up scratch fits in the p / y area, slots fit in 16 bits and the coarse
  // arrays fit in shared memory on top of everything else
  static constexpr bool ON = GEOM && NCD <= MAXDOF && ipow(2, D) * NC1 <= 65535 && setup_doubles <= 2 * NRHS * D * ipow(2, D) * NC1 &&
                             BASE + NTRI + NPAR + CBUF + 128 <= SMEM_DOUBLES;  // (+ the epilogue area of the kernel)
  // level-1 class-0 slot of level-2 node C (natural index on the m2 grid; coordinate 0 along a summed-up axis)
  HMX_DEV static int slot2(int C) {
    int s = 0;
    HMX_UNROLL
    for (int a = 0; a < D; ++a) {
      const int ca = C % m2(a);
      C /= m2(a);
      s += ca * stride1(a);
    }
    return s;
  }
};

// start of column j (minus j) of an n x n lower triangle packed by columns: entry (i, j), i >= j, sits at colbase + i
HMX_DEV constexpr int tri_colbase(int n, int j) { return j * n - (j * (j - 1)) / 2 - j; }

// parents of fine node slot i (parity-major): the level-1 nodes floor(c / 2) and floor(c / 2) + (c mod 2); the
// P1 interpolation is the mean of the two (they coincide for a level-1 node)
template <class CS, class PG>
HMX_DEV unsigned coarse_parents(int i) {
  constexpr int D = CS::D, H = CS::H;
  int c[3];
  PG::decode(i, c);
  unsigned ia = 0, ib = 0;
  HMX_UNROLL
  for (int a = 0; a < D; ++a) {
    const int lo = c[a] >> 1, hi = (lo + (c[a] & 1)) % H;
    ia += (unsigned)(lo * CS::stride1(a));
    ib += (unsigned)(hi * CS::stride1(a));
  }
  return ia | (ib << 16);
}

// Galerkin coarse matrix of this macro point, inverted: s_ei <- (E + gamma Z Z^T)^-1, packed lower triangle.
// Every thread of the CTA calls it; `work` (>= CS::setup_doubles doubles) is scratch, s_cbuf holds CS::NCD doubles.
template <class CS, class CO, int NM, int NT>
HMX_DEV void coarse_setup(const double* pc, const double (&Ms)[CO::DIM * CO::DIM], const double* s_atoms, double* work,
                          double* s_ei, double* s_cbuf, double* s_red, int red_stride, int& red_flip) {
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP, true>;
  using CAI = typename CS::CAI;
  using G1 = Grid<CO::DIM, CS::H, 0>;
  constexpr int D = CS::D, T = CS::T, NV = D * (D + 1) / 2, NA = CO::NATOMS, NA1 = CS::NA1, NRC = AI::NRC, NRC1 = CS::NRC1;
  constexpr int H = CS::H, R1 = CS::R1, NS = CS::NS, NCD = CS::NCD, NTRI = CS::NTRI, NW = NT / 32;
  double* s_e1 = work;                   // [NS][D][R1] level-1 matrix, one sparse row per level-1 unknown
  double* s_ca = work + CS::E1_DOUBLES;  // [NA][T][NRC1] coefficient means of the level-1 simplices
  const int t_id = tid();

  // ---- level-1 coefficients: mean over the 2^D fine simplices inside each level-1 simplex ----
  if (NA > 0) {
    for (int idx = t_id; idx < T * NRC1; idx += NT) {
      const int tc = idx / NRC1, rc = idx - tc * NRC1;
      int oc[3];
      CAI::rdecode(rc, oc);
      double acc[NA1];
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) acc[k] = 0.0;
      HMX_UNROLL
      for (int b = 0; b < (1 << D); ++b)
        HMX_UNROLL
        for (int t = 0; t < T; ++t)
          if (kuhn_parent_type<D>(t, b) == tc) {
            int of[3] = {0, 0, 0};
            HMX_UNROLL
            for (int a = 0; a < D; ++a) of[a] = 2 * oc[a] + ((b >> a) & 1);
            const int ro = AI::ridx(of);
            HMX_UNROLL
            for (int k = 0; k < NA; ++k) acc[k] += s_atoms[(k * T + t) * NRC + ro];
          }
      HMX_UNROLL
      for (int k = 0; k < NA; ++k) s_ca[(k * T + tc) * NRC1 + rc] = acc[k] * (1.0 / (double)(1 << D));
    }
  }
  sync();

  // ---- level-1 matrix rows: thread (node c, component ci) gathers the T (D + 1) level-1 simplices around c ----
  // (same element kernel as the fine grid with h -> 2 h:  sqrt(|e|) n M  scales by 2^(D/2) / 2)
  {
    const double cf = D == 3 ? 1.4142135623730951 : 1.0;
    for (int row = t_id; row < R1; row += NT) {
      const int c1 = row / D, ci = row - c1 * D;
      int c[3];
      G1::decode(c1, c);
      for (int e = 0; e < NS * D; ++e) s_e1[e * R1 + row] = 0.0;
#ifndef HMX_EMULATE
#pragma unroll 1
#endif
      for (int t = 0; t < T; ++t) {
        double g[D + 1][D];  // M-transformed gradients of the simplex's vertex functions
        HMX_UNROLL
        for (int a = 0; a <= D; ++a)
          HMX_UNROLL
          for (int p = 0; p < D; ++p) {
            g[a][p] = 0.0;
            if (a >= 1) g[a][p] += cf * Ms[p * D + kuhn_axis<D>(t, a >= 1 ? a - 1 : 0)];
            if (a < D) g[a][p] -= cf * Ms[p * D + kuhn_axis<D>(t, a < D ? a : 0)];
          }
        HMX_UNROLL
        for (int a = 0; a <= D; ++a) {  // the simplex of type t that has the node as its vertex a
          const int ma = kuhn_pmask<D>(t, a);
          int o[3];
          G1::template shift_coords<-1>(c, ma, o);
          const int ro = CAI::ridx(o);
          double sa[NA1], e[NV], sig[NV];
          HMX_UNROLL
          for (int k2 = 0; k2 < NA1; ++k2) sa[k2] = NA > 0 ? s_ca[(k2 * T + t) * NRC1 + ro] : 0.0;
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
            const int mb = kuhn_pmask<D>(t, b);
            int sl = 0, w3 = 1;
            HMX_UNROLL
            for (int ax = 0; ax < D; ++ax) {
              sl += (((mb >> ax) & 1) - ((ma >> ax) & 1) + 1) * w3;
              w3 *= 3;
            }
            HMX_UNROLL
            for (int j = 0; j < D; ++j) {
              double tv = 0.0;
              HMX_UNROLL
              for (int p = 0; p < D; ++p) tv = fma(S[j][p], g[b][p], tv);
              s_e1[(sl * D + j) * R1 + row] += tv;
            }
          }
        }
      }
    }
  }
  sync();

  // ---- coarse matrix E = P2^T E1 P2 (P2 = identity without semi-coarsening), packed lower triangle ----
  auto decode_tri = [&](int e, int& i, int& j) {
    const double b = (double)(2 * NCD + 1);
    j = (int)((b - sqrt(b * b - 8.0 * (double)e)) * 0.5);
    if (j < 0) j = 0;
    if (j > NCD - 1) j = NCD - 1;
    while (j + 1 < NCD && tri_colbase(NCD, j + 1) + (j + 1) <= e) ++j;
    while (tri_colbase(NCD, j) + j > e) --j;
    i = e - tri_colbase(NCD, j);
  };
  // entry ((c, ki), (c2, kj)) of E1: the offset slot(s) of row (c, ki) that land on node c2 (0.0 if c2 is no neighbour)
  auto e1_entry = [&](const int (&c)[3], int ki, const int (&c2)[3], int kj) {
    const int row = G1::index(c[0], c[1], c[2]) * D + ki;
    if (H >= 3) {  // the offsets -1, 0, +1 are distinct nodes: at most one slot
      int sl = 0, w3 = 1;
      HMX_UNROLL
      for (int a = 0; a < D; ++a) {
        const int dd = (c2[a] - c[a] + H) % H;
        if (dd > 1 && dd < H - 1) return 0.0;
        sl += (dd == 0 ? 1 : (dd == 1 ? 2 : 0)) * w3;
        w3 *= 3;
      }
      return s_e1[(sl * D + kj) * R1 + row];
    }
    double sum = 0.0;  // H = 2: +1 and -1 are the same node, their slots add up
    for (int o2 = (D == 3 ? -1 : 0); o2 <= (D == 3 ? 1 : 0); ++o2) {
      if (D == 3 && (c[2] + o2 + H) % H != c2[2]) continue;
      for (int o1 = -1; o1 <= 1; ++o1) {
        if ((c[1] + o1 + H) % H != c2[1]) continue;
        for (int o0 = -1; o0 <= 1; ++o0) {
          if ((c[0] + o0 + H) % H != c2[0]) continue;
          const int sl = (o0 + 1) + 3 * (o1 + 1) + (D == 3 ? 9 * (o2 + 1) : 0);
          sum += s_e1[(sl * D + kj) * R1 + row];
        }
      }
    }
    return sum;
  };
  constexpr int NSUP = CS::SEMI ? H : 1;  // level-1 nodes summed into one level-2 node (all H along SA), weight 1
  double diag[1] = {0.0};
  for (int e = t_id; e < NTRI; e += NT) {
    int i, j;
    decode_tri(e, i, j);
    const int Ci = i / D, ki = i - Ci * D, Cj = j / D, kj = j - Cj * D;
    int ci[3] = {0, 0, 0}, cj[3] = {0, 0, 0};
    {
      int r = Ci, s = Cj;
      HMX_UNROLL
      for (int a = 0; a < D; ++a) {
        ci[a] = r % CS::m2(a);
        r /= CS::m2(a);
        cj[a] = s % CS::m2(a);
        s /= CS::m2(a);
      }
    }
    double sum = 0.0;
    for (int si = 0; si < NSUP; ++si) {
      int c[3] = {ci[0], ci[1], ci[2]};
      if (NSUP > 1) c[CS::SA] = si;
      for (int sj = 0; sj < NSUP; ++sj) {
        int c2[3] = {cj[0], cj[1], cj[2]};
        if (NSUP > 1) c2[CS::SA] = sj;
        sum += e1_entry(c, ki, c2, kj);
      }
    }
    s_ei[e] = sum;
    if (i == j) diag[0] += sum;
  }
  block_sum<1, NW>(diag, s_red + (red_flip ^= 1) * red_stride);  // (its barrier also completes s_ei)
  const double gamma = diag[0] / ((double)NCD * (double)CS::NC2);  // same component: the translation Z_k has ones there

  // ---- inverse of E + gamma Z Z^T by the symmetric sweep operator, matrix in registers ----
  // Thread (ti, tj) of a 16 x (NT / 16) grid owns the entries (ti + 16 a, tj + TGC b) of the full square (both
  // triangles: every entry then sees the same update and stays bitwise symmetric).  Sweep k in the uniform form
  //   A <- A - w w^T / A_kk - 2 e_k e_k^T,   w = A[:, k] - e_k   (after n sweeps A = -E^-1)
  // (identical to the textbook sweep, but with no special case for row / column k); it cancels harmlessly once the
  // matrix is scaled to a unit diagonal (pivots <= 1).  One column broadcast and ONE barrier per sweep.
  {
    constexpr int TGR = 16, TGC = NT / TGR, RA = (NCD + TGR - 1) / TGR, RB = (NCD + TGC - 1) / TGC;
    constexpr int NPAD = (RA * TGR > RB * TGC ? RA * TGR : RB * TGC);
    static_assert(CS::CBUF >= 3 * NPAD + 2, "column buffers of the coarse inversion");
    double* s_scale = s_cbuf;          // [NPAD] 1 / sqrt(diagonal)
    double* s_w = s_cbuf + NPAD;       // 2 x [NPAD]
    double* s_piv = s_cbuf + 3 * NPAD;  // 2
    const int ti = t_id % TGR, tj = t_id / TGR;
    for (int i = t_id; i < NPAD; i += NT) {
      s_scale[i] = i < NCD ? fast_rsqrt(s_ei[tri_colbase(NCD, i) + i] + gamma) : 0.0;
      s_w[i] = 0.0;
      s_w[NPAD + i] = 0.0;
    }
    sync();
    double A[RA][RB];
    HMX_UNROLL
    for (int a = 0; a < RA; ++a)
      HMX_UNROLL
      for (int b = 0; b < RB; ++b) {
        const int i = ti + TGR * a, j = tj + TGC * b;
        double v = 0.0;
        if (i < NCD && j < NCD) {
          v = i >= j ? s_ei[tri_colbase(NCD, j) + i] : s_ei[tri_colbase(NCD, i) + j];
          if ((i - j) % D == 0) v += gamma;
          v *= s_scale[i] * s_scale[j];
        }
        A[a][b] = v;
      }
    if (tj == 0) {
      HMX_UNROLL
      for (int a = 0; a < RA; ++a) {
        const int i = ti + TGR * a;
        if (i < NCD) s_w[i] = A[a][0] - (i == 0 ? 1.0 : 0.0);
        if (i == 0) s_piv[0] = fast_rsqrt(A[a][0]);
      }
    }
    sync();
    HMX_UNROLL
    for (int kb = 0; kb < RB; ++kb) {
      for (int kt = 0; kt < TGC; ++kt) {
        const int k = TGC * kb + kt;
        if (k >= NCD) break;
        const int cur = k & 1, nxt = cur ^ 1;
        const double rs = s_piv[cur];  // 1 / sqrt(pivot): w w^T / pivot = (w rs)(w rs)^T stays bitwise symmetric
        double wi[RA], wj[RB];
        HMX_UNROLL
        for (int a = 0; a < RA; ++a) wi[a] = s_w[cur * NPAD + ti + TGR * a] * rs;
        HMX_UNROLL
        for (int b = 0; b < RB; ++b) wj[b] = s_w[cur * NPAD + tj + TGC * b] * rs;
        HMX_UNROLL
        for (int a = 0; a < RA; ++a)
          HMX_UNROLL
          for (int b = 0; b < RB; ++b) A[a][b] = fma(-wi[a], wj[b], A[a][b]);
        if (tj == kt) {
          HMX_UNROLL
          for (int a = 0; a < RA; ++a)
            if (ti + TGR * a == k) A[a][kb] -= 2.0;
        }
        const int k1 = k + 1;
        if (k1 < NCD) {
          const bool in_kb = kt < TGC - 1;  // column k1 lies in block kb, else in kb + 1
          if (tj == (in_kb ? kt + 1 : 0)) {
            HMX_UNROLL
            for (int a = 0; a < RA; ++a) {
              const int i = ti + TGR * a;
              const double v = in_kb ? A[a][kb] : A[a][kb + 1 < RB ? kb + 1 : kb];
              if (i < NCD) s_w[nxt * NPAD + i] = v - (i == k1 ? 1.0 : 0.0);
              if (i == k1) s_piv[nxt] = fast_rsqrt(v);
            }
          }
        }
        sync();
      }
    }
    // A = -(scaled E)^-1: undo the scaling, store the lower triangle
    HMX_UNROLL
    for (int a = 0; a < RA; ++a)
      HMX_UNROLL
      for (int b = 0; b < RB; ++b) {
        const int i = ti + TGR * a, j = tj + TGC * b;
        if (i < NCD && j <= i) s_ei[tri_colbase(NCD, j) + i] = -A[a][b] * (s_scale[i] * s_scale[j]);
      }
  }
  sync();
}

// Row i = 32 rb + lane of  E^-1 r  from the lower triangle packed by columns.  With rb a compile-time constant after
// unrolling, every column j is one of three cases known at compile time: left of the block (entry (i, j) in column
// j: base s_ei + i, immediate offset), right of it (entry (j, i) in column i: base s_ei + colbase(i), immediate j), or
// inside the 32 x 32 diagonal block (per-lane choice between the two) -- two instructions per matrix entry.
template <int NCD, bool PAIR>
HMX_DEV double coarse_row_block(const double* s_ei, const double* s_r, int rb, int lane) {
  const int r0 = 32 * rb;
  const int i = r0 + lane < NCD ? r0 + lane : NCD - 1;
  const double* colb = s_ei + i;
  const double* rowb = s_ei + tri_colbase(NCD, i);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  HMX_UNROLL
  for (int j = 0; j < NCD; j += 2) {
    double rj0, rj1 = 0.0;
    if (PAIR)
      ld_pair(s_r + j, rj0, rj1);
    else {
      rj0 = s_r[j];
      if (j + 1 < NCD) rj1 = s_r[j + 1];
    }
    HMX_UNROLL
    for (int jj = 0; jj < 2; ++jj) {
      const int c = j + jj;
      if (c < NCD) {
        double a;
        if (c < r0)
          a = colb[tri_colbase(NCD, c)];
        else if (c >= r0 + 32)
          a = rowb[c];
        else
          a = (c - r0 <= lane) ? colb[tri_colbase(NCD, c)] : rowb[c];
        acc[(c >> 1) & 1 ? 2 + jj : jj] = fma(a, jj ? rj1 : rj0, acc[(c >> 1) & 1 ? 2 + jj : jj]);
      }
    }
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// z_coarse = P E^-1 P^T r for right-hand side q, by its TPR threads (l = thread within the group).
// On entry y_q = [D][N] holds the residual r (all nodes); on exit the class-0 slots hold the level-1 coarse
// solution u1 (every other slot still holds r), so that z_i = Dinv_i r_i + (u1[parent A] + u1[parent B]) / 2.
// s_r: NCD doubles of this right-hand side for the restricted residual.
// OWN (thread l holds level-1 node l and its fine nodes 2 c + m): `splus` = r[2c] + 1/2 sum_m r[2c + m], summed by the
// caller from its own registers, so that only the 2^D - 1 neighbours 2c - m are gathered here.
template <class CS, int NP, bool PAIR, bool OWN>
HMX_DEV void coarse_correct(double* y_q, const double* s_ei, double* s_r, int q, int l, const double (&splus)[CS::D]) {
  constexpr int D = CS::D, H = CS::H, NC1 = CS::NC1, NC2 = CS::NC2, NCD = CS::NCD, TPR = CS::TPR, N = NP, HC = NC1;
  group_sync(1 + q, TPR);  // r of every node of this right-hand side is in y_q
  // level-1 restriction: r1[c] = r[2c] + 1/2 sum over the 2 (2^D - 1) Kuhn neighbours 2c +- m
  if (OWN) {
    int pa[3] = {0, 0, 0}, mi[3] = {0, 0, 0};
    {
      int r = l;
      HMX_UNROLL
      for (int a = 0; a < D; ++a) {
        const int ca = r % H;
        r /= H;
        pa[a] = ca * CS::stride1(a);
        mi[a] = ((ca + H - 1) % H) * CS::stride1(a);
      }
    }
    HMX_UNROLL
    for (int k = 0; k < D; ++k) {
      double acc = 0.0;
      HMX_UNROLL
      for (int m = 1; m < (1 << D); ++m) {
        int im = 0;
        HMX_UNROLL
        for (int a = 0; a < D; ++a) im += ((m >> a) & 1) ? mi[a] : pa[a];
        acc += y_q[k * N + m * HC + im];
      }
      const double v = splus[k] + 0.5 * acc;
      if (CS::SEMI)
        y_q[k * N + l] = v;
      else
        s_r[l * D + k] = v;
    }
  }
  for (int c1 = l; !OWN && c1 < NC1; c1 += TPR) {
    int pa[3] = {0, 0, 0}, mi[3] = {0, 0, 0};
    {
      int r = c1;
      HMX_UNROLL
      for (int a = 0; a < D; ++a) {
        const int ca = r % H;
        r /= H;
        pa[a] = ca * CS::stride1(a);
        mi[a] = ((ca + H - 1) % H) * CS::stride1(a);
      }
    }
    HMX_UNROLL
    for (int k = 0; k < D; ++k) {
      double acc = 0.0;
      HMX_UNROLL
      for (int m = 1; m < (1 << D); ++m) {
        int im = 0;
        HMX_UNROLL
        for (int a = 0; a < D; ++a) im += ((m >> a) & 1) ? mi[a] : pa[a];
        acc += y_q[k * N + m * HC + c1] + y_q[k * N + m * HC + im];
      }
      const double v = y_q[k * N + c1] + 0.5 * acc;
      if (CS::SEMI)
        y_q[k * N + c1] = v;  // (the slot of a level-1 node is read by its own task only)
      else
        s_r[c1 * D + k] = v;
    }
  }
  group_sync(1 + q, TPR);
  if (CS::SEMI) {  // level 2: sum over the level-1 nodes along SA
    for (int C = l; C < NC2; C += TPR) {
      const int s = CS::slot2(C);
      HMX_UNROLL
      for (int k = 0; k < D; ++k) {
        double acc = y_q[k * N + s];
        HMX_UNROLL
        for (int t = 1; t < H; ++t) acc += y_q[k * N + s + t * CS::stride1(CS::SA)];
        s_r[C * D + k] = acc;
      }
    }
    group_sync(1 + q, TPR);
  }
  // u = E^-1 r2: one row of the packed symmetric inverse per lane, 32 consecutive rows per warp pass
  {
    constexpr int NBLK = (NCD + 31) / 32, WPG = TPR / 32;
    const int w = l >> 5, lane = l & 31;
    HMX_UNROLL
    for (int rb = 0; rb < NBLK; ++rb)
      if (w == rb % WPG) {
        const double u = coarse_row_block<NCD, PAIR>(s_ei, s_r, rb, lane);
        const int i = 32 * rb + lane;
        if (i < NCD) y_q[(i % D) * N + CS::slot2(i / D)] = u;
      }
  }
  group_sync(1 + q, TPR);
  if (CS::SEMI) {  // the coarse functions are constant along SA: copy to the other level-1 nodes of the line
    for (int c1 = l; c1 < NC1; c1 += TPR) {
      const int cs = (c1 / CS::stride1(CS::SA)) % H;
      if (cs != 0) {
        HMX_UNROLL
        for (int k = 0; k < D; ++k) y_q[k * N + c1] = y_q[k * N + c1 - cs * CS::stride1(CS::SA)];
      }
    }
    group_sync(1 + q, TPR);
  }
}

}  // namespace hmx
