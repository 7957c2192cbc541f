// Shared pieces of the micro cell kernels: launch parameters, the Kuhn/Freudenthal
// element tables of the structured periodic micro mesh, index helpers, block
// reductions and the macro element epilogue.
//
// Geometry restated: the micro mesh is the unit box split into
// NM^D cubes, each cut into D! simplices that all walk from the cube's origin corner
// to the opposite corner along one axis at a time ("type" t = the order pi_t in which
// the axes are walked).  Vertex a of a type-t simplex sits at offset
//   P(t,a) = e_{pi(0)} + .. + e_{pi(a-1)}            (a = 0..D)
// and the P1 gradients are  g_a = n (e_{pi(a-1)} - e_{pi(a)})  (terms with an index
// outside 0..D-1 dropped), so d u / d y_{pi(k)} = n (u_{P(k+1)} - u_{P(k)}).
// Periodicity (cell_problem.py:38-300 of the reference: every max-face node is a slave
// of its wrapped image) is the index wrap i mod NM.
#pragma once
#include "hmx_platform.cuh"

namespace hmx {

// General periodic micro mesh for the element-list kernel (hmx_cell_generic.cuh): any simplicial mesh
// of the unit box whose boundary nodes match periodically (cell_problem.py:16-35 of the reference accepts any such
// mesh).  Built on the host (hommx_b200/micro.py: element_list_tables); all pointers are device pointers.
struct MicroMesh {
  int n_elem, n_nodes, nnzb, nq;  // elements, PERIODIC nodes, blocks of the node-node pattern, quadrature points per element
  const int* elem_nodes;    // [n_elem][D+1] periodic node ids
  const double* elem_grad;  // [n_elem][D+1][D] gradients of the P1 basis functions
  const double* elem_vol;   // [n_elem]
  const double* elem_yq;    // [n_elem][nq][D] quadrature points (micro coordinates)
  const int* row_ptr;       // [n_nodes + 1] block rows of the periodic stiffness pattern
  const int* col;           // [nnzb]
  const int* blk_ptr;       // [nnzb + 1] contributions of every block, in a fixed order
  const int* blk_src;       //   (e * (D+1) + a) * (D+1) + b: element e couples its local vertices a (row) and b (column)
  const int* node_ptr;      // [n_nodes + 1] elements around every node
  const int* node_src;      //   e * (D+1) + a
  const int* diag;          // [n_nodes] block index of (i, i)
};
// doubles of per-CTA global scratch of the element-list kernel
HMX_HOSTDEV constexpr long long element_list_scratch(int n_elem, int n_nodes, int nnzb, int natoms1, int bs, int nrhs) {
  return (long long)natoms1 * n_elem + (long long)nnzb * bs * bs + (long long)n_nodes * bs * bs +
         5LL * nrhs * n_nodes * bs;  // atoms, matrix blocks, inverse diagonal blocks, b x r p y
}

struct CellParams {
  long long n_pts;         // macro quadrature points (= macro cells in fused mode)
  const double* x_pts;     // [n_pts][3] macro points, or nullptr when cells are given
  const int* cell_nodes;   // [n_pts][D+1] macro cell vertex ids, or nullptr
  const double* node_xyz;  // [n_nodes][3] macro vertex coordinates
  double* A_hom;           // [n_pts][m*m] homogenised tensor (nullptr: not wanted)
  double* S_loc;           // [n_pts][nb*nb] macro element matrices (nullptr: not wanted)
  int* iters;              // [n_pts] PCG iterations (max over right-hand sides), or nullptr
  double* resid;           // [n_pts] final relative residual (max over rhs), or nullptr
  const double* qp;        // [T][nq][D] quadrature points, cube-local, in units of h
  const double* qw;        // [nq] weights normalised to sum 1
  double* scratch;         // per-CTA global scratch (elasticity: corrector vectors)
  double* chi;             // [n_pts][n_rhs][bs][N_grid] correctors on the (possibly collapsed) periodic grid, or nullptr
  unsigned long long* work;  // device counter += sum over right-hand sides of PCG iterations (nullptr: off)
  int nq;
  int max_it;
  double rtol, atol;
#if !defined(HMX_VARIANT) || HMX_VARIANT == 5
  // element-list kernel only (device pointer).  The host always passes the full structure; the other kernels are
  // compiled with the prefix above (a kernel reads as many parameter bytes as it declares).
  const MicroMesh* mesh;
#endif
};

// 16-byte word for packed index tables (one LDS.128)
struct __attribute__((aligned(16))) U4 {
  unsigned x, y, z, w;
};

// compile-time integer passed as a value (generic lambdas over stencil directions)
template <int V>
struct IntTag {
  static constexpr int value = V;
};

// ---- Kuhn tables ------------------------------------------------------------------
template <int D>
HMX_HOSTDEV constexpr int kuhn_ntypes() { return D == 2 ? 2 : 6; }

// axis walked at step k of type t.  3-D order (matches the tetrahedra listed for
// dolfinx create_box in SURVEY.md A.4): (x,y,z) (x,z,y) (z,x,y) (y,x,z) (z,y,x) (y,z,x)
template <int D>
HMX_HOSTDEV constexpr int kuhn_axis(int t, int k) {
  if (D == 2) return t == 0 ? k : 1 - k;
  constexpr unsigned long long packed = (36ULL) | (24ULL << 6) | (18ULL << 12) | (33ULL << 18) | (6ULL << 24) | (9ULL << 30);
  return (int)((packed >> (6 * t + 2 * k)) & 3ULL);
}
// bit mask of the axes walked before vertex a
template <int D>
HMX_HOSTDEV constexpr int kuhn_pmask(int t, int a) {
  int m = 0;
  for (int k = 0; k < a; ++k) m |= 1 << kuhn_axis<D>(t, k);
  return m;
}

HMX_HOSTDEV constexpr int ipow(int b, int e) {  // (a loop: a recursive device function has no static stack bound)
  int r = 1;
  for (int k = 0; k < e; ++k) r *= b;
  return r;
}
HMX_HOSTDEV constexpr int popcount3(int m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1); }
HMX_HOSTDEV constexpr int sym_index(int D, int i, int j) {  // upper triangle, row major
  return i <= j ? i * D - i * (i - 1) / 2 + (j - i) : j * D - j * (j - 1) / 2 + (i - j);
}

// Periodic node grid.  COLL is a bit mask of COLLAPSED axes: when the coefficient does not depend
// on y_a (CO::YDEP), operator and load vectors are invariant under translation along axis a (the
// Kuhn mesh is too), so the correctors are constant along that axis and the cell problem is solved
// EXACTLY on a single layer of cubes (extent 1, periodic onto itself) with the element volume
// multiplied by the number of layers.  COLL = 0 is the full NM^D grid.
template <int D, int NM, int COLL = 0>
struct Grid {
  HMX_HOSTDEV static constexpr int ext(int a) { return (a < D && !((COLL >> a) & 1)) ? NM : 1; }
  static constexpr int N = ext(0) * ext(1) * ext(2);
  static constexpr int NLAYERS = ipow(NM, popcount3(COLL & ((1 << D) - 1)));
  HMX_DEV static void decode(int i, int (&c)[3]) {
    c[0] = i % ext(0);
    c[1] = (i / ext(0)) % ext(1);
    c[2] = i / (ext(0) * ext(1));
  }
  HMX_DEV static int up(int v, int a) { return v + 1 >= ext(a) ? 0 : v + 1; }
  HMX_DEV static int down(int v, int a) { return v == 0 ? ext(a) - 1 : v - 1; }
  HMX_DEV static int index(int c0, int c1, int c2) { return c0 + ext(0) * (c1 + ext(1) * c2); }
  // node at c + mask (sign=+1) or c - mask (sign=-1); mask has one bit per axis
  template <int SIGN>
  HMX_DEV static int shifted(const int (&c)[3], int mask) {
    int s[3];
    HMX_UNROLL
    for (int a = 0; a < 3; ++a) s[a] = (a < D && ((mask >> a) & 1)) ? (SIGN > 0 ? up(c[a], a) : down(c[a], a)) : c[a];
    return index(s[0], s[1], s[2]);
  }
  template <int SIGN>
  HMX_DEV static void shift_coords(const int (&c)[3], int mask, int (&s)[3]) {
    HMX_UNROLL
    for (int a = 0; a < 3; ++a) s[a] = (a < D && ((mask >> a) & 1)) ? (SIGN > 0 ? up(c[a], a) : down(c[a], a)) : c[a];
  }
};

// Parity-major node layout used by the matrix-free elasticity kernel: nodes are grouped by the
// parity of their coordinates (2^D classes of H^D slots, H = ceil(NM/2)).  The threads of a warp
// sweep cubes of ONE colour (same parities), i.e. origins 2 apart along each axis: in the natural
// layout corner b of those cubes sits at stride-2/-2NM/-2NM^2 addresses (8-way bank conflicts for
// n = 8, measured: 66 % of the shared-memory wavefronts of the first version were conflicts); in
// this layout they are consecutive doubles.  For odd NM the classes are padded (decode -> valid).
template <int D, int NM, int COLL = 0>
struct PGrid {
  using G = Grid<D, NM, COLL>;
  HMX_HOSTDEV static constexpr int H(int a) { return (G::ext(a) + 1) / 2; }
  // parity classes only along axes that have more than one node
  HMX_HOSTDEV static constexpr int bit(int a) {
    int b = 0;
    for (int k = 0; k < a; ++k) b += G::ext(k) > 1 ? 1 : 0;
    return b;
  }
  static constexpr int NCLS = 1 << bit(3);
  static constexpr int HC = H(0) * H(1) * H(2);
  static constexpr int NP = NCLS * HC;
  HMX_DEV static int index(const int (&c)[3]) {
    int cls = 0, idx = 0, s = 1;
    HMX_UNROLL
    for (int a = 0; a < D; ++a) {
      if (G::ext(a) > 1) cls |= (c[a] & 1) << bit(a);
      idx += (c[a] >> 1) * s;
      s *= H(a);
    }
    return cls * HC + idx;
  }
  HMX_DEV static bool decode(int ip, int (&c)[3]) {
    const int cls = ip / HC;
    int r = ip - cls * HC;
    bool ok = true;
    HMX_UNROLL
    for (int a = 0; a < 3; ++a) {
      c[a] = 0;
      if (a < D) {
        c[a] = 2 * (r % H(a)) + (G::ext(a) > 1 ? ((cls >> bit(a)) & 1) : 0);
        r /= H(a);
        ok = ok && c[a] < G::ext(a);
      }
    }
    return ok;
  }
};

// Atoms (the y-dependent scalars of the coefficient) are stored per element on the
// REDUCED cube set: axes none of the atoms depend on are collapsed, so e.g. a laminate
// a(y0) costs NM*T evaluations per macro point instead of NM^D*T.
// PARITY = true stores them parity-major like PGrid (restricted to the axes they depend on).
template <int D, int NM, int YDEP, bool PARITY = false>
struct AtomIdx {
  static constexpr int NDEP = popcount3(YDEP & ((1 << D) - 1));
  static constexpr int H = (NM + 1) / 2;
  static constexpr int HC = ipow(H, NDEP);
  static constexpr int NRC = PARITY ? (1 << NDEP) * HC : ipow(NM, NDEP);
  HMX_DEV static int ridx(const int (&c)[3]) {
    if (PARITY) {
      int cls = 0, idx = 0, s = 1, bit = 0;
      HMX_UNROLL
      for (int a = 0; a < D; ++a)
        if ((YDEP >> a) & 1) {
          cls |= (c[a] & 1) << bit;
          idx += (c[a] >> 1) * s;
          s *= H;
          ++bit;
        }
      return cls * HC + idx;
    }
    int r = 0, s = 1;
    HMX_UNROLL
    for (int a = 0; a < D; ++a)
      if ((YDEP >> a) & 1) {
        r += c[a] * s;
        s *= NM;
      }
    return r;
  }
  // reduced slot -> cube coordinates (collapsed axes 0); false for the padding slots of odd NM
  HMX_DEV static bool rdecode(int rc, int (&c)[3]) {
    bool ok = true;
    if (PARITY) {
      const int cls = rc / HC;
      int r = rc - cls * HC, bit = 0;
      HMX_UNROLL
      for (int a = 0; a < 3; ++a) {
        c[a] = 0;
        if (a < D && ((YDEP >> a) & 1)) {
          c[a] = 2 * (r % H) + ((cls >> bit) & 1);
          r /= H;
          ++bit;
          ok = ok && c[a] < NM;
        }
      }
      return ok;
    }
    HMX_UNROLL
    for (int a = 0; a < 3; ++a) {
      c[a] = 0;
      if (a < D && ((YDEP >> a) & 1)) {
        c[a] = rc % NM;
        rc /= NM;
      }
    }
    return ok;
  }
};

// ---- block reductions ---------------------------------------------------------------
// Sum NV values per thread over the CTA; every thread ends with the same totals (fixed
// summation order -> deterministic).  One BAR.SYNC.  `buf` holds NW*NV doubles; callers
// alternate between two buffers so that no trailing barrier is needed.
// Warp sums of NV values at once with a TRANSPOSING butterfly: while a lane still holds more than one value it keeps
// one half of them (which half: its lane bit of the step) and hands the other half to its partner, so every step
// moves half as many values as the one before; the steps that are left sum the single remaining value.
// P - 1 + 5 - log2(P) shuffles (P = NV rounded up to a power of two) instead of 5 NV: 6 instead of 15 for NV = 3, 31
// instead of 90 for NV = 18.  The total of value warp_multi_index<NV>(lane) is returned (every lane of the warp calls).
template <int NV>
struct WarpMulti {
  static_assert(NV >= 1 && NV <= 32, "at most one value per lane");
  static constexpr int P = NV <= 1 ? 1 : NV <= 2 ? 2 : NV <= 4 ? 4 : NV <= 8 ? 8 : NV <= 16 ? 16 : 32;
};
template <int NV>
HMX_DEV int warp_multi_index(int lane) {
  int k = 0;
  HMX_UNROLL
  for (int st = 0; st < 5; ++st)
    if ((WarpMulti<NV>::P >> st) > 1 && (lane & (16 >> st))) k += WarpMulti<NV>::P >> (st + 1);
  return k;
}
template <int NV>
HMX_DEV double warp_sum_multi(const double (&v)[NV], int lane) {
  constexpr int P = WarpMulti<NV>::P;
  double w[P];
  HMX_UNROLL
  for (int k = 0; k < P; ++k) w[k] = k < NV ? v[k < NV ? k : 0] : 0.0;
  HMX_UNROLL
  for (int st = 0; st < 5; ++st) {
    const int m = 16 >> st, cnt = P >> st;  // values per lane before this step (0: one value, plain butterfly)
    if (cnt > 1) {
      const bool up = (lane & m) != 0;
      HMX_UNROLL
      for (int k = 0; k < (cnt >> 1); ++k) {
        const double lo = w[k], hi = w[k + (cnt >> 1)];
        w[k] = (up ? hi : lo) + lane_xor(up ? lo : hi, m);
      }
    } else {
      w[0] += lane_xor(w[0], m);
    }
  }
  return w[0];
}

// first half of block_sum: the warp totals land in buf[warp * NV + k]; one BAR.SYNC.  For callers in which only a
// few threads need the CTA totals (they add the NW partials in warp order themselves).
template <int NV, int NW>
HMX_DEV void block_partials(const double (&v)[NV], double* buf) {
  const int lane = tid() & 31, warp = tid() >> 5;
  {
    const double t = warp_sum_multi<NV>(v, lane);
    const int k = warp_multi_index<NV>(lane);
    if ((lane & (32 / WarpMulti<NV>::P - 1)) == 0 && k < NV) buf[warp * NV + k] = t;
  }
  sync();
}
template <int NV, int NW>
HMX_DEV void block_sum(double (&v)[NV], double* buf) {
  block_partials<NV, NW>(v, buf);
  HMX_UNROLL
  for (int k = 0; k < NV; ++k) {
    double s = 0.0;
    for (int w = 0; w < NW; ++w) s += buf[w * NV + k];
    v[k] = s;
  }
}

// ---- macro element epilogue -----------------------------------------------------------
// P1 gradients of the macro simplex and |T| (hmm.py:20-28 of the reference), then
//   S_loc[i][j] = |T| sum_pq C[p][j] A_hom[p][q] C[q][i]
// with C[:,i] the gradient (Poisson) or engineering-Voigt strain (elasticity, unrolled
// dof i = a*D + k, hmm.py:31-40) of macro basis function i  (SURVEY.md A.3).
template <int D, int KIND>
HMX_DEV double macro_strain_matrix(const double* verts /* (D+1) x 3 */,
                                   double (&C)[KIND == 0 ? D : D * (D + 1) / 2][KIND == 0 ? D + 1 : (D + 1) * D]) {
  constexpr int NV = D + 1;
  double J[D][D];  // columns are edge vectors
  HMX_UNROLL
  for (int r = 0; r < D; ++r)
    HMX_UNROLL
    for (int c = 0; c < D; ++c) J[r][c] = verts[(c + 1) * 3 + r] - verts[r];
  double Ji[D][D], det;
  if (D == 2) {
    det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double id = 1.0 / det;
    Ji[0][0] = J[1][1] * id;
    Ji[0][1] = -J[0][1] * id;
    Ji[1][0] = -J[1][0] * id;
    Ji[1][1] = J[0][0] * id;
  } else {
    const double c00 = J[1][1] * J[2 % D][2 % D] - J[1][2 % D] * J[2 % D][1];
    const double c01 = J[1][2 % D] * J[2 % D][0] - J[1][0] * J[2 % D][2 % D];
    const double c02 = J[1][0] * J[2 % D][1] - J[1][1] * J[2 % D][0];
    det = J[0][0] * c00 + J[0][1] * c01 + J[0][2 % D] * c02;
    const double id = 1.0 / det;
    Ji[0][0] = c00 * id;
    Ji[1][0] = c01 * id;
    Ji[2 % D][0] = c02 * id;
    Ji[0][1] = (J[0][2 % D] * J[2 % D][1] - J[0][1] * J[2 % D][2 % D]) * id;
    Ji[1][1] = (J[0][0] * J[2 % D][2 % D] - J[0][2 % D] * J[2 % D][0]) * id;
    Ji[2 % D][1] = (J[0][1] * J[2 % D][0] - J[0][0] * J[2 % D][1]) * id;
    Ji[0][2 % D] = (J[0][1] * J[1][2 % D] - J[0][2 % D] * J[1][1]) * id;
    Ji[1][2 % D] = (J[0][2 % D] * J[1][0] - J[0][0] * J[1][2 % D]) * id;
    Ji[2 % D][2 % D] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
  }
  const double vol = fabs(det) / (D == 2 ? 2.0 : 6.0);
  // G[p][a] = d phi_a / d x_p : rows of J^-1 are the gradients of barycentric coords 1..D
  double G[D][NV];
  HMX_UNROLL
  for (int p = 0; p < D; ++p) {
    double s = 0.0;
    HMX_UNROLL
    for (int a = 1; a < NV; ++a) {
      G[p][a] = Ji[a - 1][p];
      s += Ji[a - 1][p];
    }
    G[p][0] = -s;
  }
  if (KIND == 0) {
    HMX_UNROLL
    for (int p = 0; p < D; ++p)
      HMX_UNROLL
      for (int a = 0; a < NV; ++a) C[p][a] = G[p][a];
  } else {
    HMX_UNROLL
    for (int a = 0; a < NV; ++a)
      HMX_UNROLL
      for (int k = 0; k < D; ++k) {
        const int i = a * D + k;
        HMX_UNROLL
        for (int q = 0; q < D; ++q) C[q][i] = (q == k) ? G[q][a] : 0.0;
        int v = D;
        HMX_UNROLL
        for (int r = 0; r < D; ++r)
          HMX_UNROLL
          for (int c = r + 1; c < D; ++c) {
            C[v][i] = ((r == k) ? G[c][a] : 0.0) + ((c == k) ? G[r][a] : 0.0);
            ++v;
          }
      }
  }
  return vol;
}

template <int D, int KIND>
HMX_DEV void macro_element_matrix(const double* verts /* (D+1) x 3 */, const double* Ahom, double* S) {
  constexpr int MV = KIND == 0 ? D : D * (D + 1) / 2;
  constexpr int NB = KIND == 0 ? D + 1 : (D + 1) * D;
  double C[MV][NB];
  const double vol = macro_strain_matrix<D, KIND>(verts, C);
  for (int i = 0; i < NB; ++i)
    for (int j = 0; j < NB; ++j) {
      double s = 0.0;
      HMX_UNROLL
      for (int p = 0; p < MV; ++p)
        HMX_UNROLL
        for (int q = 0; q < MV; ++q) s += C[p][j] * Ahom[p * MV + q] * C[q][i];
      S[i * NB + j] = vol * s;
    }
}

// macro point of unit `pt`: given directly, or the barycentre of the macro cell
// (mean of the vertex coordinates, hmm.py:349-352)
template <int D>
HMX_DEV void macro_point(const CellParams& P, long long pt, double (&xm)[3], double (&verts)[(D + 1) * 3]) {
  if (P.cell_nodes != nullptr) {
    HMX_UNROLL
    for (int k = 0; k < 3; ++k) xm[k] = 0.0;
    HMX_UNROLL
    for (int v = 0; v <= D; ++v) {
      const long long node = P.cell_nodes[pt * (D + 1) + v];
      HMX_UNROLL
      for (int k = 0; k < 3; ++k) {
        verts[v * 3 + k] = P.node_xyz[node * 3 + k];
        xm[k] += verts[v * 3 + k];
      }
    }
    HMX_UNROLL
    for (int k = 0; k < 3; ++k) xm[k] /= (double)(D + 1);
  } else {
    HMX_UNROLL
    for (int k = 0; k < 3; ++k) xm[k] = P.x_pts[pt * 3 + k];
    HMX_UNROLL
    for (int v = 0; v < (D + 1) * 3; ++v) verts[v] = 0.0;
  }
}

}  // namespace hmx
