// Micro cell kernel for the linear-elasticity HMM classes: one CTA per macro quadrature point.
//
// Replaces, for LinearElasticityHMM / LinearElasticityStratifiedHMM, what the reference does in
// BaseHMM._compute_local_stiffness (hmm.py:334-369) with the forms of hmm.py:887-922 /
// 1024-1067: D(D+1)/2 unit-strain corrector problems instead of n_b = D(D+1).
//
// The periodic P1 stiffness matrix of the vector problem (D x D blocks on the 7-/15-point
// stencil: 135 doubles per node in 3-D) does not fit in shared memory, so the operator is applied
// MATRIX-FREE from the reference element kernel with the coefficient evaluated in place:
//   per cube and right-hand side: gather the 2^D corner displacements, and per simplex
//   H = n sum_k Mcol_{pi(k)} (x) (u_{P(k+1)} - u_{P(k)}),  e = sym(H)  (engineering Voigt),
//   sigma = C(atoms of this element) : e   (generated `stress`, constant-folded for e.g. Hooke),
//   t_k = |e| n sigma Mcol_{pi(k)};   y_{P(k+1)} += t_k,  y_{P(k)} -= t_k.
// Cubes are processed colour by colour (2^D colours for even n) so that the in-place
// accumulation into the shared-memory result needs no atomics and is deterministic.
// Threads are grouped by right-hand side (NT = NRHS * TPR).  The right-hand sides are independent
// linear systems, so each group runs ITS OWN PCG loop and synchronises only with itself through a
// named barrier (BAR.SYNC id, TPR): the groups drift apart and fill each other's barrier and
// reduction bubbles on the FP64 pipe.  Search directions p and the product y = K p live in shared
// memory, the residuals and correctors in an L2-resident per-CTA scratch.  Structurally zero
// entries of M (known when the coefficient program is generated, CO::MZERO) are dropped from the
// element kernel at compile time.
// Preconditioner: D x D block Jacobi.  Epilogue as in the Poisson kernel:
//   A_hom[p][q] = <C>[p][q] - b_p.x_q - x_p.r_q.
#pragma once
#include "hmx_cell_common.cuh"
#include "hmx_cell_coarse.cuh"

#ifndef HMX_STAGGER
#define HMX_STAGGER 0  // SM clocks between the PCG starts of consecutive right-hand-side groups
#endif

namespace hmx {

// VGLOB = 1: the search directions p and the product y = K p live in the L2-resident scratch instead of
// shared memory -- the fallback for cells whose vectors exceed 227 KB (3-D elasticity, n >= 10): slower
// (every gather / accumulation goes to L2) but any cell size runs.
template <class CO, int NM, int NT, int COLL = 0, int VGLOB = 0>
struct ElasticityLayout {
  static constexpr int D = CO::DIM;
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int N = Grid<D, NM, COLL>::N;
  static constexpr int NRHS = D * (D + 1) / 2;
  static constexpr int NV = NRHS;  // Voigt length
  static constexpr int NP = PGrid<D, NM, COLL>::NP;  // node slots of the parity-major layout (>= N)
  static constexpr int NDOF = NP * D;
  static constexpr int TPR = NT / NRHS;  // threads per right-hand side
  // small (e.g. axis-collapsed) cells have fewer cubes per colour than a warp has lanes: then a
  // right-hand side gets a 16-/8-lane SEGMENT of a warp, the warp is the synchronisation group,
  // reductions are segmented shuffles and the right-hand sides sharing a warp iterate together
  static constexpr bool SUBW = TPR < 32;
  static constexpr int TPG = SUBW ? 32 : TPR;  // threads per synchronisation group
  static constexpr int NW = NT / 32;
  static constexpr int WPR = SUBW ? 1 : TPR / 32;  // warps per right-hand side
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  static constexpr int NSYM = D * (D + 1) / 2;
  static constexpr int NRC = AtomIdx<D, NM, CO::YDEP, true>::NRC;
  static constexpr int NREDV = 2 * NRHS > NA1 ? 2 * NRHS : NA1;
  static constexpr int o_red = 0;                              // 2 buffers [NW][NREDV]
  static constexpr int NSLOT = NW > NRHS * WPR ? NW : NRHS * WPR;
  static constexpr int o_stat = o_red + 2 * NSLOT * NREDV;     // [NRHS][4] per right-hand side: its, rz, rz0
  static constexpr int o_atoms = o_stat + 4 * NRHS;            // [NA][T][NRC]
  static constexpr int o_dinv = o_atoms + NA1 * T * NRC;       // [NSYM][NP] inverse diagonal blocks
  static constexpr int o_p = o_dinv + NSYM * NP;               // [NRHS][D][NP]
  static constexpr int o_y = o_p + (VGLOB ? 0 : NRHS * NDOF);  // [NRHS][D][N]
  // per-cube index table (even NM, no collapsed axis): the corner node slots and the atom slot of every cube in
  // sweep order, 16-bit each, so the sweep does one or two 16-byte loads instead of ~80 integer instructions
  static constexpr bool TAB = COLL == 0 && NM % 2 == 0 && NP <= 65535 && NRC <= 65535;
  // 2-D: one 16-byte word per cube (4 corners + atom slot).  3-D: a 16-byte word of corners per cube, then
  // (structure of arrays, so that neither load has bank conflicts) one 32-bit atom slot per cube
  static constexpr int o_tab = ((o_y + (VGLOB ? 0 : NRHS * NDOF) + 1) / 2) * 2;
  static constexpr int o_tab_ro = o_tab + N * 2;  // 3-D only
  // 3-D cells swept warp-slab-wise take 2x2x1 blocks of cubes per thread (elasticity_sweep_blocks)
#ifndef HMX_NO_BLOCK_SWEEP
  static constexpr bool BLK = D == 3 && TAB && !SUBW && (NM / 2) % WPR == 0;
#else
  static constexpr bool BLK = false;
#endif
  // ... and when the blocks of one plane sit in one warp pass, the first contribution to a node of each sweep is
  // a plain store: y needs no zeroing between the sweeps
  static constexpr bool STORE1 = BLK && 32 % ((NM / 2) * (NM / 2)) == 0;
  // two-level preconditioner (hmx_cell_coarse.cuh): inverse coarse matrix, parent table, one column buffer
  static constexpr int o_ei = o_tab + (TAB ? N * 2 + (D == 3 ? (N + 1) / 2 : 0) : 0);  // [NTRI] packed lower triangle
  using CS = CoarseSpace<CO, NM, NT, COLL, VGLOB, o_ei>;
  static constexpr int o_par = o_ei + (CS::ON ? CS::NTRI : 0);                           // [NP] 2 x 16-bit level-1 slots
  static constexpr int o_cbuf = o_par + (CS::ON ? (NP + 1) / 2 : 0);                     // [CS::CBUF]
  static constexpr int o_epi = ((o_cbuf + (CS::ON ? CS::CBUF : 0) + 1) / 2) * 2;  // epilogue: A_hom, macro strain matrix, |T|
  static constexpr int total = o_epi + NRHS * NRHS + NV * (D + 1) * D + 2;
  static_assert(!CS::ON || (CS::setup_doubles <= 2 * NRHS * NDOF && !SUBW), "coarse set-up scratch must fit in the p / y area");
  static constexpr int scratch_doubles = (VGLOB ? 4 : 2) * NRHS * NDOF;  // x and r (and p, y) per CTA
  static_assert(NT % NRHS == 0 && NT % 32 == 0 && (TPR % 32 == 0 || 32 % TPR == 0),
                "block size must be NRHS * (a multiple or a divisor of 32)");
};

// inverse of a symmetric DxD matrix given by its upper triangle (row major)
template <int D>
HMX_DEV void sym_inverse(const double* a, double* inv) {
  if (D == 2) {
    const double det = a[0] * a[2] - a[1] * a[1];
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    inv[0] = a[2] * id;
    inv[1] = -a[1] * id;
    inv[2] = a[0] * id;
  } else {
    // a = [a00 a01 a02 a11 a12 a22]
    const double a00 = a[0], a01 = a[1], a02 = a[2 % (D * (D + 1) / 2)], a11 = a[3 % (D * (D + 1) / 2)],
                 a12 = a[4 % (D * (D + 1) / 2)], a22 = a[5 % (D * (D + 1) / 2)];
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    inv[0] = c00 * id;
    inv[1] = c01 * id;
    inv[2 % (D * (D + 1) / 2)] = c02 * id;
    inv[3 % (D * (D + 1) / 2)] = (a00 * a22 - a02 * a02) * id;
    inv[4 % (D * (D + 1) / 2)] = (a01 * a02 - a00 * a12) * id;
    inv[5 % (D * (D + 1) / 2)] = (a00 * a11 - a01 * a01) * id;
  }
}

// The T simplices of one cube: acc += K_cube u  (RHSMODE: acc += the load of the unit strain -E_q).
// u, acc: [corner][component]; ro: the cube's atom slot.
template <class CO, bool RHSMODE, int COLL, int NRC_>
HMX_DEV void cube_apply(const double* pc, const double (&Ms)[CO::DIM * CO::DIM], const double* s_atoms, int ro,
                             const double (&u)[1 << CO::DIM][CO::DIM], double (&acc)[1 << CO::DIM][CO::DIM], int q,
                             double sqrtw) {
  constexpr int D = CO::DIM, T = kuhn_ntypes<D>(), NV = D * (D + 1) / 2, NA = CO::NATOMS, NA1 = NA > 0 ? NA : 1;
  constexpr int NRC = NRC_, NC = 1 << D, CM = COLL & (NC - 1);
#define HMX_MZ(p_, ax_) ((CO::MZERO >> ((p_)*D + (ax_))) & 1u)
  HMX_UNROLL
  for (int t = 0; t < T; ++t) {
    double e[NV];
    if (RHSMODE) {
      HMX_UNROLL
      for (int v = 0; v < NV; ++v) e[v] = (v == q) ? -sqrtw : 0.0;
    } else {
      // e = sym(H) accumulated directly in engineering Voigt form,
      // H[p][j] = sum_k Ms[p][pi(k)] * (u[P(k+1)][j] - u[P(k)][j]):  e[voigt(p,j)] += Ms[p][pi(k)] dk[j]
      HMX_UNROLL
      for (int v = 0; v < NV; ++v) e[v] = 0.0;
      HMX_UNROLL
      for (int k2 = 0; k2 < D; ++k2) {
        const int ax = kuhn_axis<D>(t, k2);
        if ((CM >> ax) & 1) continue;  // collapsed axis: the difference is identically zero
        const int b0 = kuhn_pmask<D>(t, k2) & ~CM, b1 = kuhn_pmask<D>(t, k2 + 1) & ~CM;
        HMX_UNROLL
        for (int j = 0; j < D; ++j) {
          const double dk = u[b1][j] - u[b0][j];
          HMX_UNROLL
          for (int p = 0; p < D; ++p)
            if (!HMX_MZ(p, ax)) {
              // Voigt slot of the (p, j) entry: diagonal -> p, off-diagonal pairs in the order (0,1)[,(0,2),(1,2)]
              const int lo = p < j ? p : j, hi = p < j ? j : p;
              const int v = p == j ? p : D + (lo == 0 ? hi - 1 : 2);
              e[v] += Ms[p * D + ax] * dk;
            }
        }
      }
    }
    double sa[NA1], sig[NV];
    HMX_UNROLL
    for (int k2 = 0; k2 < NA1; ++k2) sa[k2] = NA > 0 ? s_atoms[(k2 * T + t) * NRC + ro] : 0.0;
    CO::stress(pc, sa, e, sig);
    double S[D][D];
    HMX_UNROLL
    for (int v = 0; v < D; ++v) S[v][v] = sig[v];
    {
      int v = D;
      HMX_UNROLL
      for (int r = 0; r < D; ++r)
        HMX_UNROLL
        for (int c = r + 1; c < D; ++c) {
          S[r][c] = S[c][r] = sig[v];
          ++v;
        }
    }
    HMX_UNROLL
    for (int k2 = 0; k2 < D; ++k2) {
      const int ax = kuhn_axis<D>(t, k2);
      if ((CM >> ax) & 1) continue;  // both ends are the same node: the two forces cancel
      const int b0 = kuhn_pmask<D>(t, k2) & ~CM, b1 = kuhn_pmask<D>(t, k2 + 1) & ~CM;
      // single-term columns (e.g. the rotation axis of C4's Jacobian) go straight into the two
      // accumulators as FMAs; longer ones are summed once and added / subtracted
      int nterms = 0;
      HMX_UNROLL
      for (int p = 0; p < D; ++p) nterms += HMX_MZ(p, ax) ? 0 : 1;
      HMX_UNROLL
      for (int j = 0; j < D; ++j) {
        if (nterms == 1) {
          HMX_UNROLL
          for (int p = 0; p < D; ++p)
            if (!HMX_MZ(p, ax)) {
              acc[b1][j] += S[j][p] * Ms[p * D + ax];
              acc[b0][j] -= S[j][p] * Ms[p * D + ax];
            }
        } else {
          double tk = 0.0;
          HMX_UNROLL
          for (int p = 0; p < D; ++p)
            if (!HMX_MZ(p, ax)) tk += S[j][p] * Ms[p * D + ax];
          acc[b1][j] += tk;
          acc[b0][j] -= tk;
        }
      }
    }
  }
#undef HMX_MZ
}

// entry `idx` of the per-cube index table (ElasticityLayout::TAB): corner node slots and atom slot
template <int D>
HMX_DEV void table_entry(const U4* s_tab, int ncubes, int idx, int (&node)[1 << D], int& ro) {
  const U4 w0 = s_tab[idx];
  const unsigned w[4] = {w0.x, w0.y, w0.z, w0.w};
  HMX_UNROLL
  for (int b = 0; b < (1 << D); ++b) node[b] = (int)((w[b >> 1] >> (16 * (b & 1))) & 0xffffu);
  ro = D == 3 ? (int)reinterpret_cast<const unsigned*>(s_tab + ncubes)[idx] : (int)(w[2] & 0xffffu);
}

// Block sweep (3-D, SLAB scheme with the index table):  y += K p  with every thread taking a 2x2x1 BLOCK of cubes
// per last-axis parity instead of one cube per colour.  The four cubes are visited in the order
// A=(0,0) -> B=(1,0) -> C=(1,1) -> D=(0,1); the face shared by consecutive cubes keeps its displacements and its
// partial forces in registers, so the block costs 20 node loads of p and 20 node updates of y where four separate
// colours cost 32 + 32 (-37 % shared-memory traffic, the second ceiling of this kernel after the FP64 pipe).
// A node is written back as soon as this thread has no further contribution to it.  Nodes shared with the
// neighbouring blocks (block-local index 0 or 2 along x1 / x2; all of them belong to this warp, which owns whole
// planes of the last axis) are updated in four phases FA..FD, one per cube, separated by __syncwarp: in one phase
// all threads update the same block-local nodes, i.e. distinct nodes of the cell:
//   FA: (0,0) and a first part of (0,1)   FB: (1,0), (2,0)   FC: (2,1), (2,2)   FD: (0,1), (1,1), (0,2), (1,2).
// The old value of y is read at the START of its phase, straight into the accumulator the cube then adds to, and
// the phase ends with plain stores: the read latency hides behind the cube's arithmetic and costs no register.
template <class CO, int NM, int NT, int COLL, int VGLOB>
HMX_DEV void elasticity_sweep_blocks(const double* pc, const double (&Ms)[CO::DIM * CO::DIM], const double* s_atoms,
                                     const double* s_p, double* s_y, int q, int l, double sqrtw, const U4* s_tab) {
  using L = ElasticityLayout<CO, NM, NT, COLL, VGLOB>;
  constexpr int D = 3, N = L::NP, NRC = L::NRC, NC = 8;
  constexpr int HALF = NM / 2, TOT = HALF * HALF * HALF;  // cubes per colour
  constexpr int SLABSZ = TOT / L::WPR;
  static_assert(CO::DIM == 3 && COLL == 0 && NM % 2 == 0 && (NM / 2) % L::WPR == 0 && !L::SUBW && L::TAB, "block sweep");
  const int wig = l >> 5, lig = l & 31;
  const double* p_q = s_p + (size_t)q * D * N;
  double* y_q = s_y + (size_t)q * D * N;
  // displacements of the corners with (b & mask) == val
  auto load = [&](double (&u)[NC][D], const int (&node)[NC], int mask, int val) {
    HMX_UNROLL
    for (int b = 0; b < NC; ++b)
      if ((b & mask) == val) {
        HMX_UNROLL
        for (int j = 0; j < D; ++j) u[b][j] = p_q[j * N + node[b]];
      }
  };
  // start of a phase: the corners (b & mask) == val are completed by this cube -> their accumulators take up the
  // old y.  `carried` corners already hold the previous cube's part; `fresh` ones are the first contribution of
  // the whole sweep to their node when `first` (L::STORE1): nothing to read
  auto open = [&](double (&acc)[NC][D], const int (&node)[NC], int mask, int val, int carried, bool first, int fresh) {
    HMX_UNROLL
    for (int b = 0; b < NC; ++b) {
      const bool sel = (b & mask) == val, car = (carried >> b) & 1;
      HMX_UNROLL
      for (int j = 0; j < D; ++j) {
        const bool old_y = sel && !(L::STORE1 && ((fresh >> b) & 1) && first);
        if (car) {
          if (old_y) acc[b][j] += y_q[j * N + node[b]];
        } else {
          acc[b][j] = old_y ? y_q[j * N + node[b]] : 0.0;
        }
      }
    }
  };
  auto close = [&](const double (&acc)[NC][D], const int (&node)[NC], int mask, int val) {
    HMX_UNROLL
    for (int b = 0; b < NC; ++b)
      if ((b & mask) == val) {
        HMX_UNROLL
        for (int j = 0; j < D; ++j) y_q[j * N + node[b]] = acc[b][j];
      }
  };
  for (int cz = 0; cz < 2; ++cz) {
    const bool first = cz == 0;  // parity 0 reaches every node of the cell
    for (int k0 = wig * SLABSZ; k0 < (wig + 1) * SLABSZ; k0 += 32) {
      const int k = k0 + lig;
      const bool on = k < (wig + 1) * SLABSZ;
      const int e0 = 4 * cz * TOT + (on ? k : k0);  // colour index = cx + 2 cy + 4 cz, TOT entries per colour
      int node[NC], ro, node2[NC], ro2;
      double u[NC][D], acc[NC][D], v[NC][D], bcc[NC][D];
      // A = (cx, cy) = (0, 0); completes its x1 = 0 face [(0,0), and (0,1) for now]
      table_entry<D>(s_tab, L::N, e0, node, ro);
      load(u, node, 0, 0);
      open(acc, node, 1, 0, 0x00, first, 0x55);
      table_entry<D>(s_tab, L::N, e0 + TOT, node2, ro2);
      cube_apply<CO, false, COLL, NRC>(pc, Ms, s_atoms, ro, u, acc, q, sqrtw);
      load(v, node2, 1, 1);
      if (on) close(acc, node, 1, 0);
      warp_sync();
      // B = (1, 0): its x1 = 0 face is A's x1 = 1 face; completes its x2 = 0 face [(1,0), (2,0)]
      HMX_UNROLL
      for (int b = 0; b < NC; b += 2)
        HMX_UNROLL
        for (int j = 0; j < D; ++j) {
          v[b][j] = u[b | 1][j];
          bcc[b][j] = acc[b | 1][j];
        }
      open(bcc, node2, 2, 0, 0x55, first, 0x11);
      table_entry<D>(s_tab, L::N, e0 + 3 * TOT, node, ro);
      cube_apply<CO, false, COLL, NRC>(pc, Ms, s_atoms, ro2, v, bcc, q, sqrtw);
      load(u, node, 2, 2);
      if (on) close(bcc, node2, 2, 0);
      warp_sync();
      // C = (1, 1): its x2 = 0 face is B's x2 = 1 face; completes its x1 = 1 face [(2,1), (2,2)]
      HMX_UNROLL
      for (int b = 0; b < NC; ++b)
        if (!(b & 2)) {
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            u[b][j] = v[b | 2][j];
            acc[b][j] = bcc[b | 2][j];
          }
        }
      open(acc, node, 1, 1, 0x33, first, 0);
      table_entry<D>(s_tab, L::N, e0 + 2 * TOT, node2, ro2);
      cube_apply<CO, false, COLL, NRC>(pc, Ms, s_atoms, ro, u, acc, q, sqrtw);
      load(v, node2, 1, 0);
      if (on) close(acc, node, 1, 1);
      warp_sync();
      // D = (0, 1): its x1 = 1 face is C's x1 = 0 face; completes everything it touches [(0,1), (1,1), (0,2), (1,2)]
      HMX_UNROLL
      for (int b = 1; b < NC; b += 2)
        HMX_UNROLL
        for (int j = 0; j < D; ++j) {
          v[b][j] = u[b & ~1][j];
          bcc[b][j] = acc[b & ~1][j];
        }
      open(bcc, node2, 0, 0, 0xaa, first, 0x22);
      cube_apply<CO, false, COLL, NRC>(pc, Ms, s_atoms, ro2, v, bcc, q, sqrtw);
      if (on) close(bcc, node2, 0, 0);
      warp_sync();  // next block of this warp / (with the group barrier) next parity
    }
    group_sync(1 + q, L::TPR);  // the other last-axis parity touches the neighbouring warps' planes
  }
}

// One sweep over the cubes by the TPR threads of right-hand side q:  y += K p  (RHSMODE = false)
// or  y += b_q  (RHSMODE = true: every element carries the unit strain -E_q; hmm.py:898-903).
// Ms = sqrt(|e|) n M, so that e and sigma both carry sqrt(|e|) and the nodal forces the full |e|.
template <class CO, int NM, int NT, bool RHSMODE, int COLL = 0, int VGLOB = 0>
HMX_DEV void elasticity_sweep(const double* pc, const double (&Ms)[CO::DIM * CO::DIM], const double* s_atoms,
                              const double* s_p, double* s_y, int q, int l, double sqrtw, const U4* s_tab) {
  using L = ElasticityLayout<CO, NM, NT, COLL, VGLOB>;
  using G = Grid<CO::DIM, NM, COLL>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP, true>;
  using PG = PGrid<CO::DIM, NM, COLL>;
  constexpr int D = L::D, T = L::T, N = L::NP, NV = L::NV, NA = L::NA, NA1 = L::NA1, NRC = L::NRC;
  constexpr int NC = 1 << D;  // corners
  // colours per axis: even extent -> 2 classes of ext/2 cubes; odd extent -> a third class holding the last index
#define HMX_NCOL(a_) (G::ext(a_) % 2 == 0 ? 2 : 3)
#define HMX_HALF(a_) (G::ext(a_) / 2)
#define HMX_MZ(p_, ax_) ((CO::MZERO >> ((p_)*D + (ax_))) & 1u)

  // SLAB scheme (even NM, last-axis half-count divisible by the warps of the group): colours are
  // ordered with the parity of the last axis outermost and every warp owns a slab of last-axis
  // indices.  Inside one last-axis parity the warps then write disjoint node planes, so the
  // colours of that parity only need the warp's own lock-step order (__syncwarp) and the group
  // barrier is needed twice per sweep instead of 2^D times.
  constexpr bool SLAB = !L::SUBW && COLL == 0 && (NM % 2 == 0) && ((NM / 2) % L::WPR == 0);
  constexpr int NCOLT = (D > 0 ? HMX_NCOL(0) : 1) * (D > 1 ? HMX_NCOL(1) : 1) * (D > 2 ? HMX_NCOL(2) : 1);
  constexpr int NCOL_LAST = HMX_NCOL(D - 1);
  const int wig = l >> 5, lig = l & 31;  // warp within the group, lane
  for (int col = 0; col < NCOLT; ++col) {
    int cls[3], cnt[3], total = 1;
    {
      // the last axis parity is the slowest digit of `col`
      int r = col;
      HMX_UNROLL
      for (int a = 0; a < 3; ++a) {
        cls[a] = a < D ? r % HMX_NCOL(a) : 0;
        if (a < D) r /= HMX_NCOL(a);
        cnt[a] = a < D ? (cls[a] == 2 ? 1 : HMX_HALF(a)) : 1;
        total *= cnt[a];
      }
    }
    // SLAB: this warp's share of the colour = its slab of the last axis
    const int slab = SLAB ? total / L::WPR : total;
    const int kbeg = SLAB ? wig * slab + lig : l;
    const int kend = SLAB ? (wig + 1) * slab : total;
    const int kstep = SLAB ? 32 : L::TPR;
    for (int k = kbeg; k < kend; k += kstep) {
      int o[3];
      {
        int r = k;
        HMX_UNROLL
        for (int a = 0; a < 3; ++a) {
          const int idx = cnt[a] > 0 ? r % cnt[a] : 0;
          if (cnt[a] > 0) r /= cnt[a];
          o[a] = a < D ? (cls[a] == 2 ? G::ext(a) - 1 : 2 * idx + cls[a]) : 0;
        }
      }
      // Along a collapsed axis both ends of an edge are the same node: corners are addressed through their
      // canonical index (collapsed bits cleared), steps along collapsed axes contribute nothing and are
      // dropped at compile time -- a third of the element work for C4's fibre coefficient.
      constexpr int CM = COLL & (NC - 1);
      int node[NC];
      int ro;
      if (L::TAB) {
        table_entry<D>(s_tab, L::N, col * total + k, node, ro);
      } else {
        HMX_UNROLL
        for (int b = 0; b < NC; ++b) {
          node[b] = 0;
          if (b & CM) continue;
          int cb[3];
          G::template shift_coords<1>(o, b, cb);
          node[b] = PG::index(cb);
        }
        ro = AI::ridx(o);
      }
      double u[NC][D], acc[NC][D];
      HMX_UNROLL
      for (int b = 0; b < NC; ++b)
        HMX_UNROLL
        for (int j = 0; j < D; ++j) {
          acc[b][j] = 0.0;
          u[b][j] = (RHSMODE || (b & CM)) ? 0.0 : s_p[(q * D + j) * N + node[b]];
        }
      cube_apply<CO, RHSMODE, COLL, NRC>(pc, Ms, s_atoms, ro, u, acc, q, sqrtw);
      HMX_UNROLL
      for (int b = 0; b < NC; ++b) {
        if (b & CM) continue;
        HMX_UNROLL
        for (int j = 0; j < D; ++j) s_y[(q * D + j) * N + node[b]] += acc[b][j];
      }
    }
    if (SLAB && (col + 1) % (NCOLT / NCOL_LAST) != 0)
      warp_sync();  // next colour has the same last-axis parity: only this warp's order matters
    else if (L::SUBW)
      warp_sync();  // the group is the warp
    else
      group_sync(1 + q, L::TPR);
  }
#undef HMX_MZ
#undef HMX_NCOL
#undef HMX_HALF
}

template <class CO, int NM, int NT, int COLL = 0, int VGLOB = 0>
HMX_DEV void elasticity_cell_body(const CellParams& P) {
  static_assert((COLL & CO::YDEP) == 0, "only axes the coefficient does not depend on can be collapsed");
  using L = ElasticityLayout<CO, NM, NT, COLL, VGLOB>;
  using G = Grid<CO::DIM, NM, COLL>;
  using AI = AtomIdx<CO::DIM, NM, CO::YDEP, true>;
  using PG = PGrid<CO::DIM, NM, COLL>;
  // N counts node SLOTS of the parity-major layout; padding slots (odd NM) hold zeros in every vector
  // and in the preconditioner, so the vector loops need no validity test.
  constexpr int D = L::D, T = L::T, N = L::NP, NRHS = L::NRHS, NV = L::NV, NDOF = L::NDOF, TPR = L::TPR, NW = L::NW;
  constexpr int WPR = L::WPR, NA = L::NA, NA1 = L::NA1, NSYM = L::NSYM, NRC = L::NRC;
  constexpr int NPT = (N + TPR - 1) / TPR;  // node slots per thread within its right-hand side
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;

  double* sm = dyn_smem();
  double* s_red = sm + L::o_red;
  double* s_stat = sm + L::o_stat;
  double* s_atoms = sm + L::o_atoms;
  double* s_dinv = sm + L::o_dinv;
  double* g_x = P.scratch + (size_t)bid() * L::scratch_doubles;  // [NRHS][D][NP]
  double* g_r = g_x + NRHS * NDOF;
  U4* s_tab = reinterpret_cast<U4*>(sm + L::o_tab);
  double* s_p = VGLOB ? g_r + NRHS * NDOF : sm + L::o_p;  // VGLOB: "s_" vectors are in the L2 scratch too
  double* s_y = VGLOB ? g_r + 2 * NRHS * NDOF : sm + L::o_y;
  using CS = typename L::CS;
  constexpr bool TWO = CS::ON;  // additive two-level preconditioner (hmx_cell_coarse.cuh)
  double* s_ei = sm + L::o_ei;
  unsigned* s_par = reinterpret_cast<unsigned*>(sm + L::o_par);
  double* s_cbuf = sm + L::o_cbuf;

  const int t_id = tid();
  const int q = t_id / TPR, l = t_id - q * TPR;
  const int lane = t_id & 31, warp = t_id >> 5;
  const double h = 1.0 / (double)NM;
  // |e| times the number of identical layers a collapsed grid stands for
  const double vol = (D == 2 ? 0.5 * h * h : h * h * h / 6.0) * (double)G::NLAYERS;
  const double sqrtw = sqrt(vol);
  int red_flip = 0;

  if (TWO) {  // level-1 parents of every fine node slot: depends on the micro grid only
    for (int i = t_id; i < N; i += NT) s_par[i] = coarse_parents<CS, PG>(i);
  }
  if (L::TAB) {
    // the cube -> (corner node slots, atom slot) table depends on the micro grid only: built once per launch.
    // Entry order = sweep order: colour (last-axis parity slowest), then the cube index inside the colour.
    constexpr int HALF = NM / 2, TOT = ipow(HALF, D), NC = 1 << D;
    for (int idx = t_id; idx < N; idx += NT) {
      const int col = idx / TOT;
      int k = idx - col * TOT, cc = col, o[3] = {0, 0, 0};
      HMX_UNROLL
      for (int a = 0; a < D; ++a) {
        o[a] = 2 * (k % HALF) + (cc % 2);
        k /= HALF;
        cc /= 2;
      }
      unsigned w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      HMX_UNROLL
      for (int b = 0; b < NC; ++b) {
        int cb[3];
        G::template shift_coords<1>(o, b, cb);
        w[b >> 1] |= (unsigned)PG::index(cb) << (16 * (b & 1));
      }
      w[NC >> 1] |= (unsigned)AI::ridx(o);
      s_tab[idx] = U4{w[0], w[1], w[2], w[3]};
      if (D == 3) reinterpret_cast<unsigned*>(s_tab + L::N)[idx] = w[4];
    }
    sync();
  }

  for (long long pt = bid(); pt < P.n_pts; pt += nblocks()) {
    double xm[3], verts[(D + 1) * 3];
    macro_point<D>(P, pt, xm, verts);
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double Mn[D * D], Ms[D * D];  // n M and sqrt(|e|) n M;  M[p*D+i] = d theta_i / d x_p  (hmm.py:1015-1016)
    CO::dtheta(xm, Mn);
    HMX_UNROLL
    for (int k = 0; k < D * D; ++k) {
      Mn[k] *= (double)NM;
      Ms[k] = Mn[k] * sqrtw;
    }

    // ---- 1. atoms ----
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
    if (!TWO)
      for (int i = t_id; i < NRHS * NDOF; i += NT) s_y[i] = 0.0;  // the accumulation target
    sync();

    // ---- 2. atom means, block-Jacobi preconditioner ----
    double smean[NA1];
    HMX_UNROLL
    for (int k = 0; k < NA1; ++k) smean[k] = 0.0;
    if (NA > 0) {
      for (int idx = t_id; idx < T * NRC; idx += NT) {
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) smean[k] += s_atoms[k * T * NRC + idx];
      }
      block_sum<NA1, NW>(smean, s_red + (red_flip ^= 1) * L::NSLOT * L::NREDV);
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) smean[k] *= 1.0 / (double)(T * ipow(NM, AI::NDEP));  // padding slots hold 0
    }
    for (int i = t_id; i < N; i += NT) {
      int c[3];
      if (!PG::decode(i, c)) {
        HMX_UNROLL
        for (int k = 0; k < NSYM; ++k) s_dinv[k * N + i] = 0.0;
        continue;
      }
      double blk[NSYM];
      HMX_UNROLL
      for (int k = 0; k < NSYM; ++k) blk[k] = 0.0;
      HMX_UNROLL
      for (int t = 0; t < T; ++t) {
        HMX_UNROLL
        for (int a = 0; a <= D; ++a) {
          int o[3];
          G::template shift_coords<-1>(c, kuhn_pmask<D>(t, a), o);
          const int ro = AI::ridx(o);
          double sa[NA1];
          HMX_UNROLL
          for (int k = 0; k < NA1; ++k) sa[k] = NA > 0 ? s_atoms[(k * T + t) * NRC + ro] : 0.0;
          // m = M g_a = Mn (e_{pi(a-1)} - e_{pi(a)})
          double m[D];
          HMX_UNROLL
          for (int p = 0; p < D; ++p) {
            m[p] = 0.0;
            if (a >= 1) m[p] += Mn[p * D + kuhn_axis<D>(t, a >= 1 ? a - 1 : 0)];
            if (a < D) m[p] -= Mn[p * D + kuhn_axis<D>(t, a < D ? a : 0)];
          }
          double ej[D][NV], sg[D][NV];
          HMX_UNROLL
          for (int j = 0; j < D; ++j) {
            HMX_UNROLL
            for (int v = 0; v < D; ++v) ej[j][v] = (v == j) ? m[v] : 0.0;
            int v = D;
            HMX_UNROLL
            for (int r = 0; r < D; ++r)
              HMX_UNROLL
              for (int cc = r + 1; cc < D; ++cc) {
                ej[j][v] = ((cc == j) ? m[r] : 0.0) + ((r == j) ? m[cc] : 0.0);
                ++v;
              }
            CO::stress(pc, sa, ej[j], sg[j]);
          }
          HMX_UNROLL
          for (int j = 0; j < D; ++j)
            HMX_UNROLL
            for (int j2 = j; j2 < D; ++j2) {
              double s = 0.0;
              HMX_UNROLL
              for (int v = 0; v < NV; ++v) s += sg[j][v] * ej[j2][v];
              blk[sym_index(D, j, j2)] += vol * s;
            }
        }
      }
      double inv[NSYM];
      sym_inverse<D>(blk, inv);
      HMX_UNROLL
      for (int k = 0; k < NSYM; ++k) s_dinv[k * N + i] = inv[k];
    }
    if constexpr (TWO) {
      // coarse Galerkin matrix of this point, inverted (uses the p / y area as scratch)
      coarse_setup<CS, CO, NM, NT>(pc, Ms, s_atoms, s_p, s_ei, s_cbuf, s_red, L::NSLOT * L::NREDV, red_flip);
      for (int i = t_id; i < NRHS * NDOF; i += NT) s_y[i] = 0.0;  // the accumulation target
    }
    sync();  // the preconditioner is complete before any group starts

    // ---- 3./4. one PCG per right-hand side, each group on its own named barrier ----
    // group-wide sum of one value per thread; buffers alternate so no trailing barrier is needed
    auto group_sum = [&](double v) {
      if (L::SUBW) return seg_sum(v, TPR);  // the right-hand side owns a segment of one warp
      v = warp_sum(v);
      double* buf = s_red + (red_flip ^= 1) * L::NSLOT * L::NREDV;
      if (lane == 0) buf[warp] = v;
      group_sync(1 + q, TPR);
      double s = 0.0;
      HMX_UNROLL
      for (int ww = 0; ww < WPR; ++ww) s += buf[q * WPR + ww];
      return s;
    };
    auto group_barrier = [&]() {
      if (L::SUBW)
        warp_sync();
      else
        group_sync(1 + q, TPR);
    };
    elasticity_sweep<CO, NM, NT, true, COLL, VGLOB>(pc, Ms, s_atoms, s_p, s_y, q, l, sqrtw, s_tab);  // y = b_q
    double rz, rz0;
    // z = M^-1 r at node slot i of this right-hand side (two-level: r sits in y_q, the level-1 coarse solution in
    // its class-0 slots, where the residual of those nodes is taken from the scratch instead)
    // OWN: thread l of a right-hand side owns level-1 node l and the 2^D fine nodes 2 c(l) + bits(j), j = its node
    // index (TPR = number of level-1 nodes, e.g. the 8^3 cell with 64 threads per right-hand side): the parents of
    // its nodes are c(l) and c(l) + bits(j) -- no table look-up, the first parent is loaded once
    constexpr bool OWN = TWO && TPR == CS::NC1 && NPT == (1 << D);
    int own_lo[D], own_hi[D];  // slot offsets of c(l) and of c(l) + 1 along every axis
    if (OWN) {
      int rr = l;
      HMX_UNROLL
      for (int a = 0; a < D; ++a) {
        const int ca = rr % CS::H;
        rr /= CS::H;
        own_lo[a] = ca * CS::stride1(a);
        own_hi[a] = ((ca + 1) % CS::H) * CS::stride1(a);
      }
    }
    auto precond_node = [&](int i, int j, double (&r)[D], double (&z)[D]) {
      HMX_UNROLL
      for (int c = 0; c < D; ++c) r[c] = (TWO && i >= CS::NC1) ? s_y[(q * D + c) * N + i] : g_r[(q * D + c) * N + i];
      int ia = 0, ib = 0;
      if (OWN) {
        ia = l;
        HMX_UNROLL
        for (int a = 0; a < D; ++a) ib += ((j >> a) & 1) ? own_hi[a] : own_lo[a];
      } else if (TWO) {
        const unsigned par = s_par[i];
        ia = (int)(par & 0xffffu);
        ib = (int)(par >> 16);
      }
      HMX_UNROLL
      for (int c = 0; c < D; ++c) {
        double zz = 0.0;
        HMX_UNROLL
        for (int c2 = 0; c2 < D; ++c2) zz += s_dinv[sym_index(D, c, c2) * N + i] * r[c2];
        if (TWO) zz += 0.5 * (s_y[(q * D + c) * N + ia] + s_y[(q * D + c) * N + ib]);
        z[c] = zz;
      }
    };
    if constexpr (TWO) {
      double splus[D];  // OWN: r[2c] + 1/2 sum of this thread's other nodes = its half of the level-1 restriction
      HMX_UNROLL
      for (int c = 0; c < D; ++c) splus[c] = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            const double rr = s_y[(q * D + c) * N + i];  // r = b stays in y for the restriction
            g_r[(q * D + c) * N + i] = rr;
            g_x[(q * D + c) * N + i] = 0.0;
            if (OWN) splus[c] += j == 0 ? rr : 0.5 * rr;
          }
        }
      }
      coarse_correct<CS, N, (L::o_cbuf % 2 == 0 && CS::NCD % 2 == 0), OWN>(s_y + (size_t)q * D * N, s_ei, s_cbuf + q * CS::NCD, q, l, splus);
      double part = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          double r[D], z[D];
          precond_node(i, j, r, z);
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            part += r[c] * z[c];
            s_p[(q * D + c) * N + i] = z[c];
          }
        }
      }
      if (!L::STORE1) {  // the sweep accumulates into y: clear it once nobody reads the coarse solution any more
        group_sync(1 + q, TPR);
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j) {
          const int i = l + j * TPR;
          if (i < N) {
            HMX_UNROLL
            for (int c = 0; c < D; ++c) s_y[(q * D + c) * N + i] = 0.0;
          }
        }
      }
      rz = rz0 = group_sum(part);  // its barrier also publishes p (and the cleared y)
    } else {
      double part = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          double r[D], z[D];
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            r[c] = s_y[(q * D + c) * N + i];
            s_y[(q * D + c) * N + i] = 0.0;
            g_r[(q * D + c) * N + i] = r[c];
            g_x[(q * D + c) * N + i] = 0.0;
          }
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            z[c] = 0.0;
            HMX_UNROLL
            for (int c2 = 0; c2 < D; ++c2) z[c] += s_dinv[sym_index(D, c, c2) * N + i] * r[c2];
            part += r[c] * z[c];
            s_p[(q * D + c) * N + i] = z[c];
          }
        }
      }
      rz = rz0 = group_sum(part);  // its barrier also publishes p and the zeroed y
      if (L::SUBW) warp_sync();    // (segmented shuffles carry no memory ordering)
    }
    int it = 0;
    bool active = rz0 > P.atol * P.atol;
    const double tol2 = fmax(P.rtol * P.rtol * rz0, P.atol * P.atol);
#if HMX_STAGGER > 0
    // The groups start in lock step and every iteration costs each of them the same, so without this they all sweep
    // (FP64 pipe saturated) and then all sit in their latency-bound phases (reductions, coarse correction: pipe idle)
    // at the same time.  Offsetting group q by q / NRHS of an iteration lets one group's bubbles fill with the
    // others' sweeps for the whole solve; costs < 1 % of a cell once.
    if (!L::SUBW) spin_cycles((long long)q * HMX_STAGGER);
#endif
    // SUBW: the right-hand sides sharing a warp loop together (the converged ones only keep y clean)
    while (L::SUBW ? warp_any(active && it < P.max_it) : (active && it < P.max_it)) {
      const bool mine = active && it < P.max_it;
      if (mine) ++it;
      if constexpr (L::BLK)  // y = K p
        elasticity_sweep_blocks<CO, NM, NT, COLL, VGLOB>(pc, Ms, s_atoms, s_p, s_y, q, l, sqrtw, s_tab);
      else
        elasticity_sweep<CO, NM, NT, false, COLL, VGLOB>(pc, Ms, s_atoms, s_p, s_y, q, l, sqrtw, s_tab);
      // (two-level) the old x and r come from the L2 scratch: issue those loads before the reduction of p.Kp, so
      // that their latency runs under the reduction and its barrier instead of after it
      constexpr int NPF = TWO ? NPT : 1;
      double r_old[NPF][D], x_old[NPF][D];
      if (TWO) {
        HMX_UNROLL
        for (int j = 0; j < NPF; ++j) {
          const int i = l + j * TPR;
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            r_old[j][c] = i < N ? g_r[(q * D + c) * N + i] : 0.0;
            x_old[j][c] = i < N ? g_x[(q * D + c) * N + i] : 0.0;
          }
        }
      }
      double part = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c) part += s_p[(q * D + c) * N + i] * s_y[(q * D + c) * N + i];
        }
      }
      const double pAp = group_sum(part);
      const double alpha = (mine && pAp > 0.0) ? fast_div(rz, pAp) : 0.0;
      if constexpr (TWO) {
        // x += alpha p, r -= alpha K p; r goes to the scratch and, for the restriction, into y
        double splus[D];
        HMX_UNROLL
        for (int c = 0; c < D; ++c) splus[c] = 0.0;
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j) {
          const int i = l + j * TPR;
          if (i < N) {
            HMX_UNROLL
            for (int c = 0; c < D; ++c) {
              const int a = (q * D + c) * N + i;
              const double rr = r_old[j < NPF ? j : 0][c] - alpha * s_y[a];
              g_x[a] = x_old[j < NPF ? j : 0][c] + alpha * s_p[a];
              g_r[a] = rr;
              s_y[a] = rr;
              if (OWN) splus[c] += j == 0 ? rr : 0.5 * rr;
            }
          }
        }
        coarse_correct<CS, N, (L::o_cbuf % 2 == 0 && CS::NCD % 2 == 0), OWN>(s_y + (size_t)q * D * N, s_ei, s_cbuf + q * CS::NCD, q, l, splus);
        // z = M^-1 r overwrites r in y (the class-0 slots hold the coarse solution other threads still read: the z
        // of those nodes -- at most NKEEP per thread -- waits in registers)
        constexpr int NKEEP = (CS::NC1 + TPR - 1) / TPR;
        double zk[NKEEP][D];
        part = 0.0;
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j) {
          const int i = l + j * TPR;
          if (i < N) {
            double r[D], z[D];
            precond_node(i, j, r, z);
            HMX_UNROLL
            for (int c = 0; c < D; ++c) {
              part += r[c] * z[c];
              if (j < NKEEP && i < CS::NC1)
                zk[j < NKEEP ? j : 0][c] = z[c];
              else
                s_y[(q * D + c) * N + i] = z[c];
            }
          }
        }
        const double rz_new = group_sum(part);
        const double beta = fast_div(rz_new, rz);
        rz = rz_new;
        if (!(rz_new > tol2)) active = false;
        if (active) {
          HMX_UNROLL
          for (int j = 0; j < NPT; ++j) {
            const int i = l + j * TPR;
            if (i < N) {
              HMX_UNROLL
              for (int c = 0; c < D; ++c) {
                const int a = (q * D + c) * N + i;
                const double z = (j < NKEEP && i < CS::NC1) ? zk[j < NKEEP ? j : 0][c] : s_y[a];
                s_p[a] = z + beta * s_p[a];
              }
            }
          }
          if (!L::STORE1) {  // clear y for the next sweep once nobody reads the coarse solution any more
            group_barrier();
            HMX_UNROLL
            for (int j = 0; j < NPT; ++j) {
              const int i = l + j * TPR;
              if (i < N) {
                HMX_UNROLL
                for (int c = 0; c < D; ++c) s_y[(q * D + c) * N + i] = 0.0;
              }
            }
          }
          group_barrier();  // publish p (and the cleared y) to the group
        }
        continue;
      }
      part = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          double r[D];
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            const int a = (q * D + c) * N + i;
            const double yv = s_y[a];
            if (!L::STORE1) s_y[a] = 0.0;
            r[c] = g_r[a];
            if (mine) {
              g_x[a] += alpha * s_p[a];
              r[c] -= alpha * yv;
              g_r[a] = r[c];
            }
          }
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            double z = 0.0;
            HMX_UNROLL
            for (int c2 = 0; c2 < D; ++c2) z += s_dinv[sym_index(D, c, c2) * N + i] * r[c2];
            part += r[c] * z;
          }
        }
      }
      const double rz_new = group_sum(part);
      const double beta = mine ? fast_div(rz_new, rz) : 0.0;
      if (mine) {
        rz = rz_new;
        if (!(rz_new > tol2)) active = false;
      }
      if (mine && active) {
        HMX_UNROLL
        for (int j = 0; j < NPT; ++j) {
          const int i = l + j * TPR;
          if (i < N) {
            double r[D];
            HMX_UNROLL
            for (int c = 0; c < D; ++c) r[c] = g_r[(q * D + c) * N + i];
            HMX_UNROLL
            for (int c = 0; c < D; ++c) {
              double z = 0.0;
              HMX_UNROLL
              for (int c2 = 0; c2 < D; ++c2) z += s_dinv[sym_index(D, c, c2) * N + i] * r[c2];
              const int a = (q * D + c) * N + i;
              s_p[a] = z + beta * s_p[a];
            }
          }
        }
      }
      if (L::SUBW || active) group_barrier();  // publish p (and the zeroed y) to the group
    }
    if (l == 0) {
      s_stat[4 * q + 0] = (double)it;
      s_stat[4 * q + 1] = rz;
      s_stat[4 * q + 2] = rz0;
    }

    // ---- 5. epilogue: b -> y again, A_hom = <C> - b_p.x_q - x_p.r_q ----
    if (L::STORE1 || TWO) {  // the loop left y = K p (two-level: the residual and the coarse solution) behind
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c) s_y[(q * D + c) * N + i] = 0.0;
        }
      }
      group_barrier();
    }
    // the search directions are dead: their shared memory takes the correctors x, so that the cross products
    // x_p . r_q below read every x_p from shared memory instead of the L2 scratch (the right-hand-side sweep does
    // not read p)
    constexpr bool XS = !VGLOB;
    if (XS) {
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c) s_p[(q * D + c) * N + i] = g_x[(q * D + c) * N + i];
        }
      }
    }
    const double* x_all = XS ? s_p : g_x;
    elasticity_sweep<CO, NM, NT, true, COLL, VGLOB>(pc, Ms, s_atoms, s_p, s_y, q, l, sqrtw, s_tab);
    sync();  // x, r, b of every right-hand side are visible to the whole CTA
    {
      double z[2 * NRHS];
      HMX_UNROLL
      for (int k = 0; k < 2 * NRHS; ++k) z[k] = 0.0;
      HMX_UNROLL
      for (int j = 0; j < NPT; ++j) {
        const int i = l + j * TPR;
        if (i < N) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            const double xq = x_all[(q * D + c) * N + i], rq = g_r[(q * D + c) * N + i];
            HMX_UNROLL
            for (int p = 0; p < NRHS; ++p) {
              z[p] += s_y[(p * D + c) * N + i] * xq;          // b_p . x_q
              z[NRHS + p] += x_all[(p * D + c) * N + i] * rq;  // x_p . r_q
            }
          }
        }
      }
      HMX_UNROLL
      for (int k = 0; k < 2 * NRHS; ++k) z[k] = L::SUBW ? seg_sum(z[k], TPR) : warp_sum(z[k]);
      // the groups did different numbers of reductions: after the barrier above every thread
      // restarts from the same buffer parity.  One slot per (right-hand side, warp of that side).
      red_flip = 0;
      double* buf = s_red;
      if (L::SUBW ? l == 0 : lane == 0) {
        const int slot = L::SUBW ? q : warp;  // non-SUBW: warp = q * WPR + (warp within the side)
        HMX_UNROLL
        for (int k = 0; k < 2 * NRHS; ++k) buf[slot * L::NREDV + k] = z[k];
      }
      sync();
      if (P.chi != nullptr) {  // correctors in natural node order: [q][component][node]
        constexpr int NN = G::N;
        for (int i = t_id; i < N; i += NT) {
          int c[3];
          if (!PG::decode(i, c)) continue;
          const int nat = G::index(c[0], c[1], c[2]);
          for (int k = 0; k < NRHS * D; ++k) P.chi[((size_t)pt * NRHS * D + k) * NN + nat] = g_x[k * N + i];
        }
      }
      // A_hom by thread 0, then one thread per entry of the macro element matrix (with one CTA per SM nothing would
      // hide a serial 12 x 12 x 36 epilogue)
      constexpr int NB = (D + 1) * D;
      double* s_ah = sm + L::o_epi;         // [NRHS][NRHS]
      double* s_cm = s_ah + NRHS * NRHS;    // [NV][NB] macro strain matrix, then |T|
      if (t_id == 0) {
        for (int qq = 0; qq < NRHS; ++qq) {
          double e[NV], sg[NV];
          HMX_UNROLL
          for (int v = 0; v < NV; ++v) e[v] = (v == qq) ? 1.0 : 0.0;
          CO::stress(pc, smean, e, sg);
          for (int p = 0; p < NRHS; ++p) {
            double z1 = 0.0, z2 = 0.0;
            for (int ww = 0; ww < WPR; ++ww) {
              z1 += buf[(qq * WPR + ww) * L::NREDV + p];
              z2 += buf[(qq * WPR + ww) * L::NREDV + NRHS + p];
            }
            s_ah[p * NRHS + qq] = sg[p] - z1 - z2;
          }
        }
        if (P.S_loc != nullptr) {
          double Cm[NV][NB];
          s_cm[NV * NB] = macro_strain_matrix<D, 1>(verts, Cm);
          HMX_UNROLL
          for (int p = 0; p < NV; ++p)
            HMX_UNROLL
            for (int i = 0; i < NB; ++i) s_cm[p * NB + i] = Cm[p][i];
        }
        int itmax = 0;
        unsigned long long tot = 0;
        double worst = 0.0;
        for (int qq = 0; qq < NRHS; ++qq) {
          const int iq = (int)s_stat[4 * qq];
          itmax = iq > itmax ? iq : itmax;
          tot += (unsigned long long)iq;
          if (s_stat[4 * qq + 2] > P.atol * P.atol) worst = fmax(worst, sqrt(s_stat[4 * qq + 1] / s_stat[4 * qq + 2]));
        }
        if (P.iters != nullptr) P.iters[pt] = itmax;
        if (P.resid != nullptr) P.resid[pt] = worst;
        if (P.work != nullptr) atomic_add_u64(P.work, tot);
      }
      sync();
      if (P.A_hom != nullptr)
        for (int k = t_id; k < NRHS * NRHS; k += NT) P.A_hom[pt * NRHS * NRHS + k] = s_ah[k];
      if (P.S_loc != nullptr) {
        for (int k = t_id; k < NB * NB; k += NT) {
          const int i = k / NB, j = k - i * NB;
          double acc = 0.0;  // same summation order as macro_element_matrix
          HMX_UNROLL
          for (int p = 0; p < NV; ++p)
            HMX_UNROLL
            for (int q2 = 0; q2 < NV; ++q2) acc += s_cm[p * NB + j] * s_ah[p * NRHS + q2] * s_cm[q2 * NB + i];
          P.S_loc[pt * NB * NB + k] = s_cm[NV * NB] * acc;
        }
      }
    }
    sync();  // shared memory and the scratch are reused by the next macro point
  }
}

}  // namespace hmx
