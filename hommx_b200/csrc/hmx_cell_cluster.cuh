// Micro cell kernel for 3-D linear elasticity, CLUSTER variant (HMX_VARIANT = 4): one thread-block CLUSTER per macro
// point, the periodic P1 stiffness matrix of the point ASSEMBLED once and RESIDENT in the distributed shared memory
// of the cluster for the whole PCG solve (BASELINE north star (2): "thread-block cluster + DSMEM for cells that
// exceed one SM").
//
// Replaces, like hmx_cell_elasticity.cuh, what the reference does per macro cell in BaseHMM._compute_local_stiffness
// (hmm.py:334-369) with the forms of hmm.py:887-922 / 1024-1067; the operator the reference re-assembles for every
// right-hand side (cell_problem.py:367-369) is assembled once per macro point here and applied ~100 x 6 times.
//
// Why a cluster: the block stencil (3 x 3 blocks on the 15-point Kuhn stencil) is 135 doubles per node, 553 KB for an
// 8^3 cell -- it fits in no single SM.  Stored as the 7 POSITIVE directions only (K_{i,i-d} = K_{i-d,i}^T) and with
// the diagonal blocks scaled away (below) it is 63 doubles per node: 129 KB per CTA when the cell is split into
// z-slabs over a cluster of 2 CTAs (8^3), 101 KB over 5 CTAs (10^3).  The matrix-free element kernel executes 372
// FP64 instructions per node, right-hand side and iteration; the resident stencil needs 126 FMA.
//
// Decomposition: CTA `rank` of the cluster owns the PZ = NM / CL node planes z = rank PZ .. rank PZ + PZ - 1.
//   * s_K  [7][9][NOWN]   scaled blocks K~_{i,i+d} of the own nodes, d = the 7 positive Kuhn directions;
//   * s_Kh [4][9][NM^2]   the blocks K~_{k,k+d}, d_z = 1, of the plane BELOW the slab (owned by the lower neighbour;
//                         assembled redundantly here so that no matrix entry is ever read through DSMEM);
//   * s_p  [(PZ+2) NM^2][18]  search directions of all 6 right-hand sides, node-major, own planes + one HALO plane
//                         below and above.  After every update one thread hands the two boundary planes (9 KB each,
//                         contiguous) to the copy engine: cp.async.bulk shared::cta -> shared::cluster into the
//                         neighbours' halo planes, completion counted as transaction bytes on the RECEIVER's mbarrier.
//                         The part of y = K p that needs no halo runs first, the four (of 14) block products per
//                         boundary node that do wait on that mbarrier: the SM-to-SM transfer hides behind the rest.
// Threads: TPN = 2 threads per node (lanes 2j, 2j+1 of a warp), each owning 3 of the 6 right-hand sides: its x~, r~
// (and y = K p while it is formed) live in REGISTERS; a matrix block is loaded once per thread pair and row (the
// two lanes read the same address: one shared-memory wavefront) and used for 27 FMA per thread.  The apply is a pure
// gather: no colouring, no atomics, fixed summation order.
//
// Scaled system: with K_ii = L_i L_i^T (3 x 3 Cholesky) the kernel iterates on K~ = L^-1 K L^-T, x~ = L^T x,
// r~ = L^-1 r: the diagonal blocks become identities (not stored, not multiplied), block Jacobi becomes the identity
// and PCG on K~ with M~^-1 = I + L^T P E^-1 P^T L is, in exact arithmetic, the PCG of the matrix-free kernel with
// M^-1 = blockdiag(K)^-1 + P E^-1 P^T (hmx_cell_coarse.cuh) iterate by iterate.  Every thread keeps the L of its node.
//
// Cluster-wide reductions (p.Kp; r~.r~ together with the restricted residual P^T r): warp shuffles -> CTA partial
// -> every CTA stores its partial into every CTA's receive buffer (st.async: DSMEM stores counted as transaction
// bytes on the RECEIVER's mbarrier) -> every CTA waits on its own mbarrier and adds the partials in rank order: all
// CTAs hold bit-identical scalars and take identical branches.  No barrier.cluster inside the iteration (its
// release is a MEMBAR.ALL.GPU, its acquire a CCTL.IVALL): the two exchanges alternate, and a CTA can only send round
// k + 1 of one after it has received round k of the other from every peer, which also orders the reuse of the halo
// planes and receive buffers.  r.z needs no second reduction: r~.z~ = r~.r~ + (P^T r).E^-1 (P^T r).
// Coarse space: the semi-coarsened space of hmx_cell_coarse.cuh (level-1 Kuhn P1 summed along the first micro axis
// when the coefficient does not depend on it), set up by the same coarse_setup in every CTA; restriction = line sums
// along x (warp shuffles) followed by the 2-D Kuhn restriction of the lines a CTA owns.
#pragma once
#include "hmx_cell_common.cuh"
#include "hmx_cell_coarse.cuh"

namespace hmx {

template <class CO, int NM, int NT, int CL, int TPN = 2>
struct ClusterLayout {
  static constexpr int D = CO::DIM;
  static_assert(D == 3, "the cluster variant is the 3-D elasticity kernel");
  static constexpr int T = kuhn_ntypes<D>();
  static constexpr int NRHS = D * (D + 1) / 2;
  static constexpr int NV = NRHS;
  static constexpr int NVEC = NRHS * D;    // values per node: all right-hand sides x components
  static constexpr int NRL = NRHS / TPN;   // right-hand sides per thread
  static constexpr int NVL = NRL * D;      // values per thread
  static constexpr int NH = (1 << D) - 1;  // positive stencil directions
  static constexpr int NB = D * D;
  static constexpr int PZ = NM / CL;       // node planes per CTA
  static constexpr int NPL = NM * NM;      // nodes per plane
  static constexpr int NOWN = PZ * NPL;
  static constexpr int NPB = (PZ + 2) * NPL;  // node slots of the p buffer: halo plane, own planes, halo plane
  static constexpr int NW = NT / 32;
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  using AI = AtomIdx<D, NM, CO::YDEP, true>;
  static constexpr int NRC = AI::NRC;
  using CS = CoarseSpace<CO, NM, NT, 0, 0, 0>;
  // two-level preconditioner: the semi-coarsened space along axis 0 (the z-slabs must not cut the summed axis and the
  // line sums run along the lanes of a warp)
  // line sums by shuffles when a line of NM nodes x TPN lanes is an aligned power-of-two segment of a warp
  static constexpr bool LSHFL = (TPN * NM) <= 32 && ((TPN * NM) & (TPN * NM - 1)) == 0;
  // ... which is also a condition of the two-level method: with the generic line sums (NM barrier steps) and a 75-unknown
  // coarse solve the 10^3 cell needs 152 instead of 379 iterations but pays 21.8 instead of 8.7 us for each -- 6.8k
  // against 7.3k cell solves/s (measured; -DHMX_CLUSTER_TWO_ANY=1 builds it anyway: tests, experiments)
#ifndef HMX_CLUSTER_TWO_ANY
#define HMX_CLUSTER_TWO_ANY 0
#endif
  static constexpr bool TWO = CS::GEOM && CS::SEMI && CS::SA == 0 && CS::NCD <= CS::MAXDOF && (LSHFL || HMX_CLUSTER_TWO_ANY);
  // the scratch of coarse_setup lives in the matrix area of the CTA when it fits (8^3), else in its global scratch
  // (10^3: 243 KB; set-up only)
  static constexpr bool SUG = TWO && CS::setup_doubles > ((1 << D) - 1) * D * D * (NM / CL) * NM * NM + 4 * D * D * NM * NM;
  static constexpr int NCD = TWO ? CS::NCD : 2;
  static constexpr int NTRI = TWO ? CS::NTRI : 0;
  static constexpr int NC2 = TWO ? CS::NC2 : 1;
  static constexpr int H = NM / 2;
  static constexpr int NBLK = (NCD + 31) / 32;
  static constexpr int NLINE = PZ * NM;  // x-lines this CTA owns
  static constexpr int NREC = ((NRHS * NCD + NRHS > NRHS * NRHS ? NRHS * NCD + NRHS : NRHS * NRHS) + 1) / 2 * 2;  // doubles a CTA sends per exchange
  static constexpr int EPART = NT / (NRHS * NRHS);  // node partitions of the epilogue's cross products
  HMX_HOSTDEV static constexpr int imax(int a, int b) { return a > b ? a : b; }
  // ---- shared memory (doubles) ----
  static constexpr int o_bar = 0;                                    // 3 mbarriers: halo planes, p.Kp exchange, residual exchange
  static constexpr int o_red = o_bar + 3 * HMX_MBAR_BYTES / 8;       // [NW][8] warp partials
  static constexpr int o_xch = o_red + NW * 8;                       // [CL][8] p.Kp partials of every CTA
  static constexpr int o_recv = o_xch + CL * 8;                      // [CL][NREC] restricted residual + r~.r~ partials
  static constexpr int o_scal = o_recv + CL * NREC;                  // [8] r~.r~ of every right-hand side
  static constexpr int o_u = o_scal + 8;                             // [NRHS][NCD] coarse solution
  // work area: line sums [NLINE][NVEC] -> then (after the exchange) the summed restricted residual [NRHS][NCD] and
  // the partial dot products [NRHS][NBLK]; also the column buffers of coarse_setup and the epilogue's partials
  static constexpr int WORK = imax(imax(NLINE * NVEC, NRHS * NCD + NRHS * NBLK + 2),
                                   imax(imax(TWO ? CS::CBUF : 0, EPART * NRHS * NRHS), NRHS * NRHS + NV * (D + 1) * D + 2));
  static constexpr int o_work = o_u + NRHS * NCD;
  static constexpr int o_ei = ((o_work + WORK + 1) / 2) * 2;         // [NTRI] inverse coarse matrix, packed
  // ... which moves to the global scratch (read through L1 by the coarse solve) when it is what pushes the CTA over
  // 227 KB (10^3 on 5 CTAs: 22.8 KB)
  static constexpr bool EIG = TWO && (o_ei + NTRI + 2 + NPB * NVEC + NH * NB * NOWN + 4 * NB * NPL) * 8 > 232448;
  static constexpr int o_p = ((o_ei + (EIG ? 0 : NTRI) + 1) / 2) * 2;  // [NPB][NVEC]
  static constexpr int o_K = o_p + NPB * NVEC;                       // [NH][NB][NOWN]
  static constexpr int o_Kh = o_K + NH * NB * NOWN;                  // [4][NB][NPL]
  static constexpr int total = o_Kh + 4 * NB * NPL;
  // set-up aliases: atoms [NA1][T][NRC] and the inverse Cholesky factors [6][NPB] live in the p area, the scratch
  // of coarse_setup in the matrix area (both dead before the first search direction / matrix block is written)
  // (coefficients with many atoms on all three axes: the atoms go to this CTA's global scratch instead -- set-up only)
  static constexpr bool ATG = NA1 * T * NRC + 6 * NPB + 6 * NOWN > NPB * NVEC;
  static constexpr int o_atoms = o_p;
  static constexpr int o_li = o_p + (ATG ? 0 : NA1 * T * NRC);  // [6][NPB] inverse factors (own + halo planes)
  static constexpr int o_lf = o_li + 6 * NPB;                   // [6][NOWN] factors of the own nodes
  static_assert(6 * NPB + 6 * NOWN <= NPB * NVEC, "Cholesky factors must fit in the p area during set-up");
  // diagonal-block partials of the two halves of the simplex types, [2][27][NOWN] own + [2][9][2 NPL] halo: matrix area
  static_assert(2 * 27 * NOWN + 2 * 9 * 2 * NPL <= NH * NB * NOWN + 4 * NB * NPL, "diagonal partials must fit in the matrix area");
  static constexpr int PLANE_BYTES = NPL * NVEC * 8;
  static_assert(PLANE_BYTES % 16 == 0 && (o_p * 8) % 16 == 0 && HMX_MBAR_BYTES % 8 == 0, "bulk copies of whole node planes");
  static_assert(2 * NOWN * NVEC <= NH * NB * NOWN + 4 * NB * NPL, "epilogue copies of r~ and b~ must fit in the matrix area");
  // per-CTA global scratch: b~ of the own nodes (+ atoms, + inverse coarse matrix, + coarse set-up scratch)
  static constexpr int g_atoms = NOWN * NVEC;
  static constexpr int g_ei = g_atoms + (ATG ? NA1 * T * NRC : 0);
  static constexpr int g_setup = ((g_ei + (EIG ? NTRI : 0) + 1) / 2) * 2;
  static constexpr int scratch_doubles = g_setup + (SUG ? CS::setup_doubles : 0);
  static_assert(NM % CL == 0, "the cluster splits the cell into slabs of whole node planes");
  static_assert(TPN >= 1 && TPN <= 2 && NRHS % TPN == 0 && NT % 32 == 0 && NT >= NOWN * TPN && 32 % TPN == 0, "TPN threads per own node");
  static_assert(NT >= NRHS * NCD + NRHS && NT >= NRHS * NRHS, "one thread per coarse unknown and right-hand side in the exchanges");
  static_assert(total * 8 <= 232448, "the cluster kernel's shared memory exceeds 227 KB per CTA: use a larger cluster");
};

// engineering-Voigt strain of the vector basis function phi e_j whose mapped gradient is m
HMX_DEV void cl_basis_strain(const double (&m)[3], int j, double (&e)[6]) {
  HMX_UNROLL
  for (int v = 0; v < 3; ++v) e[v] = (v == j) ? m[v] : 0.0;
  int v = 3;
  HMX_UNROLL
  for (int r = 0; r < 3; ++r)
    HMX_UNROLL
    for (int c = r + 1; c < 3; ++c) {
      e[v] = ((c == j) ? m[r] : 0.0) + ((r == j) ? m[c] : 0.0);
      ++v;
    }
}

// vertex b >= a of a type-t simplex with P(t, b) - P(t, a) = dir (dir = 0: b = a), or -1
HMX_HOSTDEV constexpr int cl_bsel(int t, int a, int dir) {
  int r = -1;
  for (int b = a; b <= 3; ++b)
    if ((kuhn_pmask<3>(t, b) & ~kuhn_pmask<3>(t, a)) == dir) r = b;
  return r;
}

// blk += the contributions of the simplex types 3 PART .. 3 PART + 2 to the block K_{i, i+DIR} of the periodic stiffness
// matrix, i = the node with (global, periodic) coordinates c:  DIR = 0 the diagonal block, DIR = 1..7 the positive Kuhn
// direction with that bit mask.  blk[r * 3 + s] couples component r at node i with component s at node i + DIR.
// With RHS the load vectors of all right-hand sides ride along with the diagonal block:
//   rhs[q * 3 + j] += their share of  b_q[i][j] = -|e| sum_elements (C E_q) : e(phi_i e_j)   (hmm.py:898-903).
// Everything about the mesh is folded at compile time (types, vertices, directions); one copy of the code per
// (DIR, PART), called from the own-node and the halo-node loops.
template <class CO, int NM, int DIR, int PART, bool RHS>
HMX_DEV_NOINLINE void cl_block(const int* c_in, const double* pc, const double* Mn, const double* s_atoms, double vol, double* blk,
                               double* rhs) {
  using G = Grid<3, NM, 0>;
  using AI = AtomIdx<3, NM, CO::YDEP, true>;
  constexpr int D = 3, T = 6, NV = 6, NA = CO::NATOMS, NA1 = NA > 0 ? NA : 1, NRC = AI::NRC;
  const int c[3] = {c_in[0], c_in[1], c_in[2]};
  HMX_UNROLL
  for (int tt = 0; tt < 3; ++tt) {
    const int t = 3 * PART + tt;
    HMX_UNROLL
    for (int a = 0; a <= D; ++a) {
      // the type-t simplex in which node i is vertex a: its edge a -> b in direction DIR, if it has one
      const int b = cl_bsel(t, a, DIR);
      if (b < 0) continue;
      int o[3];
      G::template shift_coords<-1>(c, kuhn_pmask<D>(t, a), o);
      const int ro = AI::ridx(o);
      double sa[NA1];
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) sa[k] = NA > 0 ? s_atoms[(k * T + t) * NRC + ro] : 0.0;
      double ga[D], gb[D];  // M-transformed gradients of the vertex functions a and b
      HMX_UNROLL
      for (int p = 0; p < D; ++p) {
        ga[p] = 0.0;
        if (a >= 1) ga[p] += Mn[p * D + kuhn_axis<D>(t, a >= 1 ? a - 1 : 0)];
        if (a < D) ga[p] -= Mn[p * D + kuhn_axis<D>(t, a < D ? a : 0)];
        gb[p] = 0.0;
        if (b >= 1) gb[p] += Mn[p * D + kuhn_axis<D>(t, b >= 1 ? b - 1 : 0)];
        if (b < D) gb[p] -= Mn[p * D + kuhn_axis<D>(t, b < D ? b : 0)];
      }
      double eb[D][NV];
      HMX_UNROLL
      for (int s = 0; s < D; ++s) cl_basis_strain(gb, s, eb[s]);
      HMX_UNROLL
      for (int r = 0; r < D; ++r) {
        double ea[NV], sg[NV];
        cl_basis_strain(ga, r, ea);
        CO::stress(pc, sa, ea, sg);
        if (RHS) {
          HMX_UNROLL
          for (int q = 0; q < NV; ++q) rhs[q * D + r] -= vol * sg[q];  // (C ea)[q] = ea : C : E_q
        }
        HMX_UNROLL
        for (int s = 0; s < D; ++s) {
          double v = 0.0;
          HMX_UNROLL
          for (int k = 0; k < NV; ++k) v = fma(sg[k], eb[s][k], v);
          blk[r * D + s] = fma(vol, v, blk[r * D + s]);
        }
      }
    }
  }
}

// Cholesky factor of a symmetric positive definite 3 x 3 block (row-major full storage, lower part used) and its
// inverse, both lower triangular, packed {00, 10, 11, 20, 21, 22}
HMX_DEV void cl_chol3(const double (&a)[9], double (&l)[6], double (&li)[6]) {
  const double i0 = fast_rsqrt(a[0]);
  l[0] = a[0] * i0;
  l[1] = a[3] * i0;
  l[3] = a[6] * i0;
  const double d1 = a[4] - l[1] * l[1];
  const double i1 = fast_rsqrt(d1);
  l[2] = d1 * i1;
  l[4] = (a[7] - l[3] * l[1]) * i1;
  const double d2 = a[8] - l[3] * l[3] - l[4] * l[4];
  const double i2 = fast_rsqrt(d2);
  l[5] = d2 * i2;
  li[0] = i0;
  li[2] = i1;
  li[5] = i2;
  li[1] = -l[1] * i0 * i1;
  li[4] = -l[4] * i1 * i2;
  li[3] = -(l[3] * li[0] + l[4] * li[1]) * i2;
}
// entry (r, s) of a packed lower-triangular 3 x 3 matrix, r >= s
HMX_HOSTDEV constexpr int cl_tri(int r, int s) { return r * (r + 1) / 2 + s; }

template <class CO, int NM, int NT, int CL, int TPN = 2>
HMX_DEV void elasticity_cluster_cell_body(const CellParams& P) {
  using L = ClusterLayout<CO, NM, NT, CL, TPN>;
  using AI = typename L::AI;
  using CS = typename L::CS;
  constexpr int D = 3, T = 6, NRHS = 6, NV = 6, NVEC = L::NVEC, NRL = L::NRL, NVL = L::NVL, NH = 7, NB = 9;
  constexpr int PZ = L::PZ, NPL = L::NPL, NOWN = L::NOWN, NPB = L::NPB, NW = L::NW, NA = L::NA, NA1 = L::NA1, NRC = L::NRC;
  constexpr int NCD = L::NCD, NBLK = L::NBLK, NREC = L::NREC, H = L::H;
  constexpr bool TWO = L::TWO;
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;

  double* sm = dyn_smem();
  MBar* s_bar = reinterpret_cast<MBar*>(sm + L::o_bar);  // halo planes (bulk copies of the neighbours)
  MBar* s_bar_x = reinterpret_cast<MBar*>(sm + L::o_bar + HMX_MBAR_BYTES / 8);      // p.Kp partials of all CTAs
  MBar* s_bar_c = reinterpret_cast<MBar*>(sm + L::o_bar + 2 * HMX_MBAR_BYTES / 8);  // restricted residual + r~.r~ partials
  double* s_red = sm + L::o_red;
  double* s_xch = sm + L::o_xch;
  double* s_recv = sm + L::o_recv;
  double* s_scal = sm + L::o_scal;
  double* s_u = sm + L::o_u;
  double* s_work = sm + L::o_work;
  double* g_b = P.scratch + (size_t)bid() * L::scratch_doubles;
  double* s_ei = L::EIG ? g_b + L::g_ei : sm + L::o_ei;
  double* coarse_work = L::SUG ? g_b + L::g_setup : sm + L::o_K;
  double* s_p = sm + L::o_p;
  double* s_K = sm + L::o_K;
  double* s_Kh = sm + L::o_Kh;
  double* s_atoms = L::ATG ? g_b + L::g_atoms : sm + L::o_atoms;
  double* s_li = sm + L::o_li;  // [6][NPB] inverse Cholesky factors of the diagonal blocks (set-up only)
  double* s_lf = sm + L::o_lf;  // [6][NOWN] Cholesky factors of the own nodes (set-up only)

  const int t_id = tid(), lane = t_id & 31, warp = t_id >> 5;
  const int rank = cluster_rank();
  const int jn = t_id / TPN, h = t_id - jn * TPN;  // own node slot, half
  const bool own = jn < NOWN;
  const int j = own ? jn : 0;
  const int cx = j % NM, cy = (j / NM) % NM, zl = j / NPL;
  const int zg = rank * PZ + zl;
  // neighbour offsets inside a plane (periodic wrap), in node slots
  const int oxp = cx + 1 < NM ? 1 : 1 - NM, oxm = cx > 0 ? -1 : NM - 1;
  const int oyp = cy + 1 < NM ? NM : NM - NPL, oym = cy > 0 ? -NM : NPL - NM;
  const int pb = j + NPL;  // own slot in the p buffer
  const int r_lo = (rank + CL - 1) % CL, r_up = (rank + 1) % CL;
  const double hh = 1.0 / (double)NM;
  const double vol = hh * hh * hh / 6.0;
  const double sqrtw = sqrt(vol);
  int red_flip = 0;
  unsigned halo_parity = 0, x_parity = 0, c_parity = 0;
  // offset of direction mask m (bits x, y, z) from node slot s of the p buffer, forwards / backwards
  auto fwd = [&](int m) { return ((m & 1) ? oxp : 0) + ((m & 2) ? oyp : 0) + ((m & 4) ? NPL : 0); };
  auto bwd = [&](int m) { return ((m & 1) ? oxm : 0) + ((m & 2) ? oym : 0) - ((m & 4) ? NPL : 0); };

  if (t_id == 0) {
    mbar_init(s_bar, 1);
    mbar_init(s_bar_x, 1);
    mbar_init(s_bar_c, 1);
    mbar_fence_init();
  }
  cluster_sync();  // every CTA's mbarrier exists before a peer's copy can signal it

  for (long long pt = cluster_id(); pt < P.n_pts; pt += nclusters()) {
    double xm[3], verts[(D + 1) * 3];
    macro_point<D>(P, pt, xm, verts);
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double Mn[D * D], Ms[D * D];
    CO::dtheta(xm, Mn);
    HMX_UNROLL
    for (int k = 0; k < D * D; ++k) {
      Mn[k] *= (double)NM;
      Ms[k] = Mn[k] * sqrtw;
    }

    // ---- 1. atoms (every CTA computes all of them: its own slab's elements and the coarse space need them) ----
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
          for (int a = 0; a < D; ++a) y[a] = ((double)c[a] + P.qp[(t * P.nq + qq) * D + a]) * hh;
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
      for (int idx = t_id; idx < T * NRC; idx += NT) {
        HMX_UNROLL
        for (int k = 0; k < NA; ++k) smean[k] += s_atoms[k * T * NRC + idx];
      }
      block_sum<NA1, NW>(smean, s_red);
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) smean[k] *= 1.0 / (double)(T * ipow(NM, AI::NDEP));
      sync();  // s_red is reused below
    }

    // ---- 2. coarse matrix of this point, inverted (scratch: the matrix area) ----
    if constexpr (TWO) coarse_setup<CS, CO, NM, NT>(pc, Ms, s_atoms, coarse_work, s_ei, s_work, s_red, NW * 8 / 2, red_flip);

    // ---- 3. diagonal blocks: the two halves of the simplex types in parallel (warp-uniform when NOWN is a multiple
    //         of 32), partial sums through the matrix area ----
    {
      double* s_dp = s_K;                    // [2][27][NOWN] partial diagonal blocks + load vectors of the own nodes
      double* s_dh = s_K + 2 * 27 * NOWN;    // [2][9][2 NPL] partial diagonal blocks of the two halo planes
      for (int task = t_id; task < 2 * (NOWN + 2 * NPL); task += NT) {
        const bool halo = task >= 2 * NOWN;
        const int u = halo ? task - 2 * NOWN : task, cnt = halo ? 2 * NPL : NOWN;
        const int part = u / cnt, k = u - part * cnt;
        int c[3];
        if (halo) {
          const int up = k / NPL, kk = k - up * NPL;
          c[0] = kk % NM;
          c[1] = kk / NM;
          c[2] = (rank * PZ + (up ? PZ : NM - 1)) % NM;
        } else {
          c[0] = k % NM;
          c[1] = (k / NM) % NM;
          c[2] = rank * PZ + k / NPL;
        }
        double blk[9], rhs[18];
        HMX_UNROLL
        for (int e = 0; e < 9; ++e) blk[e] = 0.0;
        HMX_UNROLL
        for (int e = 0; e < 18; ++e) rhs[e] = 0.0;
        if (halo) {
          if (part == 0)
            cl_block<CO, NM, 0, 0, false>(c, pc, Mn, s_atoms, vol, blk, rhs);
          else
            cl_block<CO, NM, 0, 1, false>(c, pc, Mn, s_atoms, vol, blk, rhs);
          HMX_UNROLL
          for (int e = 0; e < 9; ++e) s_dh[(part * 9 + e) * 2 * NPL + k] = blk[e];
        } else {
          if (part == 0)
            cl_block<CO, NM, 0, 0, true>(c, pc, Mn, s_atoms, vol, blk, rhs);
          else
            cl_block<CO, NM, 0, 1, true>(c, pc, Mn, s_atoms, vol, blk, rhs);
          HMX_UNROLL
          for (int e = 0; e < 9; ++e) s_dp[(part * 27 + e) * NOWN + k] = blk[e];
          HMX_UNROLL
          for (int e = 0; e < 18; ++e) s_dp[(part * 27 + 9 + e) * NOWN + k] = rhs[e];
        }
      }
      sync();
      // Cholesky factors; b~ = L^-1 b of every right-hand side -> scratch
      for (int task = t_id; task < NOWN + 2 * NPL; task += NT) {
        const bool halo = task >= NOWN;
        const int k = halo ? task - NOWN : task;
        double blk[9], lf[6], li[6];
        HMX_UNROLL
        for (int e = 0; e < 9; ++e)
          blk[e] = halo ? s_dh[e * 2 * NPL + k] + s_dh[(9 + e) * 2 * NPL + k] : s_dp[e * NOWN + k] + s_dp[(27 + e) * NOWN + k];
        cl_chol3(blk, lf, li);
        const int slot = halo ? (k < NPL ? k : (PZ + 1) * NPL + (k - NPL)) : NPL + k;
        HMX_UNROLL
        for (int e = 0; e < 6; ++e) s_li[e * NPB + slot] = li[e];
        if (!halo) {
          HMX_UNROLL
          for (int e = 0; e < 6; ++e) s_lf[e * NOWN + k] = lf[e];
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            double b[D];
            HMX_UNROLL
            for (int r = 0; r < D; ++r) b[r] = s_dp[(9 + q * D + r) * NOWN + k] + s_dp[(27 + 9 + q * D + r) * NOWN + k];
            HMX_UNROLL
            for (int r = 0; r < D; ++r) {
              double v = 0.0;
              HMX_UNROLL
              for (int s = 0; s <= r; ++s) v += li[cl_tri(r, s)] * b[s];
              g_b[(size_t)k * NVEC + q * D + r] = v;
            }
          }
        }
      }
      sync();
    }

    // ---- 4. scaled off-diagonal blocks  K~_{i,i+d} = L_i^-1 K_{i,i+d} L_{i+d}^-T ----
    // dst[(r * 3 + s) * stride] (=|+=) the scaled block
    auto scale_store = [&](const double (&blk)[9], int slot, int nb_slot, double* dst, int stride, bool add) {
      double li[6], lj[6];
      HMX_UNROLL
      for (int e = 0; e < 6; ++e) {
        li[e] = s_li[e * NPB + slot];
        lj[e] = s_li[e * NPB + nb_slot];
      }
      double t1[9];
      HMX_UNROLL
      for (int r = 0; r < D; ++r)
        HMX_UNROLL
        for (int s = 0; s < D; ++s) {
          double v = 0.0;
          HMX_UNROLL
          for (int k = 0; k <= r; ++k) v += li[cl_tri(r, k)] * blk[k * D + s];
          t1[r * D + s] = v;
        }
      HMX_UNROLL
      for (int r = 0; r < D; ++r)
        HMX_UNROLL
        for (int s = 0; s < D; ++s) {
          double v = 0.0;
          HMX_UNROLL
          for (int k = 0; k <= s; ++k) v += t1[r * D + k] * lj[cl_tri(s, k)];
          if (add)
            dst[(r * D + s) * stride] += v;
          else
            dst[(r * D + s) * stride] = v;
        }
    };
    {
      // own nodes: thread (node k, half `part` of the simplex types); part 0 stores, part 1 adds after a barrier
      constexpr bool PAR = NT >= 2 * NOWN;
      const int part = PAR ? t_id / NOWN : 0, k = PAR ? t_id - part * NOWN : t_id;
      const bool on = PAR ? part < 2 : k < NOWN;
      const int kx = k % NM, ky = (k / NM) % NM, kz = k / NPL;
      const int c[3] = {kx, ky, rank * PZ + kz};
      double rhs[1] = {0.0};
      auto block_dir = [&](auto dtag) {
        constexpr int DIR = decltype(dtag)::value;
        double blk[9];
        HMX_UNROLL
        for (int e = 0; e < 9; ++e) blk[e] = 0.0;
        if (on) {
          if (!PAR || part == 0) cl_block<CO, NM, DIR, 0, false>(c, pc, Mn, s_atoms, vol, blk, rhs);
          if (!PAR || part == 1) cl_block<CO, NM, DIR, 1, false>(c, pc, Mn, s_atoms, vol, blk, rhs);
        }
        const int nb = NPL + ((DIR & 1) ? (kx + 1) % NM : kx) + NM * ((DIR & 2) ? (ky + 1) % NM : ky) + NPL * (kz + ((DIR & 4) ? 1 : 0));
        double* dst = s_K + (size_t)(DIR - 1) * NB * NOWN + k;
        if (on && part == 0) scale_store(blk, NPL + k, nb, dst, NOWN, false);
        if (PAR) {
          sync();
          if (on && part == 1) scale_store(blk, NPL + k, nb, dst, NOWN, true);
        }
      };
      block_dir(IntTag<1>{});
      block_dir(IntTag<2>{});
      block_dir(IntTag<3>{});
      block_dir(IntTag<4>{});
      block_dir(IntTag<5>{});
      block_dir(IntTag<6>{});
      block_dir(IntTag<7>{});
      // the plane below: its blocks towards this slab (directions with a z step)
      for (int idx = t_id; idx < 4 * NPL; idx += NT) {
        const int dd = idx / NPL, kk = idx - dd * NPL;
        const int hx = kk % NM, hy = kk / NM;
        const int ch[3] = {hx, hy, (rank * PZ + NM - 1) % NM};
        double blk[9];
        HMX_UNROLL
        for (int e = 0; e < 9; ++e) blk[e] = 0.0;
        if (dd == 0) {
          cl_block<CO, NM, 4, 0, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
          cl_block<CO, NM, 4, 1, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
        } else if (dd == 1) {
          cl_block<CO, NM, 5, 0, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
          cl_block<CO, NM, 5, 1, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
        } else if (dd == 2) {
          cl_block<CO, NM, 6, 0, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
          cl_block<CO, NM, 6, 1, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
        } else {
          cl_block<CO, NM, 7, 0, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
          cl_block<CO, NM, 7, 1, false>(ch, pc, Mn, s_atoms, vol, blk, rhs);
        }
        const int d = 4 + dd;
        const int nb = NPL + ((d & 1) ? (hx + 1) % NM : hx) + NM * ((d & 2) ? (hy + 1) % NM : hy);
        scale_store(blk, kk, nb, s_Kh + (size_t)dd * NB * NPL + kk, NPL, false);
      }
    }
    // this thread's Cholesky factor and load vectors
    double Lf[6];
    HMX_UNROLL
    for (int e = 0; e < 6; ++e) Lf[e] = s_lf[e * NOWN + j];
    sync();  // the matrix is complete; atoms and factors (p area) are dead

    // ---- 5. PCG on the scaled system, all right-hand sides in lock step ----
    double xt[NVL], rt[NVL];
    HMX_UNROLL
    for (int k = 0; k < NVL; ++k) {
      xt[k] = 0.0;
      rt[k] = own ? g_b[(size_t)j * NVEC + h * NVL + k] : 0.0;
    }
    // sum of one value per right-hand side of this thread over the CTA -> s_red[warp][h * NRL + q]
    auto warp_partials = [&](const double (&v)[NRL]) {
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        double s = v[q];
        for (int m = TPN; m < 32; m <<= 1) s += lane_xor(s, m);
        if (lane < TPN) s_red[warp * 8 + lane * NRL + q] = s;
      }
    };
    // z~ = M~^-1 r~ and r~.z~ per right-hand side.  Contains one cluster barrier; every thread of the cluster calls it.
    auto precond = [&](const double (&r)[NVL], double (&z)[NVL], double (&rz)[NRL]) {
      double rr[NRL];
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        rr[q] = 0.0;
        HMX_UNROLL
        for (int c = 0; c < D; ++c) rr[q] += r[q * D + c] * r[q * D + c];
      }
      warp_partials(rr);
      if constexpr (TWO) {
        // physical residual L r~, summed along the x-lines
        double rp[NVL];
        HMX_UNROLL
        for (int q = 0; q < NRL; ++q)
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            double v = 0.0;
            HMX_UNROLL
            for (int s = 0; s <= c; ++s) v += Lf[cl_tri(c, s)] * r[q * D + s];
            rp[q * D + c] = own ? v : 0.0;
          }
        if constexpr (L::LSHFL) {
          HMX_UNROLL
          for (int k = 0; k < NVL; ++k) {
            double s = rp[k];
            for (int m = TPN; m < TPN * NM; m <<= 1) s += lane_xor(s, m);
            if (own && cx == 0) s_work[(cy + NM * zl) * NVEC + h * NVL + k] = s;
          }
        } else {
          for (int xx = 0; xx < NM; ++xx) {  // fixed order along the line
            if (own && cx == xx) {
              HMX_UNROLL
              for (int k = 0; k < NVL; ++k) {
                double* dst = s_work + (cy + NM * zl) * NVEC + h * NVL + k;
                *dst = xx == 0 ? rp[k] : *dst + rp[k];
              }
            }
            sync();
          }
        }
      }
      sync();
      // CTA partials -> every CTA's receive buffer
      {
        double v = 0.0;
        bool send = false;
        int slot = 0;
        if constexpr (TWO) {
          if (t_id < NRHS * NCD) {
            // restricted residual of unknown (level-2 node C = Y + H Z, component c) for right-hand side q: the 2-D
            // Kuhn restriction of the line sums, restricted to the lines this CTA owns
            const int q = t_id / NCD, i = t_id - q * NCD, C = i / D, c = i - C * D;
            const int Y = C % H, Z = C / H;
            HMX_UNROLL
            for (int e = 0; e < 7; ++e) {
              constexpr int dy[7] = {0, 1, -1, 0, 0, 1, -1}, dz[7] = {0, 0, 0, 1, -1, 1, -1};
              const int yy = (2 * Y + dy[e] + NM) % NM, zz = (2 * Z + dz[e] + NM) % NM;
              const int zloc = zz - rank * PZ;
              if (zloc >= 0 && zloc < PZ) v += (e == 0 ? 1.0 : 0.5) * s_work[(yy + NM * zloc) * NVEC + q * D + c];
            }
            send = true;
            slot = t_id;
          }
        }
        if (t_id >= NT - NRHS) {  // the last NRHS threads: r~.r~ partial of right-hand side q
          const int q = t_id - (NT - NRHS);
          v = 0.0;
          for (int w = 0; w < NW; ++w) v += s_red[w * 8 + q];
          send = true;
          slot = NRHS * NCD + q;
        }
        // every value goes to every CTA (this one included) as a store counted on the receiver's mbarrier: the wait
        // below needs no cluster barrier (whose release is a MEMBAR.ALL.GPU and whose acquire invalidates L1).  One
        // buffer per exchange suffices: a CTA sends round k + 1 only after it has received round k of the OTHER
        // exchange from every peer, which the peer sent after it had read this buffer's round k.
        if (send) {
          HMX_UNROLL
          for (int rk = 0; rk < CL; ++rk) st_async_f64(s_recv + rank * NREC + slot, v, s_bar_c, rk);
        }
        if (t_id == 0) mbar_arrive_expect_tx(s_bar_c, (unsigned)(CL * ((TWO ? NRHS * NCD : 0) + NRHS) * 8));
      }
      mbar_wait(s_bar_c, c_parity);
      c_parity ^= 1u;
      // totals in rank order: bit-identical in every CTA
      if (TWO && t_id < NRHS * NCD) {
        double v = 0.0;
        for (int rk = 0; rk < CL; ++rk) v += s_recv[rk * NREC + t_id];
        s_work[t_id] = v;  // [q][NCD]
      }
      if (t_id >= NT - NRHS) {
        const int q = t_id - (NT - NRHS);
        double v = 0.0;
        for (int rk = 0; rk < CL; ++rk) v += s_recv[rk * NREC + NRHS * NCD + q];
        s_scal[q] = v;
      }
      sync();
      if constexpr (TWO) {
        // u = E^-1 rc, one row per lane; (rc . u) partials per 32-row block
        for (int task = warp; task < NRHS * NBLK; task += NW) {
          const int q = task / NBLK, rb = task - q * NBLK;
          const double u = coarse_row_block<NCD, (L::o_work % 2 == 0 && NCD % 2 == 0)>(s_ei, s_work + q * NCD, rb, lane);
          const int i = 32 * rb + lane;
          double pr = 0.0;
          if (i < NCD) {
            s_u[q * NCD + i] = u;
            pr = u * s_work[q * NCD + i];
          }
          pr = warp_sum(pr);
          if (lane == 0) s_work[NRHS * NCD + q * NBLK + rb] = pr;
        }
        sync();
        const int Ya = cy >> 1, Za = zg >> 1;
        const int Ca = Ya + H * Za, Cb = (Ya + (cy & 1)) % H + H * ((Za + (zg & 1)) % H);
        HMX_UNROLL
        for (int q = 0; q < NRL; ++q) {
          const int qg = h * NRL + q;
          double e[D];
          HMX_UNROLL
          for (int c = 0; c < D; ++c) e[c] = 0.5 * (s_u[qg * NCD + Ca * D + c] + s_u[qg * NCD + Cb * D + c]);
          HMX_UNROLL
          for (int c = 0; c < D; ++c) {
            double v = r[q * D + c];
            HMX_UNROLL
            for (int s = c; s < D; ++s) v += Lf[cl_tri(s, c)] * e[s];
            z[q * D + c] = v;
          }
          double s = s_scal[qg];
          for (int rb = 0; rb < NBLK; ++rb) s += s_work[NRHS * NCD + qg * NBLK + rb];
          rz[q] = s;
        }
      } else {
        HMX_UNROLL
        for (int q = 0; q < NRL; ++q) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c) z[q * D + c] = r[q * D + c];
          rz[q] = s_scal[h * NRL + q];
        }
      }
    };

    // the search direction of this thread lives in its own slot of s_p only (no register copy: the slot is rewritten
    // after the CTA barrier of the p.Kp reduction, which every local reader passes after its products; the copy
    // engine has read the boundary planes by then too -- the neighbour's p.Kp partial, which the exchange waits
    // for, was sent after that neighbour had waited for the copy)
    double* p_own = s_p + (size_t)pb * NVEC + h * NVL;
    double rz[NRL], tol2[NRL];
    bool active[NRL];
    int it[NRL];
    double rz0[NRL];
    {
      double z[NVL];
      precond(rt, z, rz);
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        rz0[q] = rz[q];
        active[q] = rz0[q] > P.atol * P.atol;
        tol2[q] = fmax(P.rtol * P.rtol * rz0[q], P.atol * P.atol);
        it[q] = 0;
        HMX_UNROLL
        for (int c = 0; c < D; ++c)
          if (own) p_own[q * D + c] = active[q] ? z[q * D + c] : 0.0;
      }
    }
    int iter = 0;
    while (true) {
      // uniform over the cluster: r~.z~ comes from cluster-wide scalars that are bit-identical in every thread with the
      // same half, and every warp holds both halves
      bool mine = false;
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) mine = mine || active[q];
      if (!warp_any(mine) || iter >= P.max_it) break;
      ++iter;
      // the search directions are in place (own slots): hand the two boundary planes to the neighbours' halo planes
      sync();
      if (t_id == 0) {
        fence_async_proxy();  // the planes were written through the generic proxy
        mbar_arrive_expect_tx(s_bar, 2 * L::PLANE_BYTES);
        bulk_s2c(s_p + (size_t)(PZ + 1) * NPL * NVEC, s_p + (size_t)NPL * NVEC, L::PLANE_BYTES, s_bar, r_lo);
        bulk_s2c(s_p, s_p + (size_t)PZ * NPL * NVEC, L::PLANE_BYTES, s_bar, r_up);
      }
      // ---- y = K~ p: identity diagonal + 14 off-diagonal blocks; the products that read a halo plane come last ----
      double y[NVL];
      HMX_UNROLL
      for (int k = 0; k < NVL; ++k) y[k] = p_own[k];
      auto visit_fwd = [&](int d) {  // y_i += K~_{i,i+d} p_{i+d}
        const double* kb = s_K + (size_t)(d - 1) * NB * NOWN + j;
        const double* pn = s_p + (size_t)(pb + fwd(d)) * NVEC + h * NVL;
        double kk[NB], pp[NVL];
        HMX_UNROLL
        for (int e = 0; e < NB; ++e) kk[e] = kb[e * NOWN];
        HMX_UNROLL
        for (int k = 0; k < NVL; ++k) pp[k] = pn[k];
        HMX_UNROLL
        for (int q = 0; q < NRL; ++q)
          HMX_UNROLL
          for (int r = 0; r < D; ++r)
            HMX_UNROLL
            for (int s = 0; s < D; ++s) y[q * D + r] = fma(kk[r * D + s], pp[q * D + s], y[q * D + r]);
      };
      auto visit_bwd = [&](int d) {  // y_i += K~_{i-d,i}^T p_{i-d}
        const int off = bwd(d);
        const bool below = (d & 4) && zl == 0;  // the block lives in the halo-plane copy
        const double* kb = below ? s_Kh + (size_t)(d - 4) * NB * NPL + (j + off + NPL) : s_K + (size_t)(d - 1) * NB * NOWN + (j + off);
        const int stride = below ? NPL : NOWN;
        const double* pn = s_p + (size_t)(pb + off) * NVEC + h * NVL;
        double kk[NB], pp[NVL];
        HMX_UNROLL
        for (int e = 0; e < NB; ++e) kk[e] = kb[e * stride];
        HMX_UNROLL
        for (int k = 0; k < NVL; ++k) pp[k] = pn[k];
        HMX_UNROLL
        for (int q = 0; q < NRL; ++q)
          HMX_UNROLL
          for (int r = 0; r < D; ++r)
            HMX_UNROLL
            for (int s = 0; s < D; ++s) y[q * D + s] = fma(kk[r * D + s], pp[q * D + r], y[q * D + s]);
      };
      const bool top = zl == PZ - 1, bot = zl == 0;
      HMX_UNROLL
      for (int d = 1; d <= NH; ++d) {
        if (!((d & 4) && top)) visit_fwd(d);
        if (!((d & 4) && bot)) visit_bwd(d);
      }
      mbar_wait(s_bar, halo_parity);  // the neighbours' boundary planes have landed in the halo planes
      halo_parity ^= 1u;
      HMX_UNROLL
      for (int d = 4; d <= NH; ++d) {
        if (top) visit_fwd(d);
        if (bot) visit_bwd(d);
      }
      // ---- p.Kp over the cluster ----
      {
        double part[NRL];
        HMX_UNROLL
        for (int q = 0; q < NRL; ++q) {
          part[q] = 0.0;
          HMX_UNROLL
          for (int c = 0; c < D; ++c) part[q] += own ? p_own[q * D + c] * y[q * D + c] : 0.0;
        }
        warp_partials(part);
        sync();
        if (t_id < NRHS) {
          double v = 0.0;
          for (int w = 0; w < NW; ++w) v += s_red[w * 8 + t_id];
          HMX_UNROLL
          for (int rk = 0; rk < CL; ++rk) st_async_f64(s_xch + rank * 8 + t_id, v, s_bar_x, rk);
        }
        if (t_id == 0) mbar_arrive_expect_tx(s_bar_x, (unsigned)(CL * NRHS * 8));
        mbar_wait(s_bar_x, x_parity);
        x_parity ^= 1u;
      }
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        double pAp = 0.0;
        for (int rk = 0; rk < CL; ++rk) pAp += s_xch[rk * 8 + h * NRL + q];
        const double alpha = (active[q] && pAp > 0.0) ? fast_div(rz[q], pAp) : 0.0;
        if (active[q]) ++it[q];
        HMX_UNROLL
        for (int c = 0; c < D; ++c) {
          xt[q * D + c] = own ? fma(alpha, p_own[q * D + c], xt[q * D + c]) : 0.0;  // (idle threads keep zeros)
          rt[q * D + c] = own ? fma(-alpha, y[q * D + c], rt[q * D + c]) : 0.0;
        }
      }
      double z[NVL], rzn[NRL];
      precond(rt, z, rzn);
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        double beta = 0.0;
        if (active[q]) {
          beta = fast_div(rzn[q], rz[q]);
          rz[q] = rzn[q];
          if (!(rzn[q] > tol2[q])) active[q] = false;
        }
        HMX_UNROLL
        for (int c = 0; c < D; ++c)
          if (own) p_own[q * D + c] = active[q] ? fma(beta, p_own[q * D + c], z[q * D + c]) : 0.0;
      }
    }

    // ---- 6. epilogue: A_hom[p][q] = <C>[p][q] - b_p . x_q - x_p . r_q  (all in scaled variables) ----
    cluster_sync();  // nobody reads p or the matrix any more
    double* s_x = s_p;                  // [NOWN][NVEC] (own slots only, packed)
    double* s_r = s_K;                  // [NOWN][NVEC]
    double* s_b = s_K + NOWN * NVEC;    // [NOWN][NVEC]
    if (own) {
      HMX_UNROLL
      for (int k = 0; k < NVL; ++k) {
        s_x[(size_t)j * NVEC + h * NVL + k] = xt[k];
        s_r[(size_t)j * NVEC + h * NVL + k] = rt[k];
        s_b[(size_t)j * NVEC + h * NVL + k] = g_b[(size_t)j * NVEC + h * NVL + k];
      }
    }
    sync();
    {
      constexpr int NPAIR = NRHS * NRHS, EP = L::EPART;
      const int pair = t_id % NPAIR, part = t_id / NPAIR;
      const int p = pair / NRHS, q = pair - p * NRHS;
      double acc = 0.0;
      if (part < EP) {
        for (int i = part; i < NOWN; i += EP) {
          HMX_UNROLL
          for (int c = 0; c < D; ++c)
            acc += s_b[(size_t)i * NVEC + p * D + c] * s_x[(size_t)i * NVEC + q * D + c] +
                   s_x[(size_t)i * NVEC + p * D + c] * s_r[(size_t)i * NVEC + q * D + c];
        }
        s_work[part * NPAIR + pair] = acc;
      }
      sync();
      if (t_id < NPAIR) {
        double v = 0.0;
        for (int e = 0; e < EP; ++e) v += s_work[e * NPAIR + t_id];
        cluster_map(s_recv, 0)[rank * NREC + t_id] = v;
      }
    }
    if (P.chi != nullptr && own) {  // correctors x = L^-T x~ in natural node order: [q][component][node]
      constexpr int NN = NM * NM * NM;
      const int nat = cx + NM * (cy + NM * zg);
      // inverse of the packed lower-triangular factor
      double Li[6];
      Li[0] = 1.0 / Lf[0];
      Li[2] = 1.0 / Lf[2];
      Li[5] = 1.0 / Lf[5];
      Li[1] = -Lf[1] * Li[0] * Li[2];
      Li[4] = -Lf[4] * Li[2] * Li[5];
      Li[3] = -(Lf[3] * Li[0] + Lf[4] * Li[1]) * Li[5];
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        HMX_UNROLL
        for (int c = 0; c < D; ++c) {
          double v = 0.0;
          HMX_UNROLL
          for (int s = c; s < D; ++s) v += Li[cl_tri(s, c)] * xt[q * D + s];
          P.chi[((size_t)pt * NRHS * D + (h * NRL + q) * D + c) * NN + nat] = v;
        }
      }
    }
    // iteration statistics of all right-hand sides: one thread of each half of node 0
    if (rank == 0 && t_id < TPN) {
      HMX_UNROLL
      for (int q = 0; q < NRL; ++q) {
        s_xch[t_id * NRL + q] = (double)it[q];
        s_xch[8 + t_id * NRL + q] = rz0[q] > P.atol * P.atol ? sqrt(rz[q] / rz0[q]) : 0.0;
      }
    }
    cluster_sync();
    if (rank == 0) {
      // A_hom by thread 0, then one thread per entry of the macro element matrix
      constexpr int NBM = (D + 1) * D;
      double* s_ah = s_work;                 // [NRHS][NRHS]
      double* s_cm = s_work + NRHS * NRHS;   // [NV][NBM] macro strain matrix, then |T|
      if (t_id == 0) {
        for (int qq = 0; qq < NRHS; ++qq) {
          double e[NV], sg[NV];
          HMX_UNROLL
          for (int v = 0; v < NV; ++v) e[v] = (v == qq) ? 1.0 : 0.0;
          CO::stress(pc, smean, e, sg);
          for (int p = 0; p < NRHS; ++p) {
            double zsum = 0.0;
            for (int rk = 0; rk < CL; ++rk) zsum += s_recv[rk * NREC + p * NRHS + qq];
            s_ah[p * NRHS + qq] = sg[p] - zsum;
          }
        }
        if (P.S_loc != nullptr) {
          double Cm[NV][NBM];
          s_cm[NV * NBM] = macro_strain_matrix<D, 1>(verts, Cm);
          HMX_UNROLL
          for (int p = 0; p < NV; ++p)
            HMX_UNROLL
            for (int i = 0; i < NBM; ++i) s_cm[p * NBM + i] = Cm[p][i];
        }
        int itmax = 0;
        unsigned long long tot = 0;
        double worst = 0.0;
        for (int qq = 0; qq < NRHS; ++qq) {
          const int iq = (int)s_xch[qq];
          itmax = iq > itmax ? iq : itmax;
          tot += (unsigned long long)iq;
          worst = fmax(worst, s_xch[8 + qq]);
        }
        if (P.iters != nullptr) P.iters[pt] = itmax;
        if (P.resid != nullptr) P.resid[pt] = worst;
        if (P.work != nullptr) atomic_add_u64(P.work, tot);
      }
      sync();
      if (P.A_hom != nullptr)
        for (int k = t_id; k < NRHS * NRHS; k += NT) P.A_hom[pt * NRHS * NRHS + k] = s_ah[k];
      if (P.S_loc != nullptr) {
        for (int k = t_id; k < NBM * NBM; k += NT) {
          const int i = k / NBM, jj = k - i * NBM;
          double acc = 0.0;  // same summation order as macro_element_matrix
          HMX_UNROLL
          for (int p = 0; p < NV; ++p)
            HMX_UNROLL
            for (int q2 = 0; q2 < NV; ++q2) acc += s_cm[p * NBM + jj] * s_ah[p * NRHS + q2] * s_cm[q2 * NBM + i];
          P.S_loc[pt * NBM * NBM + k] = s_cm[NV * NBM] * acc;
        }
      }
    }
    cluster_sync();  // shared memory (also the receive buffers the peers write into) is reused by the next point
  }
}

}  // namespace hmx
