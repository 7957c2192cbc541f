// Micro cell kernel for GENERAL periodic micro meshes (HMX_VARIANT = 5): the element-list kernel of SURVEY 8f row 4.
//
// The reference's periodic cell problem accepts any micro mesh of the unit box whose boundary nodes match
// (create_periodic_boundary_conditions, cell_problem.py:16-300; BaseHMM._setup_cell_problem_variables, hmm.py:178-207);
// the stencil / matrix-free kernels of this package need the structured Kuhn split.  This kernel takes the mesh as
// DATA (MicroMesh, hmx_cell_common.cuh: element list, P1 gradients, volumes, quadrature points, the block pattern of
// the periodic stiffness matrix and fixed-order gather lists) and serves all four classes: Poisson / elasticity, 2-D /
// 3-D, optional stratification Jacobian M.  One CTA per macro point:
//   1. atoms at the quadrature points of every element -> element means (as the structured kernels: the coefficient is
//      affine in its y-dependent atoms, so the element mean of the tensor is the tensor of the mean atoms);
//   2. the periodic stiffness matrix in block CSR, every block the sum of its element contributions
//        |e| sigma(eps(phi_a e_j)) . eps(phi_b e_k),   eps = M-mapped gradient (Poisson) / its symmetric part (elasticity),
//      gathered in the fixed order of blk_src (no atomics, deterministic); load vectors of the NRHS unit strains the
//      same way over node_src  (hmm.py:652-667, 774-789, 905-922, 1050-1067);
//   3. block-Jacobi PCG for all right-hand sides in lock step, matrix and vectors in the L2-resident per-CTA scratch;
//   4. A_hom = <C> - b_p . x_q - x_p . r_q, S_loc for the macro cell -- the same epilogue as the structured kernels.
// Correctness first: this is the general path, bandwidth-bound on L2 by construction; the structured kernels remain the
// fast path for the meshes the reference's tests and examples use.
#pragma once
#include "hmx_cell_common.cuh"

namespace hmx {

template <class CO, int NT>
struct GenericLayout {
  static constexpr int D = CO::DIM;
  static constexpr int KIND = CO::KIND;
  static constexpr int BS = KIND == 0 ? 1 : D;
  static constexpr int NV = KIND == 0 ? D : D * (D + 1) / 2;  // length of a "strain": gradient / Voigt strain
  static constexpr int NRHS = NV;
  static constexpr int NW = NT / 32;
  static constexpr int NA = CO::NATOMS;
  static constexpr int NA1 = NA > 0 ? NA : 1;
  static constexpr int NREDV = NRHS > NA1 + 1 ? NRHS : NA1 + 1;
  static constexpr int total = 2 * NW * NREDV + 4 * NRHS + NRHS * NRHS + 8;  // reduction buffers, per-RHS scalars, A_hom
  static constexpr int scratch_doubles = 0;  // depends on the mesh: element_list_scratch(), computed by the host
  static_assert(NT % 32 == 0, "whole warps");
};

// sigma = C(atoms) : eps  for either kind; eps / sigma have NV entries
template <class CO>
HMX_DEV void generic_stress(const double* pc, const double* sa, const double* eps, double* sig) {
  constexpr int D = CO::DIM;
  if (CO::KIND == 0) {
    constexpr int NSYM = D * (D + 1) / 2;
    double A[NSYM];
    CO::tensor(pc, sa, A);
    HMX_UNROLL
    for (int p = 0; p < D; ++p) {
      double v = 0.0;
      HMX_UNROLL
      for (int q = 0; q < D; ++q) v += A[sym_index(D, p, q)] * eps[q];
      sig[p] = v;
    }
  } else {
    CO::stress(pc, sa, eps, sig);
  }
}

// "strain" of the basis function phi e_j whose M-mapped gradient is m
template <class CO>
HMX_DEV void generic_strain(const double* m, int j, double* e) {
  constexpr int D = CO::DIM;
  if (CO::KIND == 0) {
    HMX_UNROLL
    for (int v = 0; v < D; ++v) e[v] = m[v];
  } else {
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
}

// inverse of a small symmetric positive definite matrix (BS = 1, 2, 3), full row-major storage
template <int BS>
HMX_DEV void generic_inverse(const double* a, double* inv) {
  if (BS == 1) {
    inv[0] = a[0] != 0.0 ? 1.0 / a[0] : 0.0;
  } else if (BS == 2) {
    const double det = a[0] * a[3 % (BS * BS)] - a[1] * a[2 % (BS * BS)];
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    inv[0] = a[3 % (BS * BS)] * id;
    inv[1] = -a[1] * id;
    inv[2 % (BS * BS)] = -a[2 % (BS * BS)] * id;
    inv[3 % (BS * BS)] = a[0] * id;
  } else {
    constexpr int N2 = BS * BS;
    const double a00 = a[0], a01 = a[1], a02 = a[2 % N2], a10 = a[3 % N2], a11 = a[4 % N2], a12 = a[5 % N2], a20 = a[6 % N2],
                 a21 = a[7 % N2], a22 = a[8 % N2];
    const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    const double id = det != 0.0 ? 1.0 / det : 0.0;
    inv[0] = c00 * id;
    inv[1] = (a02 * a21 - a01 * a22) * id;
    inv[2 % N2] = (a01 * a12 - a02 * a11) * id;
    inv[3 % N2] = c01 * id;
    inv[4 % N2] = (a00 * a22 - a02 * a20) * id;
    inv[5 % N2] = (a02 * a10 - a00 * a12) * id;
    inv[6 % N2] = c02 * id;
    inv[7 % N2] = (a01 * a20 - a00 * a21) * id;
    inv[8 % N2] = (a00 * a11 - a01 * a10) * id;
  }
}

template <class CO, int NT>
HMX_DEV void generic_cell_body(const CellParams& P) {
  using L = GenericLayout<CO, NT>;
  constexpr int D = L::D, BS = L::BS, NV = L::NV, NRHS = L::NRHS, NW = L::NW, NA = L::NA, NA1 = L::NA1, NVX = D + 1;
  constexpr int NPC1 = CO::NPC > 0 ? CO::NPC : 1;
  constexpr int NVEC = NRHS * BS;  // values per node
  const MicroMesh mm = *P.mesh;
  const int NE = mm.n_elem, NP = mm.n_nodes, NZ = mm.nnzb, NQ = mm.nq;
  const long long NDOF = (long long)NP * BS;

  double* sm = dyn_smem();
  double* s_red = sm;                               // 2 x [NW][NREDV]
  double* s_scal = sm + 2 * NW * L::NREDV;          // [4][NRHS]
  double* s_ah = s_scal + 4 * NRHS;                 // [NRHS][NRHS]
  double* g = P.scratch + (size_t)bid() * (size_t)element_list_scratch(NE, NP, NZ, NA1, BS, NRHS);
  double* g_atoms = g;                              // [NA1][NE]
  double* g_K = g_atoms + (size_t)NA1 * NE;         // [NZ][BS*BS]
  double* g_dinv = g_K + (size_t)NZ * BS * BS;      // [NP][BS*BS]
  double* g_b = g_dinv + (size_t)NP * BS * BS;      // [NRHS][NDOF]
  double* g_x = g_b + NRHS * NDOF;
  double* g_r = g_x + NRHS * NDOF;
  double* g_p = g_r + NRHS * NDOF;
  double* g_y = g_p + NRHS * NDOF;
  const int t_id = tid();
  int red_flip = 0;

  for (long long pt = bid(); pt < P.n_pts; pt += nblocks()) {
    double xm[3], verts[(D + 1) * 3];
    macro_point<D>(P, pt, xm, verts);
    double pc[NPC1];
    CO::point_consts(xm, pc);
    double M[D * D];  // M[p*D+i] = d theta_i / d x_p  (hmm.py:756-757, 1015-1016)
    CO::dtheta(xm, M);

    // ---- 1. atoms: element means; their volume-weighted mean over the cell ----
    double sacc[NA1 + 1];
    HMX_UNROLL
    for (int k = 0; k <= NA1; ++k) sacc[k] = 0.0;
    for (int e = t_id; e < NE; e += NT) {
      double acc[NA1];
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) acc[k] = 0.0;
      if (NA > 0) {
        for (int q = 0; q < NQ; ++q) {
          double y[D], s[NA1];
          HMX_UNROLL
          for (int a = 0; a < D; ++a) y[a] = mm.elem_yq[((size_t)e * NQ + q) * D + a];
          CO::atoms(pc, y, s);
          const double w = P.qw[q];
          HMX_UNROLL
          for (int k = 0; k < NA1; ++k) acc[k] += w * s[k];
        }
      }
      const double ve = mm.elem_vol[e];
      HMX_UNROLL
      for (int k = 0; k < NA1; ++k) {
        g_atoms[(size_t)k * NE + e] = acc[k];
        sacc[k] += ve * acc[k];
      }
      sacc[NA1] += ve;
    }
    block_sum<NA1 + 1, NW>(sacc, s_red + (red_flip ^= 1) * NW * L::NREDV);  // (its barrier also completes g_atoms)
    const double Yvol = sacc[NA1];  // |Y| (hmm.py:101)
    double smean[NA1];
    HMX_UNROLL
    for (int k = 0; k < NA1; ++k) smean[k] = sacc[k] / Yvol;

    // M-mapped gradient of local vertex a of element e
    auto mapped_grad = [&](int e, int a, double (&m)[D]) {
      double gphys[D];
      HMX_UNROLL
      for (int i = 0; i < D; ++i) gphys[i] = mm.elem_grad[((size_t)e * NVX + a) * D + i];
      HMX_UNROLL
      for (int p = 0; p < D; ++p) {
        double v = 0.0;
        HMX_UNROLL
        for (int i = 0; i < D; ++i) v += M[p * D + i] * gphys[i];
        m[p] = v;
      }
    };

    // ---- 2. matrix blocks and load vectors, gathered in fixed order ----
    for (int s = t_id; s < NZ; s += NT) {
      double blk[BS * BS];
      HMX_UNROLL
      for (int k = 0; k < BS * BS; ++k) blk[k] = 0.0;
      for (int c = mm.blk_ptr[s]; c < mm.blk_ptr[s + 1]; ++c) {
        const int src = mm.blk_src[c];
        const int b = src % NVX, a = (src / NVX) % NVX, e = src / (NVX * NVX);
        double ma[D], mb[D], sa[NA1];
        mapped_grad(e, a, ma);
        mapped_grad(e, b, mb);
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) sa[k] = g_atoms[(size_t)k * NE + e];
        const double ve = mm.elem_vol[e];
        HMX_UNROLL
        for (int j = 0; j < BS; ++j) {
          double ea[NV], sg[NV];
          generic_strain<CO>(ma, j, ea);
          generic_stress<CO>(pc, sa, ea, sg);
          HMX_UNROLL
          for (int j2 = 0; j2 < BS; ++j2) {
            double eb[NV];
            generic_strain<CO>(mb, j2, eb);
            double v = 0.0;
            HMX_UNROLL
            for (int k = 0; k < NV; ++k) v += sg[k] * eb[k];
            blk[j * BS + j2] += ve * v;
          }
        }
      }
      HMX_UNROLL
      for (int k = 0; k < BS * BS; ++k) g_K[(size_t)s * BS * BS + k] = blk[k];
    }
    for (int i = t_id; i < NP; i += NT) {
      double bb[NRHS * BS];
      HMX_UNROLL
      for (int k = 0; k < NRHS * BS; ++k) bb[k] = 0.0;
      for (int c = mm.node_ptr[i]; c < mm.node_ptr[i + 1]; ++c) {
        const int src = mm.node_src[c];
        const int a = src % NVX, e = src / NVX;
        double ma[D], sa[NA1];
        mapped_grad(e, a, ma);
        HMX_UNROLL
        for (int k = 0; k < NA1; ++k) sa[k] = g_atoms[(size_t)k * NE + e];
        const double ve = mm.elem_vol[e];
        HMX_UNROLL
        for (int j = 0; j < BS; ++j) {
          double ea[NV], sg[NV];
          generic_strain<CO>(ma, j, ea);
          generic_stress<CO>(pc, sa, ea, sg);
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) bb[q * BS + j] -= ve * sg[q];  // -(C E_q) : eps(phi_a e_j) = -(C eps)[q]
        }
      }
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q)
        HMX_UNROLL
        for (int j = 0; j < BS; ++j) g_b[q * NDOF + (long long)i * BS + j] = bb[q * BS + j];
    }
    sync();
    for (int i = t_id; i < NP; i += NT) {
      double blk[BS * BS], inv[BS * BS];
      HMX_UNROLL
      for (int k = 0; k < BS * BS; ++k) blk[k] = g_K[(size_t)mm.diag[i] * BS * BS + k];
      generic_inverse<BS>(blk, inv);
      HMX_UNROLL
      for (int k = 0; k < BS * BS; ++k) g_dinv[(size_t)i * BS * BS + k] = inv[k];
    }
    sync();

    // ---- 3. PCG, all right-hand sides in lock step (rows are owned by threads: i = t_id, t_id + NT, ...) ----
    auto precond_row = [&](int i, const double* r, double* z) {  // r, z: [NRHS][BS] of row i
      double dv[BS * BS];
      HMX_UNROLL
      for (int k = 0; k < BS * BS; ++k) dv[k] = g_dinv[(size_t)i * BS * BS + k];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q)
        HMX_UNROLL
        for (int j = 0; j < BS; ++j) {
          double v = 0.0;
          HMX_UNROLL
          for (int k = 0; k < BS; ++k) v += dv[j * BS + k] * r[q * BS + k];
          z[q * BS + j] = v;
        }
    };
    double part[NRHS];
    HMX_UNROLL
    for (int q = 0; q < NRHS; ++q) part[q] = 0.0;
    for (int i = t_id; i < NP; i += NT) {
      double r[NVEC], z[NVEC];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q)
        HMX_UNROLL
        for (int j = 0; j < BS; ++j) r[q * BS + j] = g_b[q * NDOF + (long long)i * BS + j];
      precond_row(i, r, z);
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q)
        HMX_UNROLL
        for (int j = 0; j < BS; ++j) {
          const long long a = q * NDOF + (long long)i * BS + j;
          g_x[a] = 0.0;
          g_r[a] = r[q * BS + j];
          g_p[a] = z[q * BS + j];
          part[q] += r[q * BS + j] * z[q * BS + j];
        }
    }
    block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);  // (its barrier publishes p)
    double rz[NRHS], rz0[NRHS], tol2[NRHS];
    bool active[NRHS];
    int it[NRHS];
    HMX_UNROLL
    for (int q = 0; q < NRHS; ++q) {
      rz[q] = rz0[q] = part[q];
      active[q] = rz0[q] > P.atol * P.atol;
      tol2[q] = fmax(P.rtol * P.rtol * rz0[q], P.atol * P.atol);
      it[q] = 0;
    }
    for (int iter = 0; iter < P.max_it; ++iter) {
      bool go = false;
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) go = go || active[q];
      if (!go) break;
      // y = K p (block CSR rows), p . y
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) part[q] = 0.0;
      for (int i = t_id; i < NP; i += NT) {
        double y[NVEC];
        HMX_UNROLL
        for (int k = 0; k < NVEC; ++k) y[k] = 0.0;
        for (int s = mm.row_ptr[i]; s < mm.row_ptr[i + 1]; ++s) {
          const int jn = mm.col[s];
          double kb[BS * BS];
          HMX_UNROLL
          for (int k = 0; k < BS * BS; ++k) kb[k] = g_K[(size_t)s * BS * BS + k];
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            double pj[BS];
            HMX_UNROLL
            for (int k = 0; k < BS; ++k) pj[k] = g_p[q * NDOF + (long long)jn * BS + k];
            HMX_UNROLL
            for (int j = 0; j < BS; ++j)
              HMX_UNROLL
              for (int k = 0; k < BS; ++k) y[q * BS + j] += kb[j * BS + k] * pj[k];
          }
        }
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q)
          HMX_UNROLL
          for (int j = 0; j < BS; ++j) {
            const long long a = q * NDOF + (long long)i * BS + j;
            g_y[a] = y[q * BS + j];
            part[q] += g_p[a] * y[q * BS + j];
          }
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);
      double alpha[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        alpha[q] = (active[q] && part[q] > 0.0) ? rz[q] / part[q] : 0.0;
        if (active[q]) ++it[q];
        part[q] = 0.0;
      }
      // x += alpha p, r -= alpha y, r . z   (own rows only: no barrier needed in between)
      for (int i = t_id; i < NP; i += NT) {
        double r[NVEC], z[NVEC];
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q)
          HMX_UNROLL
          for (int j = 0; j < BS; ++j) {
            const long long a = q * NDOF + (long long)i * BS + j;
            g_x[a] += alpha[q] * g_p[a];
            r[q * BS + j] = g_r[a] - alpha[q] * g_y[a];
            g_r[a] = r[q * BS + j];
          }
        precond_row(i, r, z);
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q)
          HMX_UNROLL
          for (int j = 0; j < BS; ++j) {
            part[q] += r[q * BS + j] * z[q * BS + j];
            g_y[q * NDOF + (long long)i * BS + j] = z[q * BS + j];  // z waits in y for the update of p
          }
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);  // (every thread has finished reading p of other rows)
      double beta[NRHS];
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) {
        beta[q] = 0.0;
        if (active[q]) {
          beta[q] = part[q] / rz[q];
          rz[q] = part[q];
          if (!(part[q] > tol2[q])) active[q] = false;
        }
      }
      for (int i = t_id; i < NP; i += NT) {
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q)
          HMX_UNROLL
          for (int j = 0; j < BS; ++j) {
            const long long a = q * NDOF + (long long)i * BS + j;
            g_p[a] = active[q] ? g_y[a] + beta[q] * g_p[a] : 0.0;
          }
      }
      sync();  // publish p
    }

    // ---- 4. A_hom[p][q] = <C>[p][q] - b_p . x_q - x_p . r_q ----
    for (int p = 0; p < NRHS; ++p) {
      HMX_UNROLL
      for (int q = 0; q < NRHS; ++q) part[q] = 0.0;
      for (int i = t_id; i < NP; i += NT) {
        for (int j = 0; j < BS; ++j) {
          const double bp = g_b[p * NDOF + (long long)i * BS + j], xp = g_x[p * NDOF + (long long)i * BS + j];
          HMX_UNROLL
          for (int q = 0; q < NRHS; ++q) {
            const long long a = q * NDOF + (long long)i * BS + j;
            part[q] += bp * g_x[a] + xp * g_r[a];
          }
        }
      }
      block_sum<NRHS, NW>(part, s_red + (red_flip ^= 1) * NW * L::NREDV);
      if (t_id == 0) {
        HMX_UNROLL
        for (int q = 0; q < NRHS; ++q) s_ah[p * NRHS + q] = part[q];
      }
    }
    if (P.chi != nullptr) {  // correctors on the PERIODIC nodes: [q][component][node]
      for (int i = t_id; i < NP; i += NT)
        for (int q = 0; q < NRHS; ++q)
          for (int j = 0; j < BS; ++j) P.chi[((size_t)pt * NRHS * BS + q * BS + j) * NP + i] = g_x[q * NDOF + (long long)i * BS + j];
    }
    sync();
    if (t_id == 0) {
      double Ah[NRHS * NRHS];
      for (int qq = 0; qq < NRHS; ++qq) {
        double e[NV], sg[NV];
        HMX_UNROLL
        for (int v = 0; v < NV; ++v) e[v] = (v == qq) ? 1.0 : 0.0;
        generic_stress<CO>(pc, smean, e, sg);
        for (int p = 0; p < NRHS; ++p) Ah[p * NRHS + qq] = sg[p] - s_ah[p * NRHS + qq] / Yvol;
      }
      if (P.A_hom != nullptr)
        for (int k = 0; k < NRHS * NRHS; ++k) P.A_hom[pt * NRHS * NRHS + k] = Ah[k];
      if (P.S_loc != nullptr) macro_element_matrix<D, CO::KIND>(verts, Ah, P.S_loc + pt * (D + 1) * BS * (D + 1) * BS);
      int itmax = 0;
      unsigned long long tot = 0;
      double worst = 0.0;
      for (int qq = 0; qq < NRHS; ++qq) {
        itmax = it[qq] > itmax ? it[qq] : itmax;
        tot += (unsigned long long)it[qq];
        if (rz0[qq] > P.atol * P.atol) worst = fmax(worst, sqrt(rz[qq] / rz0[qq]));
      }
      if (P.iters != nullptr) P.iters[pt] = itmax;
      if (P.resid != nullptr) P.resid[pt] = worst;
      if (P.work != nullptr) atomic_add_u64(P.work, tot);
    }
    sync();  // shared memory and the scratch are reused by the next macro point
  }
}

}  // namespace hmx
