/* hmx.h -- C ABI of libhmx.so, the B200 (sm_100a) replacement for the hot path of
 * flxrcz/hommx: the per-macro-quadrature-point periodic micro cell problems and the
 * macro stiffness assembly they feed.
 *
 * The reference has no FFI for this path: the seam is the Python method
 * BaseHMM._assemble_stiffness (src/hommx/hmm.py:298-332), which loops over macro cells,
 * calls _compute_local_stiffness (hmm.py:334-369) and MatSetValues(ADD_VALUES)
 * (hmm.py:325-330).  The entry points below are what a binding for that seam binds;
 * INTEGRATION.md shows the ctypes stub a maintainer would add to hmm.py.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Every function returns 0 on
 * success or a negative hmx_status; the message is available from hmx_last_error().  A
 * handle is not thread-safe; it owns one CUDA stream (replaceable with hmx_set_stream) and
 * its device buffers; the caller owns every buffer it passes in.  Functions ending in _dev
 * take device pointers and only enqueue work on the handle's stream (no synchronisation);
 * the others take host pointers, copy in and out on that stream and return after
 * synchronising it.  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with HMX_ERR_CUDA.
 */
#ifndef HMX_H
#define HMX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hmx_handle hmx_t;

/* Bumped whenever a signature, the layout of hmx_desc or the kernel-image contract (hmx_info, CellParams) changes;
 * a binding checks it against the value it was written for before calling anything else (a stale libhmx.so next
 * to new Python sources would otherwise be called with the wrong arguments). */
#define HMX_ABI_VERSION 8
int32_t hmx_abi_version(void);

enum hmx_status {
  HMX_OK = 0,
  HMX_ERR_ARG = -1,     /* invalid argument (the reference raises ValueError, hmm.py:104-115) */
  HMX_ERR_CUDA = -2,    /* CUDA runtime/driver failure, or no device */
  HMX_ERR_KERNEL = -3,  /* cell kernel image missing/incompatible with the descriptor */
  HMX_ERR_STATE = -4,   /* call sequence error (e.g. assemble before hmx_macro_plan) */
  HMX_ERR_NOMEM = -5
};

enum hmx_kind { HMX_POISSON = 0, HMX_ELASTICITY = 1 };

/* Describes one solver object = one (problem class, coefficient, micro mesh) combination;
 * replaces BaseHMM._setup_cell_problem_variables / _setup_cell_problem_forms
 * (hmm.py:178-207, 259-274). */
/* General periodic micro mesh: any simplicial mesh of the unit box whose boundary nodes match
 * periodically -- what create_periodic_boundary_conditions (cell_problem.py:16-300) accepts.  Replaces the structured
 * n_micro^dim description for the element-list kernel (csrc/hmx_cell_generic.cuh, kernel images built with
 * HMX_VARIANT=5).  HOST pointers; hmx_create copies everything to the device.  hommx_b200/micro.py
 * (ElementListTables) builds these arrays from the mesh. */
typedef struct hmx_micro_mesh {
  int32_t n_elem;             /* micro elements */
  int32_t n_nodes;            /* PERIODIC nodes (slaves identified with their masters) */
  int32_t nnzb;               /* blocks of the node-node pattern of the periodic stiffness matrix */
  const int32_t* elem_nodes;  /* [n_elem][dim+1] periodic node ids */
  const double* elem_grad;    /* [n_elem][dim+1][dim] gradients of the P1 basis functions */
  const double* elem_vol;     /* [n_elem] */
  const double* elem_yq;      /* [n_elem][nq][dim] quadrature points in micro coordinates */
  const int32_t* row_ptr;     /* [n_nodes+1] block rows */
  const int32_t* col;         /* [nnzb] */
  const int32_t* blk_ptr;     /* [nnzb+1] contributions of every block ... */
  const int32_t* blk_src;     /* ... as (e*(dim+1) + a)*(dim+1) + b, in the order they are summed */
  const int32_t* node_ptr;    /* [n_nodes+1] elements around every node ... */
  const int32_t* node_src;    /* ... as e*(dim+1) + a */
  const int32_t* diag;        /* [n_nodes] block index of (i, i) */
} hmx_micro_mesh;

typedef struct hmx_desc {
  int32_t dim;       /* 2 or 3 (micro and macro dimension agree, hmm.py:114-115) */
  int32_t kind;      /* hmx_kind */
  int32_t n_micro;   /* cells per axis of the structured periodic micro mesh (0 with micro_mesh) */
  int32_t nq;        /* quadrature points per micro element */
  const double* qp;  /* [T][nq][dim] points per element type, cube-local, units of h (T = dim!) */
  const double* qw;  /* [nq] weights, normalised to sum 1 */
  const void* kernel_image; /* cubin/fatbin of the cell kernel specialised for the coefficient, the micro mesh size and
                             * the solver variant (matrix-free PCG, or the direct dense-Cholesky kernel for elasticity
                             * cells of <= 192 unknowns).  The direct kernel ignores rtol / atol /
                             * max_it and reports 0 iterations. */
  size_t kernel_image_size;
  double rtol;       /* PCG: stop at sqrt(r.z) <= max(rtol*sqrt(r0.z0), atol) (ksp_rtol / ksp_atol) */
  double atol;
  int32_t max_it;    /* ksp_max_it */
  int32_t device;    /* CUDA device ordinal */
  const hmx_micro_mesh* micro_mesh; /* NULL: the structured n_micro^dim mesh; else a general periodic mesh (qp is
                                     * ignored then: the quadrature points are micro_mesh->elem_yq; qw still [nq]) */
} hmx_desc;

int hmx_create(hmx_t** out, const hmx_desc* desc);
void hmx_destroy(hmx_t* h);
/* message of the last failure on `h`; with h == NULL the last hmx_create failure */
const char* hmx_last_error(const hmx_t* h);

/* use an existing CUDA stream (cudaStream_t) for all work of this handle */
int hmx_set_stream(hmx_t* h, void* cuda_stream);
int hmx_set_tolerances(hmx_t* h, double rtol, double atol, int32_t max_it);
/* 0 = one resident wave of CTAs (default), otherwise the grid size to launch */
int hmx_set_grid(hmx_t* h, int32_t n_ctas);

/* static facts about the loaded kernel: info[0]=dynamic smem bytes, [1]=threads per CTA,
 * [2]=right-hand sides per point, [3]=m (A_hom is m x m), [4]=n_b (S_loc is n_b x n_b),
 * [5]=CTAs per SM, [6]=SM count, [7]=per-CTA scratch doubles */
int hmx_kernel_info(const hmx_t* h, int32_t info[8]);
/* thread-block clusters of the loaded kernel: info[0]=CTAs per cluster (1: ordinary launch; > 1: the cluster variant of
 * the 3-D elasticity kernel, csrc/hmx_cell_cluster.cuh -- one macro point per cluster, the assembled stencil resident
 * in the cluster's distributed shared memory), info[1]=clusters resident on the device at once (= the grid launched) */
int hmx_cluster_info(const hmx_t* h, int32_t info[2]);

/* Homogenised tensors at given macro points: replaces the n_b corrector solves and n_b^2
 * assemble_scalar calls of _compute_local_stiffness (hmm.py:354-364) in the d-RHS form of
 * BasePeriodicHMM.compute_effective_tensor (hmm.py:1219-1245).
 *   x_pts [n_pts][3], A_hom [n_pts][m][m], iters [n_pts] (may be NULL), resid [n_pts] (may be NULL) */
int hmx_cell_tensors(hmx_t* h, int64_t n_pts, const double* x_pts, double* A_hom, int32_t* iters, double* resid);
int hmx_cell_tensors_dev(hmx_t* h, int64_t n_pts, const double* x_pts, double* A_hom, int32_t* iters, double* resid);

/* Same solve, additionally returning the correctors chi_q (BasePeriodicHMM.correctors, hmm.py:1211-1213):
 * chi [n_pts][n_rhs][bs][N_grid] nodal values on the periodic micro grid in natural node order (x fastest;
 * collapsed axes have extent 1), defined up to an additive constant per component.  A_hom may be NULL.
 * Device pointers. */
int hmx_cell_correctors_dev(hmx_t* h, int64_t n_pts, const double* x_pts, double* A_hom, double* chi);

/* Macro stiffness assembly: replaces BaseHMM._assemble_stiffness (hmm.py:298-332) for the
 * `n_cells` macro cells given (a rank's owned cells, hmm.py:307).
 *   cell_nodes [n_cells][dim+1]  macro vertex ids (geometry dofmap)
 *   node_xyz   [n_nodes][3]      macro vertex coordinates (msh.geometry.x)
 *   gather_ptr [nnz+1], gather_src [gather_ptr[nnz]]: for CSR value slot s the sources
 *       S_loc_flat[gather_src[j]], j in [gather_ptr[s], gather_ptr[s+1]), S_loc_flat being
 *       the [n_cells][n_b][n_b] row-major local matrices (row = unrolled dof i, hmm.py:31-40,
 *       325-330).  The fixed order makes the sum deterministic.
 *   csr_vals   [nnz]  out: summed values (un-BC'd, what self._A holds before hmm.py:442)
 *   S_loc      [n_cells][n_b*n_b] out, may be NULL;  iters/resid [n_cells], may be NULL */
int hmx_assemble_macro(hmx_t* h, int64_t n_cells, const int32_t* cell_nodes, int64_t n_nodes, const double* node_xyz,
                       int64_t nnz, const int64_t* gather_ptr, const int32_t* gather_src, double* csr_vals,
                       double* S_loc, int32_t* iters, double* resid);
int hmx_assemble_macro_dev(hmx_t* h, int64_t n_cells, const int32_t* cell_nodes, int64_t n_nodes,
                           const double* node_xyz, int64_t nnz, const int64_t* gather_ptr, const int32_t* gather_src,
                           double* csr_vals, double* S_loc, int32_t* iters, double* resid);

/* The two stages of hmx_assemble_macro_dev can also be run separately: call
 * hmx_assemble_macro_dev with nnz = 0 (cell kernel only, S_loc required), then this gather. */
int hmx_gather_csr_dev(hmx_t* h, int64_t nnz, const int64_t* gather_ptr, const int32_t* gather_src, const double* S_loc,
                       double* csr_vals);

/* Sum over all points and right-hand sides of the PCG iterations executed by the cell kernels
 * of this handle since the last reset (the "I" of the algorithmic FLOP count).
 * Synchronises the stream. */
int hmx_rhs_iterations(hmx_t* h, int64_t* total, int32_t reset);

/* Halo exchange helpers for macro cells sharded over several GPUs (the one exchange step of
 * the reference: MatAssembly of shared rows, hmm.py:442).  pack: buf[j] = csr_vals[slots[j]];
 * the caller sums `buf` across ranks (torch.distributed / NCCL all-reduce); unpack writes it
 * back.  All pointers are device pointers. */
int hmx_halo_pack_dev(hmx_t* h, const double* csr_vals, const int64_t* slots, int64_t n, double* buf);
int hmx_halo_unpack_dev(hmx_t* h, double* csr_vals, const int64_t* slots, int64_t n, const double* buf);
/* The whole exchange for a host that holds an NCCL communicator (SURVEY 8b proposal `hmx_halo_sum`; replaces the PETSc
 * assembly of shared rows, hmm.py:442): pack, ncclAllReduce(buf, n doubles, sum) on the handle's stream, unpack.
 * `nccl_comm` is the caller's ncclComm_t; libhmx looks ncclAllReduce up in the NCCL library already loaded in the
 * process (no link-time dependency).  Every rank of the communicator must call it, with the same n (the shared-slot
 * list is global: slots touched by more than one rank, in the same order everywhere).  All pointers are device
 * pointers. */
int hmx_halo_sum_dev(hmx_t* h, void* nccl_comm, double* csr_vals, const int64_t* slots, int64_t n, double* buf);

/* SURVEY 8f row 4, "higher-order macro quadrature": macro element matrices from homogenised tensors at nq macro
 * quadrature points per cell,  S_loc = |T| C^T (sum_q weights[q] A_pts[cell][q]) C  (the reference evaluates the cell
 * problem at the barycentre only, hmm.py:349-352; with P1 macro elements C is constant per cell, so a higher-order rule
 * acts on the tensor alone).  A_pts [n_cells][nq][m][m] comes from hmx_cell_tensors_dev at the mapped points; weights
 * [nq] (HOST, sum 1); S_loc [n_cells][n_b][n_b] then goes through hmx_gather_csr_dev as usual.  Device pointers. */
int hmx_macro_elements_dev(hmx_t* h, int64_t n_cells, const int32_t* cell_nodes, const double* node_xyz, int32_t nq,
                           const double* weights, const double* A_pts, double* S_loc);

/* the macro load vector on the device.  Replaces the FFCx kernel behind
 * `_assemble_vector_array(b_local.array_w, self._L, ...)` (hmm.py:445-450; L = inner(f(x), v) dx, hmm.py:131-133).
 * `image` is the cubin of csrc/hmx_load_entry.cu specialised for f (hommx_b200.codegen.build_load_program); qp [nq][dim]
 * / qw [nq] is the quadrature rule of the UFL-estimated degree on the reference simplex (HOST pointers, tiny).  Writes
 * the element vectors Fe [n_cells][(dim+1)*bs] (device); b is then their deterministic gather:
 * hmx_gather_csr_dev(h, n_dofs, ptr, src, Fe, b) with the gather map of the dof numbering.  cell_nodes / node_xyz / Fe
 * are device pointers; work is enqueued on the handle's stream. */
int hmx_macro_load_dev(hmx_t* h, const void* image, size_t image_size, int64_t n_cells, const int32_t* cell_nodes,
                       const double* node_xyz, int32_t nq, const double* qp, const double* qw, double* Fe);

/* SURVEY 8f rows 2-3 (outside the north-star hot path; the reference does this with PETSc, hmm.py:453-491):
 * Dirichlet lifting on the device -- b -= A u_bc on the free rows, constrained rows and columns zeroed with a unit
 * diagonal (MatZeroRowsColumns), b = value on the constrained dofs; bc_mask [n_dofs] is 1 on constrained dofs,
 * bc_values [n_dofs] holds their values (anything elsewhere).  The CSR index arrays are those of the pattern
 * the gather map was built for.  All pointers are device pointers. */
int hmx_macro_lift_dev(hmx_t* h, int64_t n_dofs, const int64_t* indptr, const int32_t* indices, double* csr_vals,
                       const int8_t* bc_mask, const double* bc_values, double* b);
/* Jacobi-preconditioned CG for the lifted (symmetric positive definite) macro system, x0 = 0; stops at
 * sqrt(r.z) <= max(rtol sqrt(r0.z0), atol) or max_it.  iters / resid are HOST pointers (may be NULL).
 * Dot products are reduced in two fixed-order stages (no floating-point atomics): bitwise reproducible.
 * Synchronises the stream. */
int hmx_macro_pcg_dev(hmx_t* h, int64_t n_dofs, const int64_t* indptr, const int32_t* indices, const double* csr_vals,
                      const double* b, double* x, double rtol, double atol, int32_t max_it, int32_t* iters, double* resid);

/* Roofline denominators measured on this device: FP64 FMA throughput (register-resident DFMA
 * chains on every SM) in TFLOP/s and a device-to-device copy in GB/s (read+write bytes). */
int hmx_measure_peaks(int32_t device, double* fp64_tflops, double* copy_gbs);

int hmx_sync(hmx_t* h);

#ifdef __cplusplus
}
#endif
#endif /* HMX_H */
