"""Drop-in HMM solver classes on the B200 hot path.

Same constructor arguments and ``solve()`` / ``set_boundary_conditions()`` /
``set_right_hand_side()`` / ``function_space`` surface as the reference's
``PoissonHMM``, ``PoissonStratifiedHMM``, ``LinearElasticityHMM`` and
``LinearElasticityStratifiedHMM`` (/root/reference/src/hommx/hmm.py:514-1067).  What changes is
``_assemble_stiffness`` (hmm.py:298-332): instead of a Python loop over macro cells with ``n_b``
PETSc solves and ``n_b^2`` ``assemble_scalar`` calls per cell, the owned cells go to the GPU in
one call (``hmx_assemble_macro``): coefficient evaluation, micro solves, A_hom, S_loc and the
deterministic gather into the CSR value array all happen on the device.  When macro cells are
sharded over several GPUs (``torch.distributed`` initialised, one process per GPU) the only
collective is the sum of the value slots shared between ranks.

Host code stays Python.  DOLFINx / PETSc are not installable in the build image, so the mesh,
function-space and Dirichlet objects are the light stand-ins of ``hommx_b200.mesh`` /
``hommx_b200.fem`` (a ``dolfinx.mesh.Mesh`` is accepted wherever its ``geometry`` /
``topology`` attributes suffice), coefficients are traced with ``hommx_b200.ufl`` (UFL's
operator names), and the macro linear solve -- outside the hot path, PETSc KSP in the
reference (hmm.py:482-483) -- is a Jacobi-PCG on the device (``hmx_macro_pcg_dev``) on the CSR values the
assembly left there, or scipy's sparse LU on request (``{"pc_type": "lu"}``).
"""
from __future__ import annotations

import logging

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import assembly, codegen, fem, micro, native, parallel, quadrature
from .mesh import as_simplex_mesh

_CELL_OPTION_KEYS = {"ksp_rtol", "ksp_atol", "ksp_max_it"}
# cell_solver="auto": mean PCG iterations per right-hand side (on a sample of macro cells) above which the small
# elasticity cells are factorised directly instead (measured cross-over on B200: 15-40, scripts/probe_dense.py)
DIRECT_ABOVE_ITERATIONS = 30


def _triangle_area(points):
    """hmm.py:20-23"""
    p = np.asarray(points, dtype=float)
    return 0.5 * np.linalg.norm(np.cross(p[1] - p[0], p[2] - p[0]))


def _tetrahedron_volume(points):
    """hmm.py:25-28"""
    p = np.asarray(points, dtype=float)
    return abs(np.linalg.det(np.array([p[1] - p[0], p[2] - p[0], p[3] - p[0]]))) / 6.0


def _macro_solve(A, b, options):
    """Stand-in for the macro PETSc KSP (outside the hot path).  Small systems and ``pc_type: lu`` use a
    sparse direct solve; larger ones Jacobi-preconditioned CG (the lifted HMM matrix is symmetric positive
    definite for symmetric coefficients), honouring ``ksp_rtol`` / ``ksp_atol`` / ``ksp_max_it``, with the
    direct solve as the last resort."""
    n = A.shape[0]
    if n <= 20000 or options.get("pc_type") == "lu" or options.get("ksp_type") == "preonly":
        return spla.spsolve(A.tocsc(), b)
    rtol = float(options.get("ksp_rtol", 1e-12))
    atol = float(options.get("ksp_atol", 0.0))
    d = A.diagonal()
    if np.all(d > 0):
        M = spla.LinearOperator(A.shape, lambda v: v / d)
        x, info = spla.cg(A, b, M=M, rtol=rtol, atol=atol, maxiter=int(options.get("ksp_max_it", 20000)))
        if info == 0 and np.all(np.isfinite(x)) and np.linalg.norm(A @ x - b) <= 1e-9 * max(np.linalg.norm(b), 1e-300):
            return x
    return spla.spsolve(A.tocsc(), b)


class BaseHMM:
    """Common driver (mirrors ``BaseHMM``, hmm.py:53-511)."""

    _KIND = codegen.POISSON

    def __init__(
        self,
        msh,
        A,
        f,
        msh_micro,
        eps,
        petsc_options_global_solve=None,
        petsc_options_cell_problem=None,
        petsc_options_prefix="hommx_HMM",
        *,
        Dtheta_transpose=None,
        quadrature_rule=None,
        device=None,
        shard=True,
        collapse_invariant_axes=True,
        cell_solver="auto",
        macro_quadrature_degree=None,
    ):
        self._logger = logging.getLogger(__name__)
        self._msh = as_simplex_mesh(msh)
        self._cell_mesh = as_simplex_mesh(msh_micro)
        self._coeff = A
        self._f = f
        self._eps = eps
        self._tdim = self._msh.dim
        # same checks, same messages as hmm.py:104-115
        if self._tdim not in (2, 3):
            raise ValueError("Topology should be 3D or 2D")
        if self._tdim == 2 and np.any(self._msh.x[:, 2] != 0.0):
            raise ValueError(
                "Topological dimension is different from geometrical dimension. Currently surfaces in 3D are not supported."
            )
        if self._cell_mesh.dim == 2 and np.any(self._cell_mesh.x[:, 2] != 0.0):
            raise ValueError("Topological dimension is different from geometrical dimension for micro mesh.")
        if self._tdim != self._cell_mesh.dim:
            raise ValueError("Micro and macro mesh should have the same dimensionality.")
        self._volume_function = _tetrahedron_volume if self._tdim == 3 else _triangle_area

        self._bs = 1 if self._KIND == codegen.POISSON else self._tdim
        self._V_macro = fem.FunctionSpace(self._msh, self._bs)
        self._macro_coordinates = self._V_macro.tabulate_dof_coordinates()
        self._num_basis_functions_per_cell = (self._tdim + 1) * self._bs
        self._num_global_dofs = self._V_macro.num_dofs
        self._u = fem.Function(self._V_macro)
        self._b = np.zeros(self._num_global_dofs)
        self._needs_reassembly = True
        self._bcs = []

        if petsc_options_cell_problem is None:
            petsc_options_cell_problem = {"ksp_atol": 1e-10}  # hmm.py:153-155
        self._petsc_options_cell_problem = dict(petsc_options_cell_problem)
        ignored = sorted(set(self._petsc_options_cell_problem) - _CELL_OPTION_KEYS)
        if ignored:
            self._logger.warning(
                "cell problems are solved by the CUDA PCG kernel; PETSc options %s have no meaning there and are ignored", ignored
            )
        # PETSc's default rtol (1e-5, left in force by the reference) bounds the residual of GMRES+ILU;
        # the PCG kernel stops on the energy-norm residual, for which 1e-8 gives A_hom to ~1e-12.
        self._cell_rtol = float(self._petsc_options_cell_problem.get("ksp_rtol", 1e-8))
        self._cell_atol = float(self._petsc_options_cell_problem.get("ksp_atol", 1e-10))
        self._cell_max_it = int(self._petsc_options_cell_problem.get("ksp_max_it", 10000))
        self._petsc_options_global_solve = dict(petsc_options_global_solve or {})
        self._petsc_options_prefix = petsc_options_prefix

        # ---- _setup_cell_problem_variables (hmm.py:178-207) ----
        self._program = codegen.build_program(A, self._tdim, self._KIND, Dtheta_transpose)
        pts, wts = quadrature_rule if quadrature_rule is not None else quadrature.default_rule(self._tdim, self._program.degree)
        try:
            # fast path: the structured simplicial split of the unit box (every mesh the reference's tests and examples use)
            self._structure = micro.detect_structure(self._cell_mesh)
            self._micro_tables = None
            self._qp, self._qw = micro.quadrature_table(self._structure, pts, wts)
        except ValueError as structured_only:
            # any other periodic micro mesh (the reference's MPC path accepts every mesh with matching boundary nodes,
            # cell_problem.py:16-35): the element-list kernel, csrc/hmx_cell_generic.cuh
            try:
                self._micro_tables = micro.ElementListTables(self._cell_mesh, pts, wts)
            except ValueError as e:
                raise ValueError(f"{e} (and not a structured grid: {structured_only})") from None
            self._structure = None
            self._qp, self._qw = None, self._micro_tables.qw
            self._logger.info("general periodic micro mesh (%d elements, %d periodic nodes): element-list kernel",
                              self._micro_tables.n_elem, self._micro_tables.n_nodes)
        self._cell_mesh_area = 1.0  # |Y| of the unit box (hmm.py:101)

        # ---- macro sparsity and slot map (replaces the un-preallocated AIJ of hmm.py:144-149) ----
        self._pattern = assembly.build_pattern(self._msh.cells, self._msh.num_nodes, self._bs)
        self._A_values = np.zeros(self._pattern.nnz)
        self._A = None
        self._device = device
        # exact symmetry reduction: micro axes the coefficient does not depend on are solved on one layer of
        # cubes (csrc/hmx_cell_common.cuh, Grid<.., COLL>); False solves the full n^d cell
        self._collapse = bool(collapse_invariant_axes)
        # macro quadrature: None / 0 / 1 = the barycentre of every macro cell, as the reference (hmm.py:349-352); k >= 2 =
        # the simplex rule of degree k: the cell problem is solved at every rule point and S_loc = |T| C^T (sum_q w_q
        # A_hom(x_q)) C (hmx_macro_elements_dev)
        self._macro_qdeg = int(macro_quadrature_degree or 0)
        # "pcg": the matrix-free PCG kernels; "direct": dense Cholesky per cell (elasticity cells of <= 192 unknowns,
        # csrc/hmx_cell_dense.cuh); "cluster": PCG on the assembled stencil resident in a thread-block cluster's
        # distributed shared memory (3-D elasticity, full cells; csrc/hmx_cell_cluster.cuh); "auto": direct where the
        # PCG iteration count of a sample of macro cells says it pays, cluster for cells that exceed one SM
        if cell_solver not in ("auto", "pcg", "direct", "cluster"):
            raise ValueError("cell_solver must be 'auto', 'pcg', 'direct' or 'cluster'")
        self._cell_solver = cell_solver
        self.cell_solver_used = None
        self._shard = bool(shard)  # False: assemble every macro cell on this GPU even if torch.distributed is up
        self._solver = None
        self._dev = None
        self._rank, self._world = 0, 1
        self.cell_iterations = None
        self.cell_residuals = None
        self.assembly_stats = None
        self.macro_solve_stats = None

    # ------------------------------------------------------------------ API surface
    @property
    def function_space(self):
        """Function space of the macro mesh (hmm.py:173-176)."""
        return self._V_macro

    def set_boundary_conditions(self, bcs):
        """hmm.py:276-287"""
        self._bcs = list(bcs) if isinstance(bcs, (list, tuple)) else [bcs]
        self._needs_reassembly = True

    def set_right_hand_side(self, f):
        """hmm.py:289-296"""
        self._f = f

    # ------------------------------------------------------------------ device plumbing
    def _ensure_solver(self):
        if self._solver is not None:
            return
        import torch
        import torch.distributed as dist

        if not torch.cuda.is_available():
            raise native.HmxError("no CUDA device: the HMM hot path runs on the GPU only (there is no CPU fallback)")
        if self._shard and dist.is_available() and dist.is_initialized():
            self._rank, self._world = dist.get_rank(), dist.get_world_size()
        dev = self._device
        if dev is None:
            dev = torch.cuda.current_device()
        self._device = int(dev)
        general = self._micro_tables is not None
        if general and self._cell_solver not in ("auto", "pcg"):
            raise native.HmxError(f"cell_solver='{self._cell_solver}' needs the structured micro mesh")
        mk = lambda variant: native.CellSolver(  # noqa: E731
            self._program, 0 if general else self._structure.n, self._qp, self._qw, rtol=self._cell_rtol, atol=self._cell_atol,
            max_it=self._cell_max_it, device=self._device, collapse=self._collapse, variant=None if general else variant,
            micro_tables=self._micro_tables,
        )  # fmt: skip
        can_direct = not general and native.dense_fits(self._program, self._structure.n, native.collapse_mask(self._program, self._collapse))
        if self._cell_solver == "direct" and not can_direct:
            raise native.HmxError(f"cell_solver='direct' holds at most {native.DENSE_MAX_DOF} unknowns per micro cell")
        forced = {"direct": native.DENSE, "cluster": native.CLUSTER, "pcg": native.MATRIX_FREE}.get(self._cell_solver)
        if self._cell_solver == "cluster":
            self._collapse = False  # the cluster kernel solves the full cell
        self._solver = mk(forced)
        self.cell_solver_used = {native.DENSE: "direct", native.CLUSTER: "cluster", native.ELEMENT_LIST: "pcg (element list)"}.get(
            self._solver.variant, "pcg")  # fmt: skip
        tdev = torch.device("cuda", self._device)
        # device state in LOCAL numbering: the nodes this rank's cells reference, the CSR slots they touch (+ 1 dummy)
        sh = assembly.build_local_shard(self._msh.cells, self._pattern.slot_map, self._pattern.nnz, self._rank, self._world)
        lo, hi = sh.lo, sh.hi
        d = {
            "lo": lo, "hi": hi, "shard": sh, "n_nodes": len(sh.nodes), "nnz": sh.nnz,
            "cells": torch.as_tensor(sh.cells, device=tdev),
            "xyz": torch.as_tensor(np.ascontiguousarray(self._msh.x[sh.nodes]), device=tdev),
            "ptr": torch.as_tensor(sh.gather.ptr, device=tdev),
            "src": torch.as_tensor(sh.gather.src, device=tdev),
            "vals": torch.zeros(sh.nnz + 1, dtype=torch.float64, device=tdev),  # last entry: dummy slot (stays 0)
            "S": torch.zeros((hi - lo, self._num_basis_functions_per_cell**2), dtype=torch.float64, device=tdev),
            "it": torch.zeros(hi - lo, dtype=torch.int32, device=tdev),
            "res": torch.zeros(hi - lo, dtype=torch.float64, device=tdev),
        }  # fmt: skip
        if self._world > 1:
            d["shared"] = torch.as_tensor(sh.shared, device=tdev)
            d["halo"] = parallel.HaloExchange(d["shared"], torch.zeros(len(sh.shared), dtype=torch.float64, device=tdev),
                                              self._solver.halo_pack_dev, self._solver.halo_unpack_dev)
        self._dev = d
        if self._cell_solver == "auto" and can_direct and self.cell_solver_used == "pcg":
            # K5 "only where it beats CG": PCG iterations on a sample of macro cells decide (measured cross-over on
            # B200: 15-40 iterations for 48-192 unknowns, scripts/probe_dense.py).  The sample is spread over the
            # rank's block and the mean is taken over ALL ranks (one all-reduce), so every rank runs the same kernel
            # and a sharded assembly does not depend on how the cells were split.
            idx = parallel.sample_cells(lo, hi)
            m = len(idx)
            its = 0
            if m:
                with torch.cuda.device(self._device):
                    # the buffers below are torch allocations: the probe runs on torch's stream, after them
                    self._solver.set_stream(torch.cuda.current_stream().cuda_stream)
                    cells = d["cells"][torch.as_tensor(idx - lo, device=tdev)].contiguous()
                    self._solver.rhs_iterations(reset=True)
                    self._solver.assemble_macro_dev(m, cells, d["n_nodes"], d["xyz"], 0, None, None, None, d["S"], d["it"], d["res"])
                    self._solver.sync()
                    its = self._solver.rhs_iterations(reset=True)
            mean_it = its / max(m * self._solver.m, 1)
            if self._world > 1:  # (an unsharded solver inside a multi-process job decides on its own)
                mean_it = parallel.agree_on_mean(its, m * self._solver.m, device=tdev)
            if mean_it > DIRECT_ABOVE_ITERATIONS:
                self._solver.close()
                self._solver = mk(native.DENSE)
                self.cell_solver_used = "direct"
                if self._world > 1:
                    d["halo"] = parallel.HaloExchange(d["shared"], d["halo"].buf, self._solver.halo_pack_dev, self._solver.halo_unpack_dev)
            self._logger.info("cell solver: %s (%.1f PCG iterations per right-hand side on %d sample cells)",
                              self.cell_solver_used, mean_it, m)

    # ------------------------------------------------------------------ hot path
    def _assemble_stiffness(self):
        """GPU replacement of hmm.py:298-332 (+ the exchange of hmm.py:442 when sharded)."""
        if not self._needs_reassembly:
            return
        import torch

        self._ensure_solver()
        d = self._dev
        with torch.cuda.device(self._device):
            self._solver.set_stream(torch.cuda.current_stream().cuda_stream)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            if self._macro_qdeg >= 2:
                self._cells_at_rule_points(d)
            else:
                self._solver.assemble_macro_dev(
                    d["hi"] - d["lo"], d["cells"], d["n_nodes"], d["xyz"], 0, None, None, None, d["S"], d["it"], d["res"],
                )  # fmt: skip
            ev[1].record()
            self._solver.gather_csr_dev(d["nnz"], d["ptr"], d["src"], d["S"], d["vals"])
            ev[2].record()
            if self._world > 1:
                self._halo_sum()
            ev[3].record()
            torch.cuda.synchronize()
            n_local = d["hi"] - d["lo"]
            cell_ms = ev[0].elapsed_time(ev[1])
            # observability the reference lacks (it has a tqdm bar, hmm.py:310): device timings of the last assembly
            self.assembly_stats = {
                "macro_cells": n_local, "cell_kernel_ms": cell_ms, "gather_ms": ev[1].elapsed_time(ev[2]),
                "halo_ms": ev[2].elapsed_time(ev[3]), "cell_solves_per_s": n_local / max(cell_ms, 1e-9) * 1e3,
                "rhs_iterations": self._solver.rhs_iterations(reset=True), "rank": self._rank, "world": self._world,
            }  # fmt: skip
            self._logger.info("HMM assembly on cuda:%d: %s", self._device, self.assembly_stats)
            self.cell_iterations = d["it"].cpu().numpy()
            self.cell_residuals = d["res"].cpu().numpy()
            S = d["S"]
            if bool(torch.isnan(S).any()):
                bad = torch.nonzero(torch.isnan(S).any(dim=1)).flatten().cpu().numpy() + d["lo"]
                for c in bad[:10]:  # hmm.py:320-323: logged, not raised
                    self._logger.error(f"Something went wrong when calculating local matrix on cell {c}")
            worst = self.cell_iterations.max(initial=0)
            if worst >= self._cell_max_it:  # hmm.py:427-430
                self._logger.error(f"Cell problem PCG hit ksp_max_it={self._cell_max_it} on {(self.cell_iterations >= self._cell_max_it).sum()} cells")
            self._A_values = self._full_values()
        self._A = sp.csr_matrix((self._A_values, self._pattern.indices, self._pattern.indptr), shape=(self._num_global_dofs,) * 2)
        self._needs_reassembly = False

    def _cells_at_rule_points(self, d):
        """Higher-order macro quadrature: A_hom at the points of the degree-k simplex rule in every local macro cell
        (one launch of the cell kernel over n_cells * nq points), then the element matrices from their weighted mean."""
        import torch

        n = d["hi"] - d["lo"]
        if "rule" not in d:
            pts, wts = quadrature.default_rule(self._tdim, self._macro_qdeg)
            lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)  # barycentric coordinates (nq, d+1)
            d["rule_w"] = np.ascontiguousarray(wts / wts.sum())
            d["rule"] = torch.as_tensor(lam, device=d["xyz"].device)
            nq, m = len(wts), self._solver.m
            d["A_pts"] = torch.empty((n, nq, m, m), dtype=torch.float64, device=d["xyz"].device)
            d["it_pts"] = torch.empty(n * nq, dtype=torch.int32, device=d["xyz"].device)
            d["res_pts"] = torch.empty(n * nq, dtype=torch.float64, device=d["xyz"].device)
        nq = d["rule"].shape[0]
        verts = d["xyz"][d["cells"].long()]  # (n, d+1, 3)
        x = torch.einsum("qa,eak->eqk", d["rule"], verts).reshape(n * nq, 3).contiguous()  # mapped rule points
        self._solver.cell_tensors_dev(n * nq, x, d["A_pts"], d["it_pts"], d["res_pts"])
        self._solver.macro_elements_dev(n, d["cells"], d["xyz"], d["rule_w"], d["A_pts"], d["S"])
        d["it"].copy_(d["it_pts"].view(n, nq).max(dim=1).values)
        d["res"].copy_(d["res_pts"].view(n, nq).max(dim=1).values)

    def _halo_sum(self):
        """Sum of the value slots shared between ranks: the only collective of the path.  The exchange buffer lists
        every globally shared slot; a rank that does not touch one contributes its dummy slot (kept at zero)."""
        vals = self._dev["vals"]
        vals[-1:].zero_()
        self._dev["halo"].sum(vals)

    def _full_values(self):
        """Complete CSR values (global numbering) on every rank for the macro solve; outside the hot path.
        After the halo sum shared slots are complete on every rank that touches them; each slot is contributed by
        the lowest such rank."""
        import torch
        import torch.distributed as dist

        d = self._dev
        sh = d["shard"]
        local = d["vals"][: sh.nnz]
        if self._world == 1:
            v = torch.zeros(self._pattern.nnz, dtype=torch.float64, device=local.device)
            v[torch.as_tensor(sh.slots, device=local.device)] = local
            return v.cpu().numpy()
        if "owned_slots" not in d:
            own = np.nonzero(sh.owned)[0]
            d["owned_local"] = torch.as_tensor(own, device=local.device)
            d["owned_slots"] = torch.as_tensor(sh.slots[own], device=local.device)
        v = torch.zeros(self._pattern.nnz, dtype=torch.float64, device=local.device)
        v[d["owned_slots"]] = local[d["owned_local"]]
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        return v.cpu().numpy()

    def _compute_local_stiffness(self, local_cell_index):
        """One macro cell (the finer seam, hmm.py:334-369): returns the (n_b, n_b) local matrix (barycentre rule, as
        the reference; the macro_quadrature_degree option acts in _assemble_stiffness)."""
        self._ensure_solver()
        nb = self._num_basis_functions_per_cell
        cells = np.ascontiguousarray(self._msh.cells[local_cell_index : local_cell_index + 1], dtype=np.int32)
        gp = np.arange(nb * nb + 1, dtype=np.int64)
        gs = np.arange(nb * nb, dtype=np.int32)
        _, S = self._solver.assemble_macro(cells, self._msh.x, gp, gs, want_local=True)
        return S[0]

    def cell_tensors(self, points):
        """Homogenised tensors A_hom at arbitrary macro points, shape (n, m, m)."""
        self._ensure_solver()
        return self._solver.cell_tensors(points)

    # ------------------------------------------------------------------ solve (hmm.py:434-491)
    def _assemble_load_device(self):
        """Macro load vector b_i = int f . phi_i dx on the device (hmm.py:131-133, 445-450): f is
        compiled into a small kernel (csrc/hmx_load_entry.cu), the element vectors are summed into b by the same
        deterministic gather as the stiffness values.  Returns a device tensor (every rank assembles all of b: it is
        outside the hot path and costs microseconds)."""
        import torch

        d = self._dev
        tdev = d["vals"].device
        lprog = codegen.build_load_program(self._f, self._tdim, self._bs)
        image = native.compile_load_kernel(lprog)
        qp, qw = quadrature.default_rule(self._tdim, lprog.degree)
        if "b_ptr" not in d:
            gm = assembly.build_gather(assembly.unroll_dofs(self._msh.cells, self._bs), self._num_global_dofs)
            d["b_ptr"], d["b_src"] = torch.as_tensor(gm.ptr, device=tdev), torch.as_tensor(gm.src, device=tdev)
            d["cells_all"] = torch.as_tensor(np.ascontiguousarray(self._msh.cells, dtype=np.int32), device=tdev)
            d["xyz_all"] = torch.as_tensor(np.ascontiguousarray(self._msh.x), device=tdev)
        nb = self._num_basis_functions_per_cell
        with torch.cuda.device(self._device):
            self._solver.set_stream(torch.cuda.current_stream().cuda_stream)
            Fe = torch.empty((self._msh.num_cells, nb), dtype=torch.float64, device=tdev)
            b = torch.empty(self._num_global_dofs, dtype=torch.float64, device=tdev)
            self._solver.macro_load_dev(image, self._msh.num_cells, d["cells_all"], d["xyz_all"], qp, qw, Fe)
            self._solver.gather_csr_dev(self._num_global_dofs, d["b_ptr"], d["b_src"], Fe, b)
        return b

    def solve(self):
        self._assemble_stiffness()
        opts = self._petsc_options_global_solve
        direct = opts.get("pc_type") == "lu" or opts.get("ksp_type") == "preonly" or opts.get("hmx_macro_solver") == "host"
        if not direct:
            x = self._solve_on_device(self._assemble_load_device())
            if x is not None:
                self._u.x.array[:] = x
                return self._u
        b = fem.assemble_load(self._V_macro, self._f)  # host path (sparse LU requested, or the device PCG did not converge)
        A = self._A.copy().tocsr()
        for bc in self._bcs:  # Dirichlet lifting, one condition at a time as in hmm.py:453-480
            u_bc = np.zeros(self._num_global_dofs)
            u_bc[bc.dofs] = bc.values
            b = b - A @ u_bc
            keep = np.ones(self._num_global_dofs)
            keep[bc.dofs] = 0.0
            Dk = sp.diags(keep)
            A = (Dk @ A @ Dk + sp.diags(1.0 - keep)).tocsr()
            b[bc.dofs] = bc.values
        self._b = b
        # macro solve: PETSc KSP in the reference (hmm.py:482-483, default GMRES + ILU); scipy here
        x = _macro_solve(A, b, opts)
        if not np.all(np.isfinite(x)):  # hmm.py:485-488: logged, not raised
            self._logger.error("Something went wrong in the global problem solve.")
        self._u.x.array[:] = x
        return self._u

    def _solve_on_device(self, b):
        """Dirichlet lifting (hmm.py:453-480) and the macro Krylov solve (hmm.py:482-483) on the
        GPU: `hmx_macro_lift_dev` + Jacobi-PCG `hmx_macro_pcg_dev` on the CSR values the assembly left on the device.
        Conditions are lifted one at a time like the reference does (so a dof named by two conditions behaves as it
        does there: the second lifting sees the rows and columns the first one already zeroed).  Returns None if the PCG does not converge (the caller then falls back to the host solve,
        which also handles non-symmetric coefficients)."""
        import torch

        opts = self._petsc_options_global_solve
        n = self._num_global_dofs
        d = self._dev
        tdev = d["vals"].device
        if "indptr" not in d:
            d["indptr"] = torch.as_tensor(self._pattern.indptr, device=tdev)
            d["indices"] = torch.as_tensor(self._pattern.indices, device=tdev)
        with torch.cuda.device(self._device):
            vals = torch.as_tensor(self._A_values, device=tdev).clone()  # complete values (all ranks after the halo sum)
            t_b = b.clone() if hasattr(b, "data_ptr") else torch.as_tensor(np.ascontiguousarray(b), device=tdev)
            t_x = torch.zeros(n, dtype=torch.float64, device=tdev)
            self._solver.set_stream(torch.cuda.current_stream().cuda_stream)
            for bc in self._bcs:  # one condition at a time, each lifting with the matrix the previous ones left
                mask = np.zeros(n, dtype=np.int8)
                ubc = np.zeros(n)
                mask[bc.dofs] = 1
                ubc[bc.dofs] = bc.values
                t_mask, t_ubc = torch.as_tensor(mask, device=tdev), torch.as_tensor(ubc, device=tdev)
                self._solver.macro_lift_dev(n, d["indptr"], d["indices"], vals, t_mask, t_ubc, t_b)
            it, res = self._solver.macro_pcg_dev(
                n, d["indptr"], d["indices"], vals, t_b, t_x, rtol=float(opts.get("ksp_rtol", 1e-12)),
                atol=float(opts.get("ksp_atol", 0.0)), max_it=int(opts.get("ksp_max_it", 20000)),
            )  # fmt: skip
            self.macro_solve_stats = {"iterations": it, "relative_residual": res}
            x = t_x.cpu().numpy()
            self._b = t_b.cpu().numpy()
        if res > max(float(opts.get("ksp_rtol", 1e-12)), 1e-10) * 10 or not np.all(np.isfinite(x)):
            self._logger.error(f"macro PCG stopped at relative residual {res:.2e} after {it} iterations; using the host solver")
            return None
        return x

    def plot_solution(self, u=None):
        """hmm.py:493-511 (needs pyvista, as in the reference)."""
        import pyvista as pv  # noqa: F401

        raise NotImplementedError("plotting is outside the hot path; use the mesh and u.x.array with pyvista directly")


# ----------------------------------------------------------------------------------------------
class PoissonHMM(BaseHMM):
    """-div(A(x, x/eps) grad u) = f with zero Dirichlet data on the bounding box by default
    (hmm.py:514-667)."""

    _KIND = codegen.POISSON

    def __init__(self, msh, A, f, msh_micro, eps, petsc_options_global_solve=None, petsc_options_cell_problem=None,
                 petsc_options_prefix="hommx_PoissonHMM", **kw):  # fmt: skip
        super().__init__(msh, A, f, msh_micro, eps, petsc_options_global_solve, petsc_options_cell_problem,
                         petsc_options_prefix, **kw)  # fmt: skip
        dofs = fem.boundary_nodes(self._msh)  # hmm.py:598-636
        self._bcs = [fem.dirichletbc(0.0, dofs, self._V_macro)]


class PoissonStratifiedHMM(BaseHMM):
    """Stratified micro structure A(x, theta(x)/eps); ``Dtheta_transpose(x)[p][i] = d theta_i / d x_p``
    (hmm.py:670-789).  No default boundary condition, as in the reference."""

    _KIND = codegen.POISSON

    def __init__(self, msh, A, f, msh_micro, eps, Dtheta_transpose, petsc_options_global_solve=None,
                 petsc_options_cell_problem=None, petsc_options_prefix="hommx_PoissonStratifiedHMM", **kw):  # fmt: skip
        super().__init__(msh, A, f, msh_micro, eps, petsc_options_global_solve, petsc_options_cell_problem,
                         petsc_options_prefix, Dtheta_transpose=Dtheta_transpose, **kw)  # fmt: skip


class LinearElasticityHMM(BaseHMM):
    """-div(A(x, x/eps) : e(u)) = f, A a rank-4 tensor; no default boundary condition (hmm.py:792-922)."""

    _KIND = codegen.ELASTICITY

    def __init__(self, msh, A, f, msh_micro, eps, petsc_options_global_solve=None, petsc_options_cell_problem=None,
                 petsc_options_prefix="hommx_LinearElasticityHMM", **kw):  # fmt: skip
        super().__init__(msh, A, f, msh_micro, eps, petsc_options_global_solve, petsc_options_cell_problem,
                         petsc_options_prefix, **kw)  # fmt: skip


class LinearElasticityStratifiedHMM(BaseHMM):
    """Stratified linear elasticity (hmm.py:925-1067; the reference's default prefix repeats
    "hommx_LinearElasticityHMM", hmm.py:986 -- kept)."""

    _KIND = codegen.ELASTICITY

    def __init__(self, msh, A, f, msh_micro, eps, Dtheta_transpose, petsc_options_global_solve=None,
                 petsc_options_cell_problem=None, petsc_options_prefix="hommx_LinearElasticityHMM", **kw):  # fmt: skip
        super().__init__(msh, A, f, msh_micro, eps, petsc_options_global_solve, petsc_options_cell_problem,
                         petsc_options_prefix, Dtheta_transpose=Dtheta_transpose, **kw)  # fmt: skip


# ----------------------------------------------------------------------------------------------
class PoissonPeriodicHMM:
    """Periodic homogenisation for scalar diffusion, ``A = A(y)`` (hmm.py:1070-1279): one cell
    problem per direction gives the constant tensor ``A_hom`` (``compute_effective_tensor``,
    hmm.py:1219-1245) -- the one-point special case of the HMM cell kernel -- followed by a
    constant-coefficient macro problem (hmm.py:1247-1255).  No default boundary condition
    (hmm.py:1132).  ``A_hom`` and ``correctors`` (hmm.py:1211-1217) come from ``hmx_cell_correctors_dev``."""

    def __init__(self, msh, A, f, msh_micro, eps, petsc_options_global_solve=None, petsc_options_cell_problem=None,
                 petsc_options_prefix="hommx_periodicHMM", *, quadrature_rule=None, device=None):  # fmt: skip
        self._logger = logging.getLogger(__name__)
        self._msh = as_simplex_mesh(msh)
        self._cell_mesh = as_simplex_mesh(msh_micro)
        self._tdim = self._cell_mesh.dim
        if self._tdim not in (2, 3):
            raise ValueError("Only 2D and 3D periodic homogenization supported.")  # hmm.py:1098-1099
        self._coeff, self._f, self._eps = A, f, eps
        if petsc_options_cell_problem is None:
            petsc_options_cell_problem = {"ksp_atol": 1e-12}  # hmm.py:1102-1104
        self._petsc_options_cell_problem = dict(petsc_options_cell_problem)
        self._petsc_options_global_solve = petsc_options_global_solve
        self._V_macro = fem.FunctionSpace(self._msh, 1)
        self._u = fem.Function(self._V_macro)
        self._bcs = []
        self._A_hom = None
        self._correctors = None
        # the HMM machinery with a coefficient that ignores the macro point
        def A_xy(x, y):
            return A(y)

        A_xy.__wrapped__ = A  # lets the tracer find the UFL module A is written against
        self._hmm = BaseHMM(self._msh, A_xy, f, self._cell_mesh, eps, petsc_options_global_solve,
                            {"ksp_rtol": float(self._petsc_options_cell_problem.get("ksp_rtol", 1e-10)),
                             "ksp_atol": float(self._petsc_options_cell_problem.get("ksp_atol", 1e-12)),
                             "ksp_max_it": int(self._petsc_options_cell_problem.get("ksp_max_it", 10000))},
                            petsc_options_prefix, quadrature_rule=quadrature_rule, device=device, shard=False)  # fmt: skip

    @property
    def function_space(self):
        return self._V_macro

    def set_boundary_conditions(self, bcs):
        self._bcs = list(bcs) if isinstance(bcs, (list, tuple)) else [bcs]

    def set_right_hand_side(self, f):
        self._f = f

    @property
    def A_hom(self):
        return self._A_hom

    @property
    def correctors(self):
        """One ``fem.Function`` on the micro mesh per direction (hmm.py:1211-1213), slaves of the periodic
        constraint filled with their master's value; each is defined up to an additive constant."""
        if self._correctors is None:
            self.compute_effective_tensor()
        return self._correctors

    def compute_effective_tensor(self):
        """One GPU cell solve (all directions at once); returns the (d, d) tensor."""
        self._hmm._ensure_solver()
        A, chi = self._hmm._solver.cell_correctors(np.zeros((1, 3)))
        self._A_hom = A[0]
        d = self._tdim
        if self._hmm._micro_tables is not None:  # general mesh: values on the periodic nodes
            idx = (self._hmm._micro_tables.node2per,)
        else:
            n = self._hmm._structure.n
            ij = np.rint(self._cell_mesh.x[:, :d] * n).astype(np.int64) % n  # vertex -> periodic grid index
            idx = tuple(ij[:, a] for a in reversed(range(d)))  # grid arrays are (z, y, x)
        V = fem.FunctionSpace(self._cell_mesh, 1)
        self._correctors = []
        for q in range(d):
            fn = fem.Function(V)
            fn.x.array[:] = chi[0, q, 0][idx]
            self._correctors.append(fn)
        return self._A_hom

    def solve(self):
        if self._A_hom is None:
            self.compute_effective_tensor()
        msh, d = self._msh, self._tdim
        pat = self._hmm._pattern
        vals = np.zeros(pat.nnz)
        X = msh.x[:, :d]
        for c, nodes in enumerate(msh.cells):  # constant-coefficient P1 stiffness (host, outside the hot path)
            v = X[nodes]
            J = (v[1:] - v[0]).T
            Ji = np.linalg.inv(J)
            G = np.concatenate([-Ji.sum(axis=0, keepdims=True), Ji], axis=0).T  # (d, d+1)
            S = abs(np.linalg.det(J)) / (2 if d == 2 else 6) * (G.T @ self._A_hom.T @ G)
            np.add.at(vals, pat.slot_map[c], S.ravel())
        A = sp.csr_matrix((vals, pat.indices, pat.indptr), shape=(pat.n_dofs,) * 2)
        self._A = A.copy()
        b = fem.assemble_load(self._V_macro, self._f)
        n = pat.n_dofs
        for bc in self._bcs:
            u_bc = np.zeros(n)
            u_bc[bc.dofs] = bc.values
            b = b - A @ u_bc
            keep = np.ones(n)
            keep[bc.dofs] = 0.0
            Dk = sp.diags(keep)
            A = (Dk @ A @ Dk + sp.diags(1.0 - keep)).tocsr()
            b[bc.dofs] = bc.values
        self._u.x.array[:] = _macro_solve(A, b, self._petsc_options_global_solve or {})
        return self._u
