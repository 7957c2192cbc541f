"""Dump golden vectors of the UNMODIFIED reference (flxrcz/hommx 0.0.3 on DOLFINx 0.9 / dolfinx_mpc 0.9.3 /
PETSc 3.23) -> tests/golden/reference_vectors.json.

This script CANNOT run in the build image (dolfinx, ufl, basix, dolfinx_mpc, petsc4py, mpi4py are not
installable there and the reference needs Python >= 3.13); it is committed so that anybody with a DOLFINx
environment (e.g. ``pixi run`` in a checkout of the reference) can pin this repository against the reference:

    python tests/golden/make_reference_golden.py            # writes tests/golden/reference_vectors.json
    python -m pytest tests/test_reference_golden.py          # oracle (CPU) and CUDA path (-m gpu) against the file

Everything dumped is NUMBERING-INVARIANT (DOLFINx reorders cells and dofs): local matrices come with the vertex
coordinates of their rows, CSR rows/columns with their dof coordinates, tensors with the macro point.  What is
called, all through the reference's public classes and its own code path:

* ``BaseHMM._compute_local_stiffness(cell)``      (src/hommx/hmm.py:334-369)  -> S_loc per macro cell
* ``BaseHMM._assemble_stiffness()`` + ``_A.assemble()`` (hmm.py:298-332, 442) -> macro CSR pattern + values
* ``PoissonPeriodicHMM.compute_effective_tensor()`` (hmm.py:1219-1245)        -> A_hom
* ``basix.make_quadrature(cell, degree)``          (the rule FFCx integrates the forms with)
* ``dolfinx.mesh.create_unit_square / create_unit_cube`` cell lists          (the Kuhn split)

Cell problems are solved with LU (MUMPS), as the reference's tightest test does
(test/integration/test_integration_poisson.py:207-211), so that the vectors are good to ~1e-13.
The coefficient callables are tests/coefficients.py evaluated with the real ``ufl`` module (the same source text
the oracle evaluates with numpy and the product traces into CUDA).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import coefficients as Cf  # noqa: E402

LU = {"ksp_type": "preonly", "pc_type": "lu", "pc_factor_mat_solver_type": "mumps"}

# name: (class, dim, micro n, coefficient, Dtheta, macro box, macro cells)  -- C1..C4 of BASELINE.json with their real
# micro cells on scaled-down macro meshes (S_loc of a macro cell does not depend on the rest of the mesh), plus the
# stratified / tensor-valued cases of tests/cases.py that the reference API can express
CASES = {
    "c1_p2_smooth_n16": ("PoissonHMM", 2, 16, "smooth_sin", None, ((0.0, 0.0), (1.0, 1.0)), (4, 4)),
    "c2_p2_laminate_wavy_n32": ("PoissonStratifiedHMM", 2, 32, "laminate", "dtheta_wavy", ((0.0, 0.0), (1.0, 1.0)), (4, 4)),
    "c3_p3_smooth_n8": ("PoissonHMM", 3, 8, "smooth_sin", None, ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), (2, 2, 2)),
    "c4_e3_fibre_rot_n8": ("LinearElasticityStratifiedHMM", 3, 8, "hooke_fibre_3d", "dtheta_rotation_3d",
                           ((0.0, 0.0, 0.0), (1.0, 0.4, 0.1)), (4, 2, 1)),
    "p2_analytic1_n15": ("PoissonHMM", 2, 15, "analytic1", None, ((0.0, 0.0), (1.0, 1.0)), (3, 3)),
    "p2_inclusion_n16": ("PoissonStratifiedHMM", 2, 16, "inclusion", "dtheta_inclusion", ((0.0, 0.0), (1.0, 1.0)), (3, 3)),
    "p2_smooth_strat_n12": ("PoissonStratifiedHMM", 2, 12, "smooth_sin", "dtheta_test_stratified", ((0.0, 0.0), (1.0, 1.0)), (3, 3)),
    "e2_hooke_sin_n6": ("LinearElasticityHMM", 2, 6, "hooke_sin_2d", None, ((0.0, 0.0), (1.0, 0.3)), (4, 2)),
    "e3_hooke_const_n3": ("LinearElasticityHMM", 3, 3, "hooke_const_3d", None, ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), (2, 2, 2)),
    "e3_hooke_smooth_shear_n4": ("LinearElasticityStratifiedHMM", 3, 4, "hooke_smooth_3d", "dtheta_shear_3d",
                                 ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)), (2, 1, 1)),
}  # fmt: skip
PERIODIC = {"periodic_only_n8": (2, 8, "periodic_only"), "periodic_only_3d_n6": (3, 6, "periodic_only")}
EPS = 2.0**-6


def main(out_path):
    import basix
    import dolfinx
    import ufl
    from dolfinx import mesh
    from mpi4py import MPI

    import hommx
    from hommx import hmm as ref

    comm = MPI.COMM_SELF
    out = {"versions": {"dolfinx": dolfinx.__version__, "basix": basix.__version__, "ufl": ufl.__version__,
                        "hommx": getattr(hommx, "__version__", "?")},
           "cases": {}, "periodic": {}, "quadrature": {}, "meshes": {}}  # fmt: skip

    def make_mesh(dim, box, cells):
        if dim == 2:
            return mesh.create_rectangle(comm, [np.array(box[0]), np.array(box[1])], list(cells), mesh.CellType.triangle)
        return mesh.create_box(comm, [np.array(box[0]), np.array(box[1])], list(cells), mesh.CellType.tetrahedron)

    def unit_cell(dim, n):
        return mesh.create_unit_square(comm, n, n) if dim == 2 else mesh.create_unit_cube(comm, n, n, n)

    for name, (cls, dim, n, coeff, dth, box, cells) in CASES.items():
        msh, mic = make_mesh(dim, box, cells), unit_cell(dim, n)
        A = getattr(Cf, coeff)(ufl)
        f = (lambda x: 1.0) if "Poisson" in cls else (lambda x: ufl.as_vector([0.0] * (dim - 1) + [-1.0]))
        args = (msh, A, f, mic, EPS) + ((getattr(Cf, dth)(ufl),) if dth else ())
        solver = getattr(ref, cls)(*args, petsc_options_cell_problem=LU, petsc_options_global_solve={"ksp_type": "preonly", "pc_type": "lu"})
        # per-cell local matrices through the reference's own seam
        solver._setup_cell_problem_forms()
        nc = msh.topology.index_map(dim).size_local
        bs = solver._bs
        rec = {"class": cls, "dim": dim, "n_micro": n, "coeff": coeff, "dtheta": dth, "eps": EPS, "bs": int(bs), "cells": []}
        for c in range(nc):
            dofs = solver._V_macro.dofmap.cell_dofs(c)
            rec["cells"].append({"vertices": solver._macro_coordinates[dofs].tolist(),  # row/column a*bs+k <-> vertex a, component k
                                 "S_loc": solver._compute_local_stiffness(c).tolist()})
        # assembled macro matrix before boundary conditions (hmm.py:298-332, 442)
        solver._needs_reassembly = True
        solver._assemble_stiffness()
        solver._A.assemble()
        indptr, indices, data = solver._A.getValuesCSR()
        rec["csr"] = {"indptr": np.asarray(indptr).tolist(), "indices": np.asarray(indices).tolist(), "data": np.asarray(data).tolist(),
                      "dof_coordinates": solver._macro_coordinates.tolist()}  # unrolled dof d <-> node d // bs, component d % bs
        out["cases"][name] = rec
        print(name, "done:", nc, "macro cells")

    for name, (dim, n, coeff) in PERIODIC.items():
        msh, mic = make_mesh(dim, ((0.0,) * dim, (1.0,) * dim), (2,) * dim), unit_cell(dim, n)
        Axy = getattr(Cf, coeff)(ufl)
        per = ref.PoissonPeriodicHMM(msh, lambda y: Axy(None, y), lambda x: 1.0, mic, EPS, petsc_options_cell_problem=LU,
                                     petsc_options_global_solve={"ksp_type": "preonly", "pc_type": "lu"})
        out["periodic"][name] = {"dim": dim, "n_micro": n, "coeff": coeff, "A_hom": np.asarray(per.compute_effective_tensor()).tolist()}
        print(name, "done")

    # the quadrature rules FFCx would pick (basix default scheme), degree 0..4
    for cell in ("triangle", "tetrahedron"):
        for deg in range(5):
            pts, wts = basix.make_quadrature(getattr(basix.CellType, cell), deg)
            out["quadrature"][f"{cell}_{deg}"] = {"points": np.asarray(pts).tolist(), "weights": np.asarray(wts).tolist()}

    # how DOLFINx splits the unit square / cube: cells as sorted vertex-coordinate tuples
    for key, m in (("unit_square_3", mesh.create_unit_square(comm, 3, 3)), ("unit_cube_2", mesh.create_unit_cube(comm, 2, 2, 2))):
        tdim = m.topology.dim
        m.topology.create_connectivity(tdim, 0)
        conn = m.topology.connectivity(tdim, 0)
        x = m.geometry.x
        # geometry and topology vertex numbering coincide for P1 meshes created by these generators
        cells = [sorted(map(tuple, np.round(x[conn.links(c)], 12).tolist())) for c in range(m.topology.index_map(tdim).size_local)]
        out["meshes"][key] = sorted(cells)

    with open(out_path, "w") as fh:
        json.dump(out, fh)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "reference_vectors.json"))
