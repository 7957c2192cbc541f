"""PoissonHMM with a GENERAL periodic micro mesh (not in the reference's examples; its periodic constraint accepts any
micro mesh whose boundary nodes match, cell_problem.py:16-35): the set-up of examples/hmm_2d.py with the micro nodes
moved off the grid.  The structured kernels do not apply; the drop-in class switches to the element-list kernel
(csrc/hmx_cell_generic.cuh) by itself.

    python examples/general_micro_mesh.py [--macro 16] [--micro 12]
"""
import argparse

import numpy as np
from _common import report, timed_solve

from hommx_b200 import PoissonHMM, fem, mesh, ufl

ap = argparse.ArgumentParser()
ap.add_argument("--macro", type=int, default=16)
ap.add_argument("--micro", type=int, default=12)
args = ap.parse_args()


def A(x, y):
    return 1.1 + x[0] + ufl.sin(2 * ufl.pi * y[0])


def perturbed_unit_square(n, amp=0.3, seed=0):
    """create_unit_square(n, n) with every interior node moved by up to amp * h / 2 (boundary nodes stay: they match)."""
    m = mesh.create_unit_square(n, n)
    x = m.x.copy()
    ij = np.rint(x[:, :2] * n).astype(np.int64)
    inside = ((ij > 0) & (ij < n)).all(axis=1)
    x[inside, :2] += np.random.default_rng(seed).uniform(-0.5, 0.5, (int(inside.sum()), 2)) * amp / n
    return mesh.SimplexMesh(x, m.cells.copy(), 2)


msh = mesh.create_rectangle((0.0, 0.0), (5.0, 5.0), (args.macro, args.macro))
phmm = PoissonHMM(msh, A, lambda x: 0.0, perturbed_unit_square(args.micro), 1 / 2**5, petsc_options_cell_problem={"ksp_atol": 1e-9})
V = phmm.function_space
left = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
right = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 5.0))
phmm.set_boundary_conditions([fem.dirichletbc(1.0, left, V), fem.dirichletbc(0.0, right, V)])
u, ta, ts = timed_solve(phmm)
report(f"PoissonHMM 2D, general micro mesh [{phmm.cell_solver_used}]", phmm, u, ta, ts)
