"""PoissonHMM in 2-D -- the set-up of the reference's examples/hmm.py (BASELINE config 1):
A(x, y) = 1.1 + x0 + sin(2 pi y0) on [0,5]^2, u = 1 on the left edge, 0 on the right, f = 0.

    python examples/hmm_2d.py [--macro 32] [--micro 16]
"""
import argparse

import numpy as np
from _common import report, timed_solve

from hommx_b200 import PoissonHMM, fem, mesh, ufl

ap = argparse.ArgumentParser()
ap.add_argument("--macro", type=int, default=32)
ap.add_argument("--micro", type=int, default=16)
args = ap.parse_args()

eps = 1 / 2**5


def A(x, y):
    return 1.1 + x[0] + ufl.sin(2 * ufl.pi * y[0])


msh = mesh.create_rectangle((0.0, 0.0), (5.0, 5.0), (args.macro, args.macro))
msh_micro = mesh.create_unit_square(args.micro, args.micro)
phmm = PoissonHMM(msh, A, lambda x: 0.0, msh_micro, eps, petsc_options_cell_problem={"ksp_atol": 1e-9})
V = phmm.function_space
left = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
right = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 5.0))
phmm.set_boundary_conditions([fem.dirichletbc(1.0, left, V), fem.dirichletbc(0.0, right, V)])
u, ta, ts = timed_solve(phmm)
report("PoissonHMM 2D", phmm, u, ta, ts)
