"""PoissonStratifiedHMM on a wavy laminate -- the set-up of the reference's examples/diffusion/laminate.py
(BASELINE config 2): layers a = 5 / 0.05 along theta_0(x) = x1 - sin(2 pi x0), completed to the square map
theta(x) = (x1 - sin 2 pi x0, x0), u = 1 left, 0 right, f = 1.

    python examples/diffusion_laminate.py [--macro 64] [--micro 32]
"""
import argparse

import numpy as np
from _common import report, timed_solve

from hommx_b200 import PoissonStratifiedHMM, fem, mesh, ufl

ap = argparse.ArgumentParser()
ap.add_argument("--macro", type=int, default=64)
ap.add_argument("--micro", type=int, default=32)
args = ap.parse_args()


def A(x, y):
    return ufl.conditional(ufl.cos(2 * ufl.pi * y[0]) < 0, 5, 0.05)


def Dtheta_transpose(x):
    return ufl.as_matrix([[-2 * ufl.pi * ufl.cos(2 * ufl.pi * x[0]), 1.0], [1.0, 0.0]])


msh = mesh.create_unit_square(args.macro, args.macro)
msh_micro = mesh.create_unit_square(args.micro, args.micro)
hmm = PoissonStratifiedHMM(msh, A, lambda x: 1.0, msh_micro, 1e-5, Dtheta_transpose)
V = hmm.function_space
left = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
right = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 1.0))
hmm.set_boundary_conditions([fem.dirichletbc(1.0, left, V), fem.dirichletbc(0.0, right, V)])
u, ta, ts = timed_solve(hmm)
report("PoissonStratifiedHMM laminate", hmm, u, ta, ts)
