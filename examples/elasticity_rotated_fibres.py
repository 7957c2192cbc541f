"""LinearElasticityStratifiedHMM -- the set-up of the reference's examples/linear_elasticity/rotated_fibers.py
(BASELINE config 4): beam [0,1] x [0,0.4] x [0,0.1] clamped at x0 = 0 under its own weight, stiff fibres
(mu = 100 in mu = 0.001, lambda = 1) whose direction rotates about the x1 axis by gamma(x1) = pi x1 / 0.8.

    python examples/elasticity_rotated_fibres.py [--macro 20 6 6] [--micro 8] [--full-cell]
"""
import argparse

import numpy as np
from _common import report, timed_solve

from hommx_b200 import LinearElasticityStratifiedHMM, fem, mesh, ufl

ap = argparse.ArgumentParser()
ap.add_argument("--macro", type=int, nargs=3, default=[20, 6, 6])
ap.add_argument("--micro", type=int, default=8)
ap.add_argument("--full-cell", action="store_true", help="solve the full n^3 cell (no axis collapse along the fibres)")
args = ap.parse_args()
L, W, H = 1.0, 0.4, 0.1


def in_fibre(a, b, r=0.25):
    dx = ufl.acos(ufl.cos(2 * ufl.pi * (a - 1 / 2)))
    dy = ufl.acos(ufl.cos(2 * ufl.pi * (b - 1 / 2)))
    return (dx**2 + dy**2) < ((2 * ufl.pi) ** 2 * r**2)


def A(x, y):
    mu = ufl.conditional(in_fibre(y[1], y[2]), 100, 0.001)
    lam = 1
    I = ufl.Identity(3)
    i, j, k, l = ufl.indices(4)
    return ufl.as_tensor(lam * I[i, j] * I[k, l] + mu * (I[i, k] * I[j, l] + I[i, l] * I[j, k]), indices=(i, j, k, l))


def Dtheta_transpose(x):
    g = 1 / 2 * ufl.pi * x[1] / W
    R = ufl.as_matrix([[ufl.cos(g), 0.0, -ufl.sin(g)], [0.0, 1.0, 0.0], [ufl.sin(g), 0.0, ufl.cos(g)]])
    return ufl.transpose(R)


msh = mesh.create_box((0.0, 0.0, 0.0), (L, W, H), args.macro)
msh_micro = mesh.create_unit_cube(args.micro, args.micro, args.micro)
hmm = LinearElasticityStratifiedHMM(msh, A, lambda x: ufl.as_vector([0.0, 0.0, -0.05 * W**2]), msh_micro, 0.01, Dtheta_transpose,
                                    collapse_invariant_axes=not args.full_cell)  # fmt: skip
V = hmm.function_space
clamp = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
hmm.set_boundary_conditions(fem.dirichletbc(np.zeros(3), clamp, V))
u, ta, ts = timed_solve(hmm)
report("LinearElasticityStratifiedHMM rotated fibres", hmm, u, ta, ts)
tip = u.x.array.reshape(-1, 3)[np.isclose(msh.x[:, 0], L)]
print(f"mean tip deflection: {tip[:, 2].mean():.6g}")
