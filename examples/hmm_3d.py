"""PoissonHMM in 3-D -- the set-up of the reference's examples/hmm_3d.py (BASELINE config 3):
A(x, y) = 1.1 + x0 + sin(2 pi y0), f = 1, zero Dirichlet data on the boundary of the unit cube.

    python examples/hmm_3d.py [--macro 16] [--micro 8]
"""
import argparse

from _common import report, timed_solve

from hommx_b200 import PoissonHMM, mesh, ufl

ap = argparse.ArgumentParser()
ap.add_argument("--macro", type=int, default=16)
ap.add_argument("--micro", type=int, default=8)
args = ap.parse_args()


def A(x, y):
    return 1.1 + x[0] + ufl.sin(2 * ufl.pi * y[0])


msh = mesh.create_unit_cube(args.macro, args.macro, args.macro)
msh_micro = mesh.create_unit_cube(args.micro, args.micro, args.micro)
phmm = PoissonHMM(msh, A, lambda x: 1.0, msh_micro, 1 / 2**3)
u, ta, ts = timed_solve(phmm)
report("PoissonHMM 3D", phmm, u, ta, ts)
