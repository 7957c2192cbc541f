"""One launch shape of the cluster kernel for ncu (python scripts/profile_cluster.py [case] [npts]); three launches, profile the last:
   ncu --set full --import-source on --clock-control none -k hmx_cell --launch-skip 2 -c 1 -o out python scripts/profile_cluster.py"""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch

import cases as K
from hommx_b200 import native

name = sys.argv[1] if len(sys.argv) > 1 else "e3_fibre_rot_n8_c4"
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 8
case = K.BY_NAME[name]
prog = K.program(case)
qp, qw = K.tables(case, prog)
xd = torch.tensor(K.points(case, npts), device="cuda")
A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device="cuda", dtype=torch.float64)
s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, atol=1e-10, variant=native.CLUSTER)
for _ in range(3):
    s.cell_tensors_dev(npts, xd, A)
    s.sync()
print("ok", s.info)
