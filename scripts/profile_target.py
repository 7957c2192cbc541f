"""One launch of a named parity case for ncu (python scripts/profile_target.py <case> <n> <npts>)."""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native, micro, quadrature
name, n, npts = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
case = K.BY_NAME[name]; prog = K.program(case)
st = micro.default_structure(case.dim, n); qp, qw = micro.quadrature_table(st, *quadrature.default_rule(case.dim, prog.degree))
rng = np.random.default_rng(0); x = rng.uniform(0,1,(npts,3))
if case.dim==2: x[:,2]=0
xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device='cuda', dtype=torch.float64)
s = native.CellSolver(prog, n, qp, qw, rtol=1e-8)
for _ in range(3):
    s.cell_tensors_dev(npts, xd, A); s.sync()
print("ok", s.info)
