import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
import cases as K
from hommx_b200 import native
npts = int(sys.argv[1])
case = K.BY_NAME["e3_fibre_rot_n8_c4"]; prog = K.program(case); qp,qw = K.tables(case, prog)
rng = np.random.default_rng(0); x = rng.uniform(0,1,(npts,3)); x[:,1]*=0.4; x[:,2]*=0.1
xd = torch.tensor(x, device='cuda'); A = torch.empty((npts, 6, 6), device='cuda', dtype=torch.float64)
s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, collapse=True, variant=(native.DENSE if len(sys.argv) > 2 and sys.argv[2] == "dense" else None))
for _ in range(2):
    s.cell_tensors_dev(npts, xd, A); s.sync()
print("ok", s.info)
