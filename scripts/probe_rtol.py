import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import cases as K
from hommx_b200 import native
case = K.BY_NAME["e3_fibre_rot_n8_c4"]; prog = K.program(case); qp,qw = K.tables(case, prog)
rng = np.random.default_rng(5); x = rng.uniform(0,1,(296,3)); x[:,1]*=0.4; x[:,2]*=0.1
s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-13, atol=1e-14)
ref, it, res = s.cell_tensors(x, True); print("ref its", it.mean())
for rtol in (1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-9, 1e-10):
    s.set_tolerances(rtol, 1e-14)
    A, it, res = s.cell_tensors(x, True)
    err = np.abs(A-ref).reshape(len(x),-1).max(axis=1)/np.abs(ref).reshape(len(x),-1).max(axis=1)
    print(f"rtol {rtol:.0e}: mean its {it.mean():.1f} max its {it.max()}  A_hom rel err max {err.max():.2e} median {np.median(err):.2e}")
