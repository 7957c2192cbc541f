import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import cases as K
from hommx_b200 import native
for name in ("e3_fibre_rot_n4", "e3_fibre_rot_n8_c4"):
    case = K.BY_NAME[name]; prog = K.program(case); qp,qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-10, variant=2)
    x = K.points(case, 5)
    A, it, res = s.cell_tensors(x, True)
    mic = K.oracle_cell(case, prog)
    err = max(np.abs(A[k]-K.oracle_tensor(case, mic, x[k])).max()/np.abs(A[k]).max() for k in range(len(x)))
    print(name, s.info, it.tolist(), err, flush=True)
    s.close()
