"""Tiny run of every kernel family for compute-sanitizer (memcheck / racecheck), one tool per call."""
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import cases as K
from hommx_b200 import native

for name, kw in [("p2_inclusion_n16", {}), ("p3_smooth_n4", {}), ("e3_fibre_rot_n4", {}), ("e3_fibre_rot_n4", {"collapse": True}),
                 ("e2_hooke_sin_n6", {}), ("e3_fibre_rot_n8_c4", {}), ("e3_fibre_rot_n4", {"variant": 1})]:  # fmt: skip
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-6, **kw)
    s.set_grid(2)
    A, it, res = s.cell_tensors(K.points(case, 3), return_stats=True)
    print(name, kw, it.tolist(), float(np.abs(A).max()), flush=True)
    s.close()
print("done")
