"""Repro of the round-1 'look-ahead NaN' of the direct kernel K5 (csrc/hmx_cell_dense.cuh).

-DHMX_DENSE_LOOKAHEAD=2 is the look-ahead with column k + 1 updated inside the owners' branch: same FMAs per matrix
entry as variants 0 / 1, finite and oracle-exact in the CPU emulation and at `-Xptxas -O0`, NaN pivots on the device at
ptxas -O1 and above (192-unknown cells; first bad pivot at a data-dependent, run-to-run reproducible step).  Variant 1
(uniform control flow, only the stores predicated) is the product default.  Run on a B200:
    python scripts/repro_dense_lookahead.py
Measured (round 2): variant 0 finite, 1 finite, 2 NaN from step 43 / 34, 2 at -Xptxas -O0 finite."""
import os
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import cases as K
from hommx_b200 import native

case = K.BY_NAME["e3_fibre_rot_n4"]  # full 4^3 cell: 192 unknowns
prog = K.program(case)
qp, qw = K.tables(case, prog)
x = K.points(case, 8, seed=3)
for flags in ("0", "1", "2", "2 -Xptxas -O0"):
    os.environ["HMX_EXTRA_NVCC"] = f"-DHMX_DENSE_LOOKAHEAD={flags} -DHMX_DENSE_DEBUG"  # DEBUG: iters = first step with a non-finite pivot
    s = native.CellSolver(prog, case.n, qp, qw, variant=native.DENSE)
    A, first_bad, _ = s.cell_tensors(x, return_stats=True)
    print(f"HMX_DENSE_LOOKAHEAD={flags:14s} first non-finite pivot per point {first_bad}  tensors finite: {bool(np.isfinite(A).all())}")
    s.close()
