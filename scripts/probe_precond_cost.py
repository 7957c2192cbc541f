"""Set-up vs per-iteration cost of the two preconditioners of the C4 cell kernel: time per cell at capped iteration counts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import cases as K
from hommx_b200 import native

nm = sys.argv[1] if len(sys.argv) > 1 else "e3_fibre_rot_n8_c4"
case = K.BY_NAME[nm]
prog = K.program(case)
qp, qw = K.tables(case, prog)
npts = 148 * 16
x = K.points(case, npts)
xd = torch.tensor(x, device="cuda")
A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device="cuda", dtype=torch.float64)
for mode in ("jacobi", "twolevel"):
    os.environ["HMX_PRECOND"] = mode
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, atol=1e-10, threads=case.threads)
    prev = None
    for cap in (1, 41, 81, 10000):
        s.set_tolerances(1e-8, 1e-10, cap)
        best = 1e9
        for _ in range(3):
            s.rhs_iterations(reset=True)
            torch.cuda.synchronize()
            t = time.perf_counter()
            s.cell_tensors_dev(npts, xd, A)
            s.sync()
            best = min(best, time.perf_counter() - t)
        its = s.rhs_iterations(reset=True) / (npts * s.m)
        us_cell = best / (npts / 148) * 1e6  # one CTA per SM
        extra = "" if prev is None else f"  -> {(us_cell - prev[0]) / max(its - prev[1], 1e-9):6.2f} us per added iteration"
        print(f"{nm} {mode:9s} cap {cap:5d}: {us_cell:9.1f} us per cell (per SM), {its:6.1f} it/rhs{extra}")
        prev = (us_cell, its)
    s.close()
