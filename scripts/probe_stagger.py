"""Per-iteration cost of the C4 cell kernel with the right-hand-side groups started HMX_STAGGER cycles apart."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import cases as K
from hommx_b200 import native

nm = "e3_fibre_rot_n8_c4"
case = K.BY_NAME[nm]
prog = K.program(case)
qp, qw = K.tables(case, prog)
npts = 148 * 16
xd = torch.tensor(K.points(case, npts), device="cuda")
A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device="cuda", dtype=torch.float64)
for stag in [int(a) for a in sys.argv[1:]] or [0, 2000, 4000, 5600, 8000]:
    os.environ["HMX_EXTRA_NVCC"] = f"-DHMX_STAGGER={stag}"
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, atol=1e-10, threads=case.threads)
    best = 1e9
    for _ in range(4):
        s.rhs_iterations(reset=True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        s.cell_tensors_dev(npts, xd, A)
        s.sync()
        best = min(best, time.perf_counter() - t)
    its = s.rhs_iterations(reset=True) / (npts * s.m)
    print(f"{nm} stagger {stag:6d}: {best / (npts / 148) * 1e6:9.1f} us per cell (per SM), {its:6.1f} it/rhs, "
          f"{npts / best:9.0f} cells/s", flush=True)
    s.close()
