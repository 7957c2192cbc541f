"""Set-up vs per-iteration cost of the cluster kernel (variant 4) and of the matrix-free kernel: time per cell at capped iterations."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import cases as K
from hommx_b200 import native

nm = sys.argv[1] if len(sys.argv) > 1 else "e3_fibre_rot_n8_c4"
which = sys.argv[2:] or ["cluster", "matrix-free"]
case = K.BY_NAME[nm]
prog = K.program(case)
qp, qw = K.tables(case, prog)
npts = 148 * 8
xd = torch.tensor(K.points(case, npts), device="cuda")
A = torch.empty((npts, prog.n_rhs, prog.n_rhs), device="cuda", dtype=torch.float64)
for label in which:
    kw = dict(variant=native.CLUSTER) if label == "cluster" else dict(threads=case.threads, variant=native.MATRIX_FREE)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, atol=1e-10, **kw)
    units = s.info["resident_clusters"] or s.info["sms"] * s.info["ctas_per_sm"]
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
        us_cell = best / (npts / units) * 1e6
        extra = "" if prev is None else f"  -> {(us_cell - prev[0]) / max(its - prev[1], 1e-9):6.2f} us per added iteration"
        print(f"{nm} {label:11s} cap {cap:5d}: {us_cell:9.1f} us per cell (per CTA/cluster), {its:6.1f} it/rhs{extra}", flush=True)
        prev = (us_cell, its)
    s.close()
