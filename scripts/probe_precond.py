"""Block-Jacobi vs additive two-level PCG (csrc/hmx_cell_coarse.cuh) on the device: iterations, time, A_hom difference.
usage: python scripts/probe_precond.py [case ...]   (cases of tests/cases.py; default: the BASELINE config-4 cell)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import cases as K
from hommx_b200 import native

names = sys.argv[1:] or ["e3_fibre_rot_n8_c4"]
NPTS = int(os.environ.get("NPTS", "4736"))
for nm in names:
    case = K.BY_NAME[nm]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    x = K.points(case, NPTS)
    out = {}
    for mode in ("jacobi", "twolevel"):
        os.environ["HMX_PRECOND"] = mode
        if "build" in os.environ.get("PROBE", ""):
            native.compile_kernel(prog, case.n, case.threads)
            continue
        s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, atol=1e-10, threads=case.threads)
        s.cell_tensors(x[:256])
        s.rhs_iterations(reset=True)
        t = time.perf_counter()
        A, it, res = s.cell_tensors(x, return_stats=True)
        dt = time.perf_counter() - t
        its = s.rhs_iterations(reset=True) / (len(x) * s.m)
        out[mode] = A
        print(f"{nm} {mode:9s}: {len(x) / dt:10.0f} cell solves/s  mean it/rhs {its:7.1f}  max it {it.max()}  max res {res.max():.2e}  smem {s.info['smem_bytes']}")
        s.close()
    if out:
        d = np.abs(out["jacobi"] - out["twolevel"]).max() / np.abs(out["jacobi"]).max()
        print(f"{nm}: max rel. difference of A_hom between the two preconditioners {d:.2e}")
