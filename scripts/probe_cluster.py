"""The cluster-resident assembled kernel (variant 4) against the matrix-free kernel: agreement and time per cell."""
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

names = sys.argv[1:] or ["e3_fibre_rot_n4", "e3_fibre_rot_n8_c4"]
import dataclasses

for nm in names:
    nm, _, nover = nm.partition(":")  # "case[:n]": the case's coefficient on an n^3 cell
    case = K.BY_NAME[nm]
    if nover:
        case = dataclasses.replace(case, n=int(nover), threads=None)
        nm = f"{nm}[n={case.n}]"
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    npts = 148 * (8 if case.n <= 8 else 2)
    x = K.points(case, npts)
    xd = torch.tensor(x, device="cuda")
    out = {}
    for label, kw in (("matrix-free", dict(threads=case.threads, variant=native.MATRIX_FREE)), ("cluster", dict(variant=native.CLUSTER))):
        s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-8, atol=1e-10, **kw)
        A = torch.zeros((npts, prog.n_rhs, prog.n_rhs), device="cuda", dtype=torch.float64)
        best = 1e9
        for _ in range(3):
            s.rhs_iterations(reset=True)
            torch.cuda.synchronize()
            t = time.perf_counter()
            s.cell_tensors_dev(npts, xd, A)
            s.sync()
            best = min(best, time.perf_counter() - t)
        its = s.rhs_iterations(reset=True) / (npts * s.m)
        out[label] = A.cpu().numpy()
        print(f"{nm} {label:12s}: {npts / best:9.0f} cells/s, {its:6.1f} it/rhs, info {s.info}", flush=True)
        s.close()
    d = np.abs(out["cluster"] - out["matrix-free"]).max() / np.abs(out["matrix-free"]).max()
    print(f"{nm}: max |A_cluster - A_mf| / max |A| = {d:.2e}", flush=True)
