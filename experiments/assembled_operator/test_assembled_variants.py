"""Parity tests of the two assembled-operator experiments (variants 1 and 2 of the elasticity cell kernel).
Not part of the default suites: run with HMX_EXTRA_NVCC set as experiments/assembled_operator/README.md says."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cpu_emu")):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("HMX_EXTRA_NVCC", f"-DHMX_EXPERIMENTAL_VARIANTS -I{os.path.dirname(os.path.abspath(__file__))}")

import cases as K  # noqa: E402
import emu  # noqa: E402
from hommx_b200 import native  # noqa: E402

ASSEMBLED, ASSEMBLED_TMA = 1, 2
ELAST = [c for c in K.CASES if c.kind == 1 and c.dim == 3 and not c.heavy and c.n**c.dim <= 1024]
ELAST_TMA = [c for c in ELAST if (c.n**c.dim) % 32 == 0]


def _threads(case, variant):
    N = case.n**case.dim
    return N if variant == ASSEMBLED_TMA else max(64, 32 * (-(-N // 32)))  # one thread per node


def _check(s, case, n_pts):
    prog = K.program(case)
    x = K.points(case, n_pts)
    Ah = s.cell_tensors(x)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()


@pytest.mark.parametrize("case", ELAST, ids=[c.name for c in ELAST])
def test_assembled_variant_emulated(case):
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    _check(emu.EmuSolver(prog, case.n, qp, qw, rtol=case.rtol, variant=ASSEMBLED, threads=_threads(case, ASSEMBLED)), case, 2)


@pytest.mark.parametrize("case", ELAST_TMA, ids=[c.name for c in ELAST_TMA])
def test_tma_staged_variant_emulated(case):
    """(the emulation models the mbarrier protocol -- phases, transaction bytes, multi-point reuse -- not the asynchrony)"""
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    _check(emu.EmuSolver(prog, case.n, qp, qw, rtol=case.rtol, variant=ASSEMBLED_TMA, threads=_threads(case, ASSEMBLED_TMA), grid=2), case, 5)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [ASSEMBLED, ASSEMBLED_TMA])
@pytest.mark.parametrize("case", ELAST_TMA, ids=[c.name for c in ELAST_TMA])
def test_assembled_variants_on_the_device(case, variant):
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=case.rtol, variant=variant, threads=_threads(case, variant))
    _check(s, case, 64)
    s.close()
