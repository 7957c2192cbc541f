"""Committed golden vectors (tests/golden/oracle_vectors.json, made by tests/golden/make_golden.py): the
oracle must keep reproducing them (CPU), and the CUDA kernels must match them through the C ABI (GPU)."""
import json
import os

import numpy as np
import pytest

import cases as K

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))


@pytest.mark.parametrize("name", [n for n in GOLD if not K.BY_NAME[n].heavy])
def test_oracle_reproduces_golden_vectors(name):
    case = K.BY_NAME[name]
    prog = K.program(case)
    mic = K.oracle_cell(case, prog)
    for x, A in zip(GOLD[name]["x"], GOLD[name]["A_hom"]):
        got = K.oracle_tensor(case, mic, np.array(x))
        assert np.abs(got - np.array(A)).max() <= 1e-12 * np.abs(np.array(A)).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GOLD))
@pytest.mark.parametrize("collapse", [False, True])
def test_cuda_matches_golden_vectors(name, collapse):
    from hommx_b200 import native

    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-10, threads=case.threads, collapse=collapse)
    Ah = s.cell_tensors(np.array(GOLD[name]["x"]))
    for k, A in enumerate(GOLD[name]["A_hom"]):
        assert np.abs(Ah[k] - np.array(A)).max() <= 1e-10 * np.abs(np.array(A)).max()
    s.close()


DIRECT = [(n, co) for n in GOLD for co in (False, True) if K.BY_NAME[n].kind == 1 and K.BY_NAME[n].threads is None]


@pytest.mark.gpu
@pytest.mark.parametrize("name,collapse", DIRECT)
def test_direct_kernel_matches_golden_vectors(name, collapse):
    """K5 (dense Cholesky per cell) against the committed vectors, wherever the cell fits its 192 unknowns."""
    from hommx_b200 import native

    case = K.BY_NAME[name]
    prog = K.program(case)
    if not native.dense_fits(prog, case.n, native.collapse_mask(prog, collapse)):
        pytest.skip("more than 192 unknowns: PCG only")
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, variant=native.DENSE, collapse=collapse)
    Ah = s.cell_tensors(np.array(GOLD[name]["x"]))
    for k, A in enumerate(GOLD[name]["A_hom"]):
        assert np.abs(Ah[k] - np.array(A)).max() <= 1e-10 * np.abs(np.array(A)).max()
    s.close()
