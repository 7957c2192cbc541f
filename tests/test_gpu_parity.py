"""Parity tests proper: the sm_100a kernels, called through the C ABI (libhmx.so via ctypes),
against the CPU oracle on the same seeded inputs.  Tolerance on A_hom and S_loc: 1e-10 relative
(BASELINE.json north star)."""
import numpy as np
import pytest

import cases as K
from hommx_b200 import native
from oracle import hmm_oracle as ho

pytestmark = pytest.mark.gpu

ALL = [c for c in K.CASES if K.kernel_available(c)]


def _solver(case, prog):
    qp, qw = K.tables(case, prog)
    return native.CellSolver(prog, case.n, qp, qw, rtol=case.rtol, threads=case.threads)


@pytest.mark.parametrize("case", ALL, ids=[c.name for c in ALL])
def test_cell_tensor_matches_oracle(case):
    prog = K.program(case)
    s = _solver(case, prog)
    x = K.points(case, 2 if case.heavy else 4)
    Ah, it, res = s.cell_tensors(x, return_stats=True)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max(), (case.name, it, res)
    s.close()


@pytest.mark.parametrize("name", ["p2_fulltensor_strat_n9", "p3_fulltensor_shear_n5", "e2_hooke_sin_strat_n7", "e3_hooke_smooth_n4"])
def test_local_matrix_matches_oracle(name):
    case = K.BY_NAME[name]
    if not K.kernel_available(case):
        pytest.skip("kernel kind not built yet")
    prog = K.program(case)
    s = _solver(case, prog)
    cells, xyz = K.random_simplices(case.dim, 5)
    nb2 = s.nb * s.nb
    gp = np.arange(len(cells) * nb2 + 1, dtype=np.int64)  # identity gather: slot j <- S_flat[j]
    gs = np.arange(len(cells) * nb2, dtype=np.int32)
    vals, S = s.assemble_macro(cells, xyz, gp, gs, want_local=True)
    mic = K.oracle_cell(case, prog)
    for k in range(len(cells)):
        verts = xyz[cells[k]]
        So = ho.local_stiffness_from_tensor(K.oracle_tensor(case, mic, verts.mean(axis=0)), verts, mic.kind)
        assert np.abs(S[k] - So).max() <= case.tol * np.abs(So).max()
    assert np.array_equal(vals, S.reshape(-1))
    s.close()


DENSE = [(c, co) for c in ALL if c.kind == 1 and c.threads is None for co in (False, True)
         if native.dense_fits(K.program(c), c.n, native.collapse_mask(K.program(c), co))
         and (not co or native.collapse_mask(K.program(c), True))]


@pytest.mark.parametrize("case,collapse", DENSE, ids=[c.name + ("_collapsed" if co else "") for c, co in DENSE])
def test_dense_cholesky_variant_matches_oracle(case, collapse):
    """K5: small elasticity cells factorised directly (csrc/hmx_cell_dense.cuh), incl. the 192-unknown collapsed
    cell of BASELINE config 4: A_hom against the oracle, local matrices and correctors against the PCG kernel."""
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, variant=native.DENSE, collapse=collapse)
    assert s.variant == native.DENSE
    x = K.points(case, 4)
    Ah, it, res = s.cell_tensors(x, return_stats=True)
    assert np.all(it == 0) and np.all(res == 0.0)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()
    ref = native.CellSolver(prog, case.n, qp, qw, rtol=1e-12, collapse=collapse, variant=native.MATRIX_FREE)
    cells, xyz = K.random_simplices(case.dim, 5)
    nb2 = s.nb * s.nb
    gp = np.arange(len(cells) * nb2 + 1, dtype=np.int64)
    gs = np.arange(len(cells) * nb2, dtype=np.int32)
    S, Sr = s.assemble_macro(cells, xyz, gp, gs, want_local=True)[1], ref.assemble_macro(cells, xyz, gp, gs, want_local=True)[1]
    assert np.abs(S - Sr).max() <= 1e-10 * np.abs(Sr).max()
    a, b = s.cell_correctors(x[:1])[0], ref.cell_correctors(x[:1])[0]
    ax = tuple(range(2, a.ndim))
    a, b = a - a.mean(axis=ax, keepdims=True), b - b.mean(axis=ax, keepdims=True)
    assert np.abs(a - b).max() <= 1e-8 * np.abs(b).max() + 1e-14
    s.close()
    ref.close()


COLLAPSIBLE = [c for c in ALL if K.program(c).ydep != (1 << c.dim) - 1]


@pytest.mark.parametrize("case", COLLAPSIBLE, ids=[c.name for c in COLLAPSIBLE])
def test_axis_collapse_is_exact(case):
    """Axes the coefficient does not depend on collapsed to one layer of cubes: same A_hom as the full
    n^d oracle (includes BASELINE configs 1-4: laminates and fibres)."""
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=case.rtol, collapse=True)
    x = K.points(case, 3)
    Ah = s.cell_tensors(x)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()
    s.close()


def test_many_points_grid_stride():
    """More points than resident CTAs: every point is computed exactly once (persistent grid)."""
    case = K.BY_NAME["p2_smooth_n16_c1"]
    prog = K.program(case)
    s = _solver(case, prog)
    rng = np.random.default_rng(0)
    x = np.zeros((5000, 3))
    x[:, :2] = rng.uniform(0, 1, (5000, 2))
    Ah = s.cell_tensors(x)
    # A = 1.1 + x0 + sin(2 pi y0): A_hom[1,1] is the arithmetic mean 1.1 + x0 exactly
    assert np.allclose(Ah[:, 1, 1], 1.1 + x[:, 0], rtol=1e-12)
    # and it is a smooth function of x0 only: recompute a subset on a tiny grid
    s.set_grid(3)
    sub = s.cell_tensors(x[:50])
    assert np.array_equal(sub, Ah[:50])
    s.close()


def test_errors_are_reported():
    case = K.BY_NAME["p2_xonly_n7"]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    with pytest.raises(native.HmxError):  # kernel built for n=7, descriptor says n=9
        import ctypes as C

        cubin = native.compile_kernel(prog, 7)
        img = open(cubin, "rb").read()
        buf = C.create_string_buffer(img, len(img))
        d = native.hmx_desc(2, 0, 9, len(qw), qp.ctypes.data_as(C.POINTER(C.c_double)), qw.ctypes.data_as(C.POINTER(C.c_double)),
                            C.cast(buf, C.c_void_p), len(img), 1e-8, 1e-10, 100, 0)
        h = C.c_void_p()
        lib = native.load_library()
        rc = lib.hmx_create(C.byref(h), C.byref(d))
        assert rc == -3
        raise native.HmxError(lib.hmx_last_error(None).decode())


def test_empty_and_degenerate_inputs():
    """Edge cases at the C ABI: zero points / zero cells are no-ops, max_it is honoured and reported,
    and a constant coefficient (no atoms, zero right-hand sides) needs no iteration."""
    case = K.BY_NAME["p2_inclusion_n16"]
    prog = K.program(case)
    s = _solver(case, prog)
    assert s.cell_tensors(np.zeros((0, 3))).shape == (0, 2, 2)
    vals = s.assemble_macro(np.zeros((0, 3), dtype=np.int32), np.zeros((1, 3)), np.zeros(5, dtype=np.int64), np.zeros(0, dtype=np.int32))
    assert vals.shape == (4,) and np.all(vals == 0.0)
    x = K.points(case, 3)
    s.set_tolerances(1e-12, 1e-14, 5)
    A5, it, res = s.cell_tensors(x, return_stats=True)
    assert np.all(it == 5) and np.all(res > 1e-6)  # stopped early, residual reported, no exception
    s.set_tolerances(1e-10, 1e-12, 10000)
    A, it, res = s.cell_tensors(x, return_stats=True)
    assert np.all(it < 10000) and np.all(res <= 1e-10) and np.abs(A - A5).max() > 0
    s.close()
    case = K.BY_NAME["p2_xonly_n7"]
    s = _solver(case, K.program(case))
    A, it, res = s.cell_tensors(K.points(case, 4), return_stats=True)
    assert np.all(it == 0) and np.all(res == 0.0)
    s.close()
    with pytest.raises(ValueError):
        native.CellSolver(K.program(case), 7, np.zeros((2, 1, 3)), np.ones(1))  # wrong quadrature table shape


@pytest.mark.parametrize("name,kw", [("p2_inclusion_n16", {}), ("p3_fulltensor_n6", {}), ("e3_fibre_rot_n8_c4", {}),
                                     ("e3_fibre_rot_n4", {"collapse": True}), ("e2_hooke_sin_n6", {}),
                                     ("e3_hooke_smooth_shear_n6", {}),
                                     ("e3_fibre_rot_n10_l2", {}), ("e3_fibre_rot_n8_c4", {"collapse": True, "variant": 3}),
                                     ("e3_cubic_shear_n4", {"variant": 3})])  # fmt: skip
def test_results_are_bitwise_reproducible(name, kw):
    """No atomics on the data path, fixed reduction orders, colour-ordered scatter: repeated runs and
    different grid sizes give bit-identical A_hom (compute-sanitizer is closed on this pool; a data race in
    the colouring / named-barrier logic would show up here as run-to-run differences)."""
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-9, **kw)
    x = K.points(case, 300 if not case.heavy else 200, seed=11)
    ref = s.cell_tensors(x)
    for grid in (0, 7, 148):
        s.set_grid(grid)
        for _ in range(2):
            assert np.array_equal(s.cell_tensors(x), ref)
    s.close()


def test_nvrtc_built_kernel_matches_oracle(monkeypatch):
    """Hosts without nvcc build the cell kernel in-process with NVRTC (same translation unit)."""
    monkeypatch.setenv("HMX_COMPILER", "nvrtc")
    case = K.BY_NAME["e3_fibre_rot_n4"]
    prog = K.program(case)
    s = _solver(case, prog)
    x = K.points(case, 3)
    Ah = s.cell_tensors(x)
    mic = K.oracle_cell(case, prog)
    for k in range(len(x)):
        Ao = K.oracle_tensor(case, mic, x[k])
        assert np.abs(Ah[k] - Ao).max() <= case.tol * np.abs(Ao).max()
    s.close()


@pytest.mark.parametrize("precond", ["twolevel", "jacobi"])
def test_c4_full_cell_local_matrices_at_32_macro_cells(precond, monkeypatch):
    """The bench headline kernel (BASELINE config 4: full 8^3 fibre cell, 6 right-hand sides) against the oracle at
    32 random macro cells of the beam [0,1] x [0,0.4] x [0,0.1]: the 12 x 12 local matrices S_loc (not only A_hom)
    against the tensor form of the oracle at every cell and against the LITERAL restatement of hmm.py:334-369
    (n_b = 12 eps-scaled correctors, 144 integrals) at 6 of them; both preconditioners of the PCG kernel."""
    monkeypatch.setenv("HMX_PRECOND", precond)
    case = K.BY_NAME["e3_fibre_rot_n8_c4"]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    s = native.CellSolver(prog, case.n, qp, qw, rtol=1e-9, atol=1e-12)
    rng = np.random.default_rng(32)
    n = 32
    base = rng.uniform([0.05, 0.02, 0.005], [0.9, 0.36, 0.09], (n, 3))
    ref = np.zeros((4, 3))
    ref[1:] = np.eye(3) * np.array([0.05, 0.02, 0.008])
    xyz = (base[:, None, :] + ref[None] + rng.uniform(-0.002, 0.002, (n, 4, 3))).reshape(-1, 3)
    cells = np.arange(4 * n, dtype=np.int32).reshape(n, 4)
    nb2 = s.nb * s.nb
    gp = np.arange(n * nb2 + 1, dtype=np.int64)
    gs = np.arange(n * nb2, dtype=np.int32)
    _, S, it, res = s.assemble_macro(cells, xyz, gp, gs, want_local=True, return_stats=True)
    mic = K.oracle_cell(case)
    A = getattr(K.Cf, case.coeff)(K.npufl)
    Dn = getattr(K.Cf, case.dtheta)(K.npufl)
    Dt = lambda x: np.asarray(Dn(np.asarray(x, float)))[..., 0]  # noqa: E731
    for k in range(n):
        verts = xyz[cells[k]]
        So = ho.local_stiffness_from_tensor(K.oracle_tensor(case, mic, verts.mean(axis=0)), verts, mic.kind)
        assert np.abs(S[k] - So).max() <= 1e-10 * np.abs(So).max(), (k, it[k], res[k])
        if k % 6 == 0:
            Sl = ho.local_stiffness_literal(mic, A, verts, 0.01, Dt)
            assert np.abs(S[k] - Sl).max() <= 1e-10 * np.abs(Sl).max(), (k, "literal")
    if precond == "twolevel":
        assert it.max() < 200  # block Jacobi needs ~250 on this cell
    s.close()
