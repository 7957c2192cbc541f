"""BASELINE.json's configurations at their FULL sizes on the GPU, checked through size-independent
properties (the oracle needs ~30 ms per macro point and cannot follow there):

* closed forms that hold at every macro point,
* symmetry and positive definiteness of A_hom, rigid-body null space of the local macro matrices,
* two independently compiled kernels (full micro cell vs. exact axis collapse) agreeing to 1e-10,
* macro points with equal coefficient giving bit-identical tensors.

The macro points are the cell barycentres of the configuration's macro mesh (hmm.py:350)."""
import numpy as np
import pytest

import cases as K
import coefficients as Cf
from hommx_b200 import mesh, native
from oracle import npufl

pytestmark = pytest.mark.gpu


def _barycentres(msh):
    return np.ascontiguousarray(msh.x[msh.cells].mean(axis=1))


def _solver(name, collapse=False, rtol=1e-9, variant=None):
    case = K.BY_NAME[name]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    return native.CellSolver(prog, case.n, qp, qw, rtol=rtol, collapse=collapse, variant=variant)


def _spd(A, tol=1e-10):
    scale = np.abs(A).max(axis=(1, 2), keepdims=True)
    assert np.abs(A - A.transpose(0, 2, 1)).max() <= tol * scale.max()
    assert np.linalg.eigvalsh(0.5 * (A + A.transpose(0, 2, 1))).min() > 0.0


def test_c1_poisson_2d_full_size():
    """configs[0]: 32x32 macro mesh (2,048 points), 16^2 micro cell, A = 1.1 + x0 + sin(2 pi y0)."""
    x = _barycentres(mesh.create_unit_square(32, 32))
    s = _solver("p2_smooth_n16_c1")
    A, it, res = s.cell_tensors(x, return_stats=True)
    assert len(A) == 2048 and res.max() <= 1e-9
    _spd(A)
    # a depends on y0 only: the x1 direction sees the arithmetic mean 1.1 + x0 (sin sums to zero over the
    # uniform periodic grid), the x0 direction lies between the harmonic and the arithmetic mean, no coupling
    a = 1.1 + x[:, 0]
    assert np.abs(A[:, 1, 1] - a).max() <= 1e-10 * a.max()
    assert np.abs(A[:, 0, 1]).max() <= 1e-10 * a.max()
    harm = np.sqrt(a * a - 1.0)  # continuum harmonic mean of a + sin
    assert np.all(A[:, 0, 0] < a) and np.all(A[:, 0, 0] > harm)
    # points of one macro column share x0 up to the last bit only: compare through the closed form instead,
    # but equal points must give identical bits
    A2 = s.cell_tensors(np.ascontiguousarray(x[::-1]))[::-1]
    assert np.array_equal(A, A2)
    s.close()


def test_c2_wavy_laminate_full_size_closed_form():
    """configs[1]: 256x256 macro mesh (131,072 points), 32^2 micro cell, laminate 5 / 0.05 under
    theta(x) = (x1 - sin 2 pi x0, x0).  The P1 discretisation reproduces the laminate formula exactly
    (SURVEY.md 8c.6): every one of the 131,072 tensors is checked against it."""
    x = _barycentres(mesh.create_unit_square(256, 256))
    Dt = Cf.dtheta_wavy(npufl)
    full, coll = _solver("p2_laminate_wavy_n32_c2"), _solver("p2_laminate_wavy_n32_c2", collapse=True)
    A = full.cell_tensors(x)
    Ac = coll.cell_tensors(x)
    assert len(A) == 131072
    M = np.stack([np.asarray(Dt(p))[..., 0] for p in x[:512]])  # the map depends on x0 only: 512 distinct columns suffice
    key = {round(float(p[0]), 12): m for p, m in zip(x[:512], M)}
    nv = np.stack([key[round(float(p[0]), 12)][:, 0] for p in x])
    nv /= np.linalg.norm(nv, axis=1, keepdims=True)
    P = nv[:, :, None] * nv[:, None, :]
    expect = 2.525 * (np.eye(2)[None] - P) + (1.0 / (0.5 / 5 + 0.5 / 0.05)) * P
    assert np.abs(A - expect).max() <= 1e-10 * np.abs(expect).max()
    assert np.abs(Ac - expect).max() <= 1e-10 * np.abs(expect).max()
    full.close()
    coll.close()


def test_c3_poisson_3d_full_size():
    """configs[2]: 32^3 macro mesh (196,608 points), 8^3 micro cell."""
    x = _barycentres(mesh.create_unit_cube(32, 32, 32))
    full, coll = _solver("p3_smooth_n8_c3"), _solver("p3_smooth_n8_c3", collapse=True)
    A = full.cell_tensors(x)
    Ac = coll.cell_tensors(x)
    assert len(A) == 196608
    _spd(A)
    a = 1.1 + x[:, 0]
    for k in (1, 2):
        assert np.abs(A[:, k, k] - a).max() <= 1e-10 * a.max()
    off = A.copy()
    off[:, [0, 1, 2], [0, 1, 2]] = 0.0
    assert np.abs(off).max() <= 1e-10 * a.max()
    assert np.abs(A - Ac).max() <= 1e-10 * np.abs(A).max()  # two different kernels
    full.close()
    coll.close()


def test_c4_rotated_fibres_throughput_size():
    """configs[3] at the bench size: 40x16x4 hexahedra -> 15,360 macro cells, 8^3 micro cell, 6 right-hand
    sides.  Full cell (block sweep) against the axis-collapsed PCG kernel and the direct (dense Cholesky) kernel, tensor symmetries, rigid-body null
    space of the 12x12 local matrices, equal rotation angle -> identical bits."""
    msh = mesh.create_box((0.0, 0.0, 0.0), (1.0, 0.4, 0.1), (40, 16, 4))
    x = _barycentres(msh)
    assert len(x) == 15360
    full, coll = _solver("e3_fibre_rot_n8_c4"), _solver("e3_fibre_rot_n8_c4", collapse=True)
    A, it, res = full.cell_tensors(x, return_stats=True)
    Ac = coll.cell_tensors(x)
    assert res.max() <= 1e-9 and it.max() < 1000
    _spd(A)
    assert np.abs(A - Ac).max() <= 1e-10 * np.abs(A).max()
    direct = _solver("e3_fibre_rot_n8_c4", collapse=True, variant=native.DENSE)  # K5: a third, direct kernel
    Ad = direct.cell_tensors(x)
    assert np.abs(Ad - A).max() <= 1e-10 * np.abs(A).max() and np.abs(Ad - Ac).max() <= 1e-10 * np.abs(A).max()
    direct.close()
    # the coefficient sees the macro point through the angle gamma(x1) only
    order = np.lexsort((x[:, 2], x[:, 0], x[:, 1]))
    xs, As = x[order], A[order]
    same = np.flatnonzero(xs[1:, 1] == xs[:-1, 1])
    assert len(same) > 10000 and np.array_equal(As[same + 1], As[same])
    # local macro matrices: translations of the macro element carry no energy
    cells = np.ascontiguousarray(msh.cells[::8], dtype=np.int32)  # 1,920 of the macro cells
    gp = np.arange(len(cells) * 144 + 1, dtype=np.int64)  # identity gather: slot j <- S_flat[j]
    gs = np.arange(len(cells) * 144, dtype=np.int32)
    _, S = full.assemble_macro(cells, msh.x, gp, gs, want_local=True)
    S = np.asarray(S).reshape(-1, 12, 12)
    t = np.zeros((12, 3))
    for k in range(3):
        t[k::3, k] = 1.0  # unrolled dof = node * 3 + component (hmm.py:31-40)
    assert np.abs(S @ t).max() <= 1e-9 * np.abs(S).max()
    assert np.abs(S - S.transpose(0, 2, 1)).max() <= 1e-10 * np.abs(S).max()
    full.close()
    coll.close()
