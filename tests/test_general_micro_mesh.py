"""general periodic micro meshes (anything create_periodic_boundary_conditions, cell_problem.py:16-300,
accepts) through the element-list kernel csrc/hmx_cell_generic.cuh.  The oracle's MicroCell is mesh-agnostic (explicit
element loop, master map of cell_problem.py:38-300), so parity is kernel vs oracle on the SAME perturbed / re-split /
locally refined mesh -- CPU: emulator; GPU: through the C ABI and the drop-in classes."""
import os
import sys

import numpy as np
import pytest

import cases as K
import coefficients as Cf
import general_meshes as G
from hommx_b200 import codegen, micro, native, quadrature
from hommx_b200 import ufl as pufl
from oracle import hmm_oracle as ho
from oracle import npufl

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cpu_emu"))

# (name, mesh factory, dim, kind, coefficient, Dtheta)
MESHES = [
    ("p2_perturbed", lambda: G.perturbed(2, 8, 1), 2, 0, "smooth_sin", None),
    ("p2_flipped_strat", lambda: G.flipped_2d(8), 2, 0, "laminate", "dtheta_wavy"),
    ("p2_refined_fulltensor", lambda: G.refined_corner_2d(6), 2, 0, "full_tensor_2d", "dtheta_test_stratified"),
    ("p3_perturbed", lambda: G.perturbed(3, 4, 2), 3, 0, "smooth_sin", None),
    ("e2_perturbed_strat", lambda: G.perturbed(2, 6, 3), 2, 1, "hooke_sin_2d", "dtheta_test_stratified"),
    ("e3_perturbed_fibre", lambda: G.perturbed(3, 4, 4), 3, 1, "hooke_fibre_3d", "dtheta_rotation_3d"),
]
IDS = [m[0] for m in MESHES]


def _setup(spec):
    _, mk, dim, kind, coeff, dth = spec
    msh = mk()
    prog = codegen.build_program(getattr(Cf, coeff)(pufl), dim, kind, getattr(Cf, dth)(pufl) if dth else None)
    tables = micro.ElementListTables(msh, *quadrature.default_rule(dim, prog.degree))
    from oracle import ufldegree

    mic = ho.MicroCell(msh, "poisson" if kind == 0 else "elasticity", ufldegree.form_degree(getattr(Cf, coeff)(ufldegree), dim))

    def oracle(x):
        M = np.asarray(getattr(Cf, dth)(npufl)(np.asarray(x, float)))[..., 0] if dth else None
        return ho.cell_tensor(mic, getattr(Cf, coeff)(npufl), x, M)

    rng = np.random.default_rng(7)
    x = rng.uniform(0.05, 0.95, (3, 3))
    if dim == 2:
        x[:, 2] = 0.0
    return msh, prog, tables, oracle, x


def test_structured_detector_rejects_these_meshes_and_tables_are_consistent():
    for spec in MESHES:
        msh = spec[1]()
        with pytest.raises(ValueError):
            micro.detect_structure(msh)
        t = micro.ElementListTables(msh, *quadrature.default_rule(spec[2], 2))
        assert abs(t.elem_vol.sum() - 1.0) <= 1e-13
        assert t.blk_ptr[-1] == t.n_elem * (spec[2] + 1) ** 2 and t.node_ptr[-1] == t.n_elem * (spec[2] + 1)
        assert (t.col[t.diag] == np.arange(t.n_nodes)).all()
        # the P1 gradients of an element sum to zero, the block pattern is symmetric
        assert np.abs(t.elem_grad.sum(axis=1)).max() <= 1e-12
        rows = np.repeat(np.arange(t.n_nodes), np.diff(t.row_ptr))
        assert {(int(r), int(c)) for r, c in zip(rows, t.col)} == {(int(c), int(r)) for r, c in zip(rows, t.col)}
    bad = G.perturbed(2, 4, 0)
    bad.geometry.x[1, 0] += 0.01  # a node on the y = 0 face moves along the face: its partner on y = 1 does not
    with pytest.raises(ValueError, match="do not match periodically"):
        micro.ElementListTables(bad, *quadrature.default_rule(2, 2))


@pytest.mark.parametrize("spec", MESHES, ids=IDS)
def test_emulated_element_list_kernel_matches_oracle(spec):
    import emu

    msh, prog, tables, oracle, x = _setup(spec)
    s = emu.EmuSolver(prog, 0, None, None, rtol=1e-11, micro_tables=tables, grid=2)
    A, it, res = s.cell_tensors(x[:2], return_stats=True)
    for k in range(2):
        ref = oracle(x[k])
        assert np.abs(A[k] - ref).max() <= 1e-10 * np.abs(ref).max()
    assert (it > 0).all()


def test_emulated_element_list_kernel_equals_structured_kernel_on_a_structured_mesh():
    """On the structured mesh both paths discretise the same problem: same tensor, same local matrix, same correctors."""
    import emu
    from hommx_b200 import mesh

    case = K.BY_NAME["e2_hooke_sin_strat_n7"]
    prog = K.program(case)
    qp, qw = K.tables(case, prog)
    msh = mesh.create_unit_square(case.n, case.n)
    tables = micro.ElementListTables(msh, *quadrature.default_rule(2, prog.degree))
    gen = emu.EmuSolver(prog, 0, None, None, rtol=1e-12, micro_tables=tables, grid=2)
    ref = emu.EmuSolver(prog, case.n, qp, qw, rtol=1e-12)
    cells, xyz = K.random_simplices(2, 2)
    Sa, Aa = gen.local_matrices(cells, xyz)
    Sb, Ab = ref.local_matrices(cells, xyz)
    assert np.abs(Aa - Ab).max() <= 1e-11 * np.abs(Ab).max() and np.abs(Sa - Sb).max() <= 1e-11 * np.abs(Sb).max()
    x = K.points(case, 1)
    ca = gen.correctors(x)[0]  # (n_rhs, bs, n_periodic_nodes)
    cb = ref.correctors(x)[0]  # (n_rhs, bs, ny, nx)
    ij = np.rint(msh.x[:, :2] * case.n).astype(np.int64) % case.n
    full_a = ca[:, :, tables.node2per]
    full_b = cb[:, :, ij[:, 1], ij[:, 0]]
    full_a = full_a - full_a.mean(axis=-1, keepdims=True)
    full_b = full_b - full_b.mean(axis=-1, keepdims=True)
    assert np.abs(full_a - full_b).max() <= 1e-9 * np.abs(full_b).max()


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("spec", MESHES, ids=IDS)
def test_element_list_kernel_matches_oracle(spec):
    msh, prog, tables, oracle, x = _setup(spec)
    s = native.CellSolver(prog, 0, None, None, rtol=1e-11, micro_tables=tables)
    assert s.variant == native.ELEMENT_LIST
    A, it, res = s.cell_tensors(x, return_stats=True)
    for k in range(len(x)):
        ref = oracle(x[k])
        assert np.abs(A[k] - ref).max() <= 1e-10 * np.abs(ref).max()
    B = s.cell_tensors(x)
    s.set_grid(1)
    C = s.cell_tensors(x)
    assert np.array_equal(A, B) and np.array_equal(A, C)  # fixed summation orders: bitwise reproducible
    s.close()


@pytest.mark.gpu
def test_drop_in_classes_accept_a_general_micro_mesh():
    """PoissonHMM / PoissonPeriodicHMM / LinearElasticityHMM with a perturbed micro mesh: macro matrix vs the literal
    oracle assembly, A_hom and correctors vs the oracle."""
    from hommx_b200 import LinearElasticityHMM, PoissonHMM, PoissonPeriodicHMM, mesh

    mic = G.perturbed(2, 8, 5)
    m = mesh.create_unit_square(3, 3)
    hmm = PoissonHMM(m, Cf.smooth_sin(pufl), lambda x: 1.0, mic, 0.1)
    hmm._assemble_stiffness()
    assert hmm.cell_solver_used.startswith("pcg (element list)")
    omic = ho.MicroCell(mic, "poisson", hmm._program.degree)
    S5 = hmm._compute_local_stiffness(5)
    verts = m.x[m.cells[5]]
    ref = ho.local_stiffness_literal(omic, Cf.smooth_sin(npufl), verts, 0.1)
    assert np.abs(S5 - ref).max() <= 1e-10 * np.abs(ref).max()
    u = hmm.solve()
    assert np.isfinite(u.x.array).all() and np.abs(u.x.array).max() > 0

    per = PoissonPeriodicHMM(m, lambda y: 2.0 + pufl.sin(2 * pufl.pi * y[0]), lambda x: 1.0, mic, 0.1)
    Ah = per.compute_effective_tensor()
    ref = ho.cell_tensor(omic, Cf.periodic_only(npufl), np.zeros(3))
    assert np.abs(Ah - ref).max() <= 1e-10 * np.abs(ref).max()
    assert len(per.correctors) == 2 and per.correctors[0].x.array.shape[0] == mic.num_nodes

    mic3 = G.perturbed(3, 4, 6)
    m3 = mesh.create_box((0.0, 0.0, 0.0), (1.0, 0.5, 0.25), (2, 1, 1))
    el = LinearElasticityHMM(m3, Cf.hooke_smooth_3d(pufl), lambda x: pufl.as_vector([0.0, 0.0, -1.0]), mic3, 0.1)
    el._assemble_stiffness()
    omic3 = ho.MicroCell(mic3, "elasticity", el._program.degree)
    S = el._compute_local_stiffness(2)
    ref = ho.local_stiffness_literal(omic3, Cf.hooke_smooth_3d(npufl), m3.x[m3.cells[2]], 0.1)
    assert np.abs(S - ref).max() <= 1e-10 * np.abs(ref).max()
