"""End-to-end parity of the drop-in classes on the GPU: assembled macro matrix and macro solution
against the oracle (the literal restatement of hmm.py:298-369 + :434-491) on identical meshes, and
the known answers the reference's own tests pin."""
import math

import numpy as np
import pytest

import coefficients as Cf
from hommx_b200 import (
    LinearElasticityHMM,
    LinearElasticityStratifiedHMM,
    PoissonHMM,
    PoissonPeriodicHMM,
    PoissonStratifiedHMM,
    fem,
    mesh,
)
from hommx_b200 import ufl as pufl
from oracle import hmm_oracle as ho
from oracle import meshes as omesh
from oracle import npufl

pytestmark = pytest.mark.gpu
TIGHT = {"ksp_rtol": 1e-10, "ksp_atol": 1e-12}


def _np_dtheta(name):
    Dt = getattr(Cf, name)(npufl)
    return lambda x: np.asarray(Dt(np.asarray(x, float)))[..., 0]


def test_poisson_hmm_matrix_and_solution_match_oracle():
    """BASELINE config 1 scaled down: PoissonHMM, A = 1.1 + x0 + sin(2 pi y0) (examples/hmm.py:15-16)."""
    nm, n = 6, 8
    s = PoissonHMM(mesh.create_unit_square(nm, nm), Cf.smooth_sin(pufl), lambda x: 1.0, mesh.create_unit_square(n, n), 2.0**-5,
                   petsc_options_cell_problem=TIGHT)  # fmt: skip
    u = s.solve()
    macro = omesh.create_unit_square(nm, nm)
    mic = ho.MicroCell(omesh.create_unit_square(n, n), "poisson", 3)
    Ao = ho.assemble_macro(macro, mic, Cf.smooth_sin(npufl), 2.0**-5, literal=True)
    assert np.linalg.norm((s._A - Ao).toarray()) <= 1e-10 * np.linalg.norm(Ao.toarray())
    bc = ho.boundary_nodes(macro)
    uo = ho.solve_dirichlet(Ao, ho.assemble_rhs(macro, lambda x: 1.0 + 0 * x[0], 1, degree=1), bc, np.zeros(len(bc)))
    assert np.abs(u.x.array - uo).max() <= 1e-8 * np.abs(uo).max()
    # the finer seam: one local matrix
    S = s._compute_local_stiffness(5)
    So = ho.local_stiffness_literal(mic, Cf.smooth_sin(npufl), macro.x[macro.cells[5]], 2.0**-5)
    assert np.abs(S - So).max() <= 1e-10 * np.abs(So).max()


def test_poisson_stratified_matches_oracle_and_closed_form():
    """BASELINE config 2 scaled down: wavy laminate (examples/diffusion/laminate.py:101-117)."""
    nm, n = 5, 8
    m = mesh.create_unit_square(nm, nm)
    s = PoissonStratifiedHMM(m, Cf.laminate(pufl), lambda x: 1.0, mesh.create_unit_square(n, n), 1e-5, Cf.dtheta_wavy(pufl),
                             petsc_options_cell_problem=TIGHT)  # fmt: skip
    V = s.function_space
    left = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))
    right = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 1.0))
    s.set_boundary_conditions([fem.dirichletbc(1.0, left, V), fem.dirichletbc(0.0, right, V)])  # laminate.py:62-86
    u = s.solve()
    macro = omesh.create_unit_square(nm, nm)
    mic = ho.MicroCell(omesh.create_unit_square(n, n), "poisson", 0)
    Ao = ho.assemble_macro(macro, mic, Cf.laminate(npufl), 1e-5, _np_dtheta("dtheta_wavy"))
    assert np.linalg.norm((s._A - Ao).toarray()) <= 1e-10 * np.linalg.norm(Ao.toarray())
    b = ho.assemble_rhs(macro, lambda x: 1.0 + 0 * x[0], 1, degree=1)
    dofs = np.concatenate([left, right])
    uo = ho.solve_dirichlet(Ao, b, dofs, np.concatenate([np.ones(len(left)), np.zeros(len(right))]))
    assert np.abs(u.x.array - uo).max() <= 1e-8 * np.abs(uo).max()
    # closed form: <a>(I - n n^T) + <1/a>^-1 n n^T with n = M e0 / |M e0|
    x = np.array([[0.3, 0.6, 0.0]])
    Ah = s.cell_tensors(x)[0]
    M = _np_dtheta("dtheta_wavy")(x[0])
    nv = M[:, 0] / np.linalg.norm(M[:, 0])
    expect = 2.525 * (np.eye(2) - np.outer(nv, nv)) + (1.0 / (0.5 / 5 + 0.5 / 0.05)) * np.outer(nv, nv)
    assert np.allclose(Ah, expect, rtol=1e-10, atol=1e-12)


def test_reference_analytic_example_1():
    """test/integration/test_integration_poisson.py:121-143: squared L2 error < 5e-5 on 15x15 / 15x15,
    and the discrete A_hom values of SURVEY 8c.1."""

    def f(x):
        return pufl.pi**2 * (1 / 2 + 1 / pufl.sqrt(3)) * pufl.sin(pufl.pi * x[0]) * pufl.sin(pufl.pi * x[1])

    s = PoissonHMM(mesh.create_unit_square(15, 15), Cf.analytic1(pufl), f, mesh.create_unit_square(15, 15), 0.1 / 15,
                   petsc_options_cell_problem={"ksp_atol": 1e-10})  # fmt: skip
    u = s.solve()
    err = ho.l2_error_squared(omesh.create_unit_square(15, 15), u.x.array, lambda x: np.sin(np.pi * x[0]) * np.sin(np.pi * x[1]))
    assert err < 5e-5
    Ah = s.cell_tensors(np.array([[0.3, 0.4, 0.0]]))[0]
    assert abs(Ah[0, 0] - 0.500975006) < 2e-9 and abs(Ah[1, 1] - 1 / math.sqrt(3)) < 1e-9 and abs(Ah[0, 1]) < 1e-12


def test_x_only_coefficient_is_exact():
    """test_integration_poisson.py:398-473: A = 1.1 + x0 has no y-dependence -> A_hom = (1.1 + c_T0) I."""
    s = PoissonHMM(mesh.create_unit_square(3, 3), Cf.x_only(pufl), lambda x: 1.0, mesh.create_unit_square(7, 7), 0.01)
    x = np.array([[0.37, 0.1, 0.0], [0.9, 0.5, 0.0]])
    Ah = s.cell_tensors(x)
    for k in range(2):
        assert np.allclose(Ah[k], (1.1 + x[k, 0]) * np.eye(2), atol=1e-14)


def test_constant_hooke_equals_plain_fem():
    """test/integration/test_integration_linear_elasticity.py:205-322 (reference tolerance 1e-4)."""
    m = mesh.create_box((0, 0, 0), (1.0, 0.2, 0.2), (4, 2, 2))
    s = LinearElasticityHMM(m, Cf.hooke_const_3d(pufl), lambda x: pufl.as_vector([0.0, 0.0, -1.0]), mesh.create_unit_cube(3, 3, 3), 1.0)
    s._assemble_stiffness()
    Cc = np.asarray(Cf.hooke_const_3d(npufl)(np.zeros(3), np.zeros((3, 1))))[..., 0]
    K = s._A.toarray()
    Kf = np.zeros_like(K)
    for nodes in m.cells:
        verts = m.x[nodes]
        G = ho.p1_gradients(verts)
        E = []
        for a in range(4):
            for k in range(3):
                g = np.zeros((3, 3))
                g[k, :] = G[:, a]
                E.append(0.5 * (g + g.T))
        E = np.array(E)
        dofs = ho.unroll_dofs(nodes, 3)
        Kf[np.ix_(dofs, dofs)] += ho.simplex_volume(verts) * np.einsum("ijkl,akl,bij->ab", Cc, E, E)
    assert np.linalg.norm(K - Kf) <= 1e-12 * np.linalg.norm(Kf)


def test_elasticity_stratified_fibres_match_oracle():
    """BASELINE config 4 scaled down: rotated fibres (examples/linear_elasticity/rotated_fibers.py)."""
    m = mesh.create_box((0, 0, 0), (1.0, 0.4, 0.1), (3, 2, 1))
    n = 4
    f = lambda x: pufl.as_vector([0.0, 0.0, -0.05 * 0.4**2])  # noqa: E731  rotated_fibers.py:13-15,88
    s = LinearElasticityStratifiedHMM(m, Cf.hooke_fibre_3d(pufl), f, mesh.create_unit_cube(n, n, n), 0.01,
                                      Cf.dtheta_rotation_3d(pufl), petsc_options_cell_problem=TIGHT)  # fmt: skip
    V = s.function_space
    clamp = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 0.0))  # rotated_fibers.py:102-115
    s.set_boundary_conditions(fem.dirichletbc(np.zeros(3), clamp, V))
    u = s.solve()
    macro = omesh.create_box([0, 0, 0], [1.0, 0.4, 0.1], [3, 2, 1])
    mic = ho.MicroCell(omesh.create_unit_cube(n, n, n), "elasticity", 0)
    Ao = ho.assemble_macro(macro, mic, Cf.hooke_fibre_3d(npufl), 0.01, _np_dtheta("dtheta_rotation_3d"))
    assert np.linalg.norm((s._A - Ao).toarray()) <= 1e-10 * np.linalg.norm(Ao.toarray())
    b = ho.assemble_rhs(macro, lambda x: np.array([0.0, 0.0, -0.05 * 0.4**2])[:, None] + 0 * x[:1], 3, degree=1)
    dofs = ho.unroll_dofs(clamp, 3)
    uo = ho.solve_dirichlet(Ao, b, dofs, np.zeros(len(dofs)))
    assert np.abs(u.x.array - uo).max() <= 1e-8 * np.abs(uo).max()
    assert s.cell_iterations.max() < 10000
    st = s.assembly_stats  # device timings and iteration counts of the last assembly
    assert st["macro_cells"] == m.num_cells and st["cell_kernel_ms"] > 0 and st["rhs_iterations"] >= s.cell_iterations.sum()


def test_cell_solver_choice_direct_vs_pcg():
    """K5 behind the drop-in classes: BASELINE config 4's micro cell (8^3, fibre along y0 -> collapsed to 192
    unknowns, ~220 PCG iterations) is factorised directly under cell_solver="auto"; "pcg" and "direct" give the
    same macro matrix to 1e-10; an easy coefficient stays on PCG."""
    m = mesh.create_box((0, 0, 0), (1.0, 0.4, 0.1), (4, 2, 1))
    f = lambda x: pufl.as_vector([0.0, 0.0, -0.05 * 0.4**2])  # noqa: E731
    mk = lambda how, n=8, coeff=Cf.hooke_fibre_3d: LinearElasticityStratifiedHMM(  # noqa: E731
        m, coeff(pufl), f, mesh.create_unit_cube(n, n, n), 0.01, Cf.dtheta_rotation_3d(pufl),
        petsc_options_cell_problem=TIGHT, cell_solver=how)
    mats = {}
    for how in ("pcg", "direct", "auto"):
        s = mk(how)
        s._assemble_stiffness()
        mats[how] = sp_values(s)
        assert s.cell_solver_used == ("pcg" if how == "pcg" else "direct")
        assert (s.cell_iterations.max() == 0) == (how != "pcg")
    ref = np.abs(mats["pcg"]).max()
    assert np.abs(mats["direct"] - mats["pcg"]).max() <= 1e-10 * ref
    assert np.array_equal(mats["auto"], mats["direct"])
    easy = mk("auto", 4, Cf.hooke_smooth_3d)
    easy._assemble_stiffness()
    assert easy.cell_solver_used == "pcg"
    with pytest.raises(ValueError):
        mk("lu")
    big = LinearElasticityStratifiedHMM(m, Cf.hooke_fibre_3d(pufl), f, mesh.create_unit_cube(8, 8, 8), 0.01, Cf.dtheta_rotation_3d(pufl),
                                        collapse_invariant_axes=False, cell_solver="direct")  # fmt: skip
    with pytest.raises(Exception):
        big._assemble_stiffness()  # 1,536 unknowns do not fit the direct kernel: loud, no fallback


def sp_values(s):
    return np.array(s._dev["vals"].cpu().numpy(), copy=True)


def test_hmm_equals_periodic_homogenisation():
    """test/integration/test_integration_poisson.py:188-240: for A = A(y) PoissonHMM and PoissonPeriodicHMM
    agree: macro matrices (Frobenius < 1e-8) and solutions (L2 < 1e-12 in the reference, with LU)."""
    nm, n = 8, 8
    A = Cf.periodic_only(pufl)
    m = mesh.create_unit_square(nm, nm)
    hmm = PoissonHMM(m, A, lambda x: 1.0, mesh.create_unit_square(n, n), 0.1 / n, petsc_options_cell_problem=TIGHT)
    per = PoissonPeriodicHMM(m, lambda y: 2.0 + pufl.sin(2 * pufl.pi * y[0]), lambda x: 1.0, mesh.create_unit_square(n, n), 0.1 / n)
    per.set_boundary_conditions(hmm._bcs)
    u1, u2 = hmm.solve(), per.solve()
    assert np.linalg.norm((hmm._A - per._A).toarray()) < 1e-8
    assert np.abs(u1.x.array - u2.x.array).max() < 1e-12
    mic = ho.MicroCell(omesh.create_unit_square(n, n), "poisson", 3)
    Ao = ho.cell_tensor(mic, Cf.periodic_only(npufl), [0.0, 0.0, 0.0])
    assert np.abs(per.A_hom - Ao).max() <= 1e-10 * np.abs(Ao).max()
    # correctors on the micro mesh vertices (slaves carry their master's value), up to a constant
    _, chis = ho.cell_tensor(mic, Cf.periodic_only(npufl), [0.0, 0.0, 0.0], return_correctors=True)
    for q in range(2):
        got = per.correctors[q].x.array
        want = chis[q][mic.node2per]
        assert np.abs((got - got.mean()) - (want - want.mean())).max() < 1e-9


def test_device_macro_solve_matches_host_solve():
    """lifting + Jacobi-PCG on the GPU against the scipy direct solve, same assembled matrix."""
    m = mesh.create_unit_cube(6, 5, 4)
    mk = lambda opts: PoissonHMM(m, Cf.smooth_sin(pufl), lambda x: 1.0 + x[0], mesh.create_unit_cube(4, 4, 4), 0.125,
                                 petsc_options_global_solve=opts, petsc_options_cell_problem=TIGHT)  # noqa: E731
    a, b = mk(None), mk({"pc_type": "lu"})
    V = a.function_space
    top = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 1.0))
    for s in (a, b):
        s.set_boundary_conditions([s._bcs[0], fem.dirichletbc(0.25, top, V)])  # overlapping conditions: lifted one by one as in hmm.py:453-480
    ua, ub = a.solve(), b.solve()
    assert a.macro_solve_stats["iterations"] > 0 and b.macro_solve_stats is None
    assert np.abs(ua.x.array - ub.x.array).max() <= 1e-9 * np.abs(ub.x.array).max()
    # elasticity with a clamped face
    me = mesh.create_box((0, 0, 0), (1.0, 0.4, 0.1), (4, 2, 2))
    f = lambda x: pufl.as_vector([0.0, 0.0, -1.0])  # noqa: E731
    mk = lambda opts: LinearElasticityHMM(me, Cf.hooke_smooth_3d(pufl), f, mesh.create_unit_cube(4, 4, 4), 0.05,
                                          petsc_options_global_solve=opts, petsc_options_cell_problem=TIGHT)  # noqa: E731
    a, b = mk(None), mk({"ksp_type": "preonly", "pc_type": "lu"})
    clamp = fem.locate_dofs_geometrical(a.function_space, lambda x: np.isclose(x[0], 0.0))
    for s in (a, b):
        s.set_boundary_conditions(fem.dirichletbc(np.zeros(3), clamp, s.function_space))
    ua, ub = a.solve(), b.solve()
    assert np.abs(ua.x.array - ub.x.array).max() <= 1e-8 * np.abs(ub.x.array).max()
    # the device solve reduces its dot products in fixed order (block partials, no floating-point atomics): a second
    # solve of the same system returns the same bits
    first = ua.x.array.copy()
    again = a.solve().x.array
    assert np.array_equal(first, again)


@pytest.mark.parametrize("dim,bs", [(2, 1), (3, 1), (3, 3)])
def test_device_load_vector_matches_host_assembly(dim, bs):
    """SURVEY 8f row 2 / hmm.py:445-450: the macro load vector assembled on the device (generated f kernel +
    deterministic gather, hmx_macro_load_dev) against the host assembly of hommx_b200.fem."""
    from hommx_b200 import fem

    if dim == 2:
        m, mic = mesh.create_rectangle((0.0, 0.0), (1.0, 0.7), (9, 7)), mesh.create_unit_square(4, 4)
        f = lambda x: 1.0 + pufl.sin(3 * x[0]) * x[1] ** 2  # noqa: E731
        s = PoissonHMM(m, Cf.smooth_sin(pufl), f, mic, 0.1)
    elif bs == 1:
        m, mic = mesh.create_box((0.0, 0.0, 0.0), (1.0, 0.5, 0.3), (5, 4, 3)), mesh.create_unit_cube(4, 4, 4)
        f = lambda x: pufl.cos(x[0] + 2 * x[2]) + x[1]  # noqa: E731
        s = PoissonHMM(m, Cf.smooth_sin(pufl), f, mic, 0.1)
    else:
        m, mic = mesh.create_box((0.0, 0.0, 0.0), (1.0, 0.4, 0.1), (6, 3, 2)), mesh.create_unit_cube(4, 4, 4)
        f = lambda x: pufl.as_vector([0.0, x[0] * pufl.sin(x[1]), -0.05 * 0.4**2])  # noqa: E731
        s = LinearElasticityHMM(m, Cf.hooke_smooth_3d(pufl), f, mic, 0.1)
    s._ensure_solver()
    b_dev = s._assemble_load_device().cpu().numpy()
    b_host = fem.assemble_load(s._V_macro, f)
    assert np.abs(b_dev - b_host).max() <= 1e-13 * np.abs(b_host).max()
    s.set_right_hand_side(lambda x: 2.0 if bs == 1 else pufl.as_vector([1.0, 0.0, 0.0]))  # a new f loads a new module
    b2 = s._assemble_load_device().cpu().numpy()
    assert np.abs(b2 - fem.assemble_load(s._V_macro, s._f)).max() <= 1e-13 * np.abs(b2).max()


@pytest.mark.parametrize("kind", ["poisson", "elasticity"])
def test_higher_order_macro_quadrature(kind):
    """SURVEY 8f row 4 (second half): macro_quadrature_degree=k solves the cell problem at every point of the degree-k
    simplex rule and assembles S_loc = |T| C^T (sum_q w_q A_hom(x_q)) C (hmx_macro_elements_dev).  Checked against the
    oracle's tensors at the same points; degree 1 / None is the reference's barycentre rule."""
    from hommx_b200 import quadrature
    from oracle import hmm_oracle as ho

    if kind == "poisson":
        m, mic, n = mesh.create_rectangle((0.0, 0.0), (1.0, 0.8), (3, 2)), mesh.create_unit_square(8, 8), 8
        A, An = Cf.analytic2(pufl), Cf.analytic2(npufl)  # 0.33 + 0.15 (sin 2 pi x0 + sin 2 pi y0): nonlinear in x
        mk = lambda **kw: PoissonHMM(m, A, lambda x: 1.0, mic, 0.1, petsc_options_cell_problem=TIGHT, **kw)  # noqa: E731
        omic = ho.MicroCell(mic, "poisson", 4)
    else:
        m, mic, n = mesh.create_box((0.0, 0.0, 0.0), (1.0, 0.5, 0.25), (2, 1, 1)), mesh.create_unit_cube(4, 4, 4), 4
        A, An = Cf.hooke_smooth_3d(pufl), Cf.hooke_smooth_3d(npufl)
        mk = lambda **kw: LinearElasticityHMM(m, A, lambda x: pufl.as_vector([0.0, 0.0, -1.0]), mic, 0.1,
                                              petsc_options_cell_problem=TIGHT, **kw)  # noqa: E731
        omic = None
    base, hi = mk(), mk(macro_quadrature_degree=3)
    base._assemble_stiffness()
    hi._assemble_stiffness()
    d = m.dim
    pts, wts = quadrature.default_rule(d, 3)
    lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)
    wts = wts / wts.sum()
    S = hi._dev["S"].cpu().numpy().reshape(m.num_cells, hi._num_basis_functions_per_cell, -1)
    for c in (0, m.num_cells - 1):
        verts = m.x[m.cells[c]]
        xq = lam @ verts
        if omic is not None:
            Aq = np.array([ho.cell_tensor(omic, An, x) for x in xq])
        else:
            Aq = hi.cell_tensors(xq)  # (elasticity: the kernel's own tensors, pinned against the oracle elsewhere)
        ref = ho.local_stiffness_from_tensor(np.einsum("q,qij->ij", wts, Aq), verts, kind)
        assert np.abs(S[c] - ref).max() <= 1e-9 * np.abs(ref).max()
    # the rule matters for a coefficient that is nonlinear in x, and the barycentre default is untouched
    if kind == "poisson":
        assert np.abs(hi._A_values - base._A_values).max() > 1e-4 * np.abs(base._A_values).max()
    u = hi.solve()
    assert np.isfinite(u.x.array).all()
