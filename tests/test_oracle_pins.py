"""Pins the CPU oracle on the known answers the reference's own tests encode
 -- the reference itself cannot run in this image."""
import math

import numpy as np
import pytest

import coefficients as C
from oracle import hmm_oracle as ho
from oracle import meshes
from oracle import npufl as ufl


def test_periodic_identification_2d():
    """/root/reference/test/unit/test_unit.py:25-55"""
    msh = meshes.create_unit_square(10, 10)
    master = meshes.periodic_master_map(msh)
    x = msh.x
    boundary = np.isclose(x[:, :2], 0).any(axis=1) | np.isclose(x[:, :2], 1).any(axis=1)
    for v in range(len(x)):
        if not boundary[v]:
            assert master[v] == v
        if np.allclose(x[v], [1, 1, 0]):
            assert np.allclose(x[master[v]], [0, 0, 0])
            continue
        if master[v] != v:
            assert boundary[master[v]]
            diff = np.abs(x[master[v]] - x[v])
            assert np.allclose(diff, [1, 0, 0]) or np.allclose(diff, [0, 1, 0])


def test_periodic_identification_3d():
    """/root/reference/test/unit/test_unit.py:57-103"""
    msh = meshes.create_unit_cube(10, 10, 10)
    master = meshes.periodic_master_map(msh)
    x = msh.x
    boundary = np.isclose(x, 0).any(axis=1) | np.isclose(x, 1).any(axis=1)
    n_slaves = 0
    for v in range(len(x)):
        if not boundary[v]:
            assert master[v] == v
            continue
        if np.allclose(x[v], [1, 1, 1]):
            assert np.allclose(x[master[v]], [0, 0, 0])
            n_slaves += 1
            continue
        handled = False
        for i, j in ((0, 1), (0, 2), (1, 2)):
            if np.allclose(x[v][[i, j]], [1, 1]):
                assert master[v] != v
                assert np.allclose(x[master[v]][[i, j]], [0, 0])
                handled = True
        if handled:
            n_slaves += 1
            continue
        if master[v] != v:
            n_slaves += 1
            diff = np.abs(x[master[v]] - x[v])
            assert any(np.allclose(diff, e) for e in np.eye(3))
    assert n_slaves == 11**3 - 10**3


def test_analytic_example_1_cell_tensor():
    """test_integration_poisson.py:121-143: A = 1/(2+cos 2 pi y0) -> continuum diag(1/2, 1/sqrt 3).
    Discrete P1 values (6-point rule) from SURVEY.md 8c.1 (independent survey computation)."""
    for n, a00 in ((15, 0.500975006), (30, 0.500244587)):
        mic = ho.MicroCell(meshes.create_unit_square(n, n), "poisson", 3)
        Ah = ho.cell_tensor(mic, C.analytic1(ufl), [0.3, 0.4, 0.0])
        assert abs(Ah[0, 0] - a00) < 2e-9
        assert abs(Ah[1, 1] - 1 / math.sqrt(3)) < 1e-9
        assert abs(Ah[0, 1]) < 1e-14 and abs(Ah[1, 0]) < 1e-14


def _solve_poisson_hmm(n, A, f, eps, degree):
    macro = meshes.create_unit_square(n, n)
    mic = ho.MicroCell(meshes.create_unit_square(n, n), "poisson", degree)
    Amat = ho.assemble_macro(macro, mic, A, eps)
    b = ho.assemble_rhs(macro, f, 1, degree=4)
    bc = ho.boundary_nodes(macro)
    return macro, Amat, ho.solve_dirichlet(Amat, b, bc, np.zeros(len(bc)))


def test_analytic_example_1_solution():
    """test_integration_poisson.py:121-143 (squared L2 error < 5e-5 on 15x15/15x15)"""

    def f(x):
        return ufl.pi**2 * (1 / 2 + 1 / ufl.sqrt(3)) * ufl.sin(ufl.pi * x[0]) * ufl.sin(ufl.pi * x[1])

    macro, _, u = _solve_poisson_hmm(15, C.analytic1(ufl), f, 0.1 / 15, 3)
    err = ho.l2_error_squared(macro, u, lambda x: ufl.sin(ufl.pi * x[0]) * ufl.sin(ufl.pi * x[1]))
    assert err < 5e-5


def test_analytic_example_2_solution():
    """test_integration_poisson.py:146-185"""

    def f(x):
        s = 0.454545454545455 * ufl.sin(2 * ufl.pi * x[0]) + 1
        root = ufl.sqrt(s**2 - 0.206611570247934)
        return (
            3.25696945235949 * root * ufl.sin(ufl.pi * x[0]) * ufl.sin(ufl.pi * x[1])
            + ufl.pi**2 * (0.15 * ufl.sin(2 * ufl.pi * x[0]) + 0.33) * ufl.sin(ufl.pi * x[0]) * ufl.sin(ufl.pi * x[1])
            - 2.96088132032681 * s * ufl.sin(ufl.pi * x[1]) * ufl.cos(ufl.pi * x[0]) * ufl.cos(2 * ufl.pi * x[0]) / root
        )

    macro, _, u = _solve_poisson_hmm(15, C.analytic2(ufl), f, 0.1 / 15, 3)
    err = ho.l2_error_squared(macro, u, lambda x: ufl.sin(ufl.pi * x[0]) * ufl.sin(ufl.pi * x[1]))
    assert err < 5e-5


def test_literal_equals_tensor_formulation_poisson():
    """test_integration_poisson.py:188-240 pins 'n_b correctors == A_hom formulation' (< 1e-8 Frobenius)."""
    n = 8
    macro = meshes.create_unit_square(n, n)
    mic = ho.MicroCell(meshes.create_unit_square(n, n), "poisson", 3)
    A = C.periodic_only(ufl)
    K1 = ho.assemble_macro(macro, mic, A, 0.1 / n, literal=True).toarray()
    K2 = ho.assemble_macro(macro, mic, A, 0.1 / n, literal=False).toarray()
    assert np.linalg.norm(K1 - K2) < 1e-10
    # the periodic class: one tensor for the whole mesh, constant-coefficient FEM matrix
    Ah = ho.cell_tensor(mic, A, [0.0, 0.0, 0.0])
    K3 = np.zeros_like(K2)
    for nodes in macro.cells:
        S = ho.local_stiffness_from_tensor(Ah, macro.x[nodes], "poisson")
        K3[np.ix_(nodes, nodes)] += S
    assert np.linalg.norm(K3 - K1) < 1e-8


@pytest.mark.parametrize("dim", [2, 3])
def test_literal_equals_tensor_formulation_stratified(dim):
    if dim == 2:
        mic = ho.MicroCell(meshes.create_unit_square(9, 9), "poisson", 3)
        verts = np.array([[0.1, 0.2, 0], [0.3, 0.25, 0], [0.15, 0.5, 0]])
        A, Dt = C.full_tensor_2d(ufl), C.dtheta_test_stratified(ufl)
    else:
        mic = ho.MicroCell(meshes.create_unit_cube(4, 4, 4), "poisson", 3)
        verts = np.array([[0.1, 0.2, 0.3], [0.3, 0.25, 0.3], [0.15, 0.5, 0.35], [0.2, 0.3, 0.6]])
        A, Dt = C.full_tensor_3d(ufl), C.dtheta_shear_3d(ufl)
    M = lambda x: np.asarray(Dt(x))[..., 0]
    S1 = ho.local_stiffness_literal(mic, A, verts, 1 / 150, M)
    c = verts.mean(axis=0)
    S2 = ho.local_stiffness_from_tensor(ho.cell_tensor(mic, A, c, M(c)), verts, "poisson")
    assert np.abs(S1 - S2).max() < 1e-12 * np.abs(S2).max()


@pytest.mark.parametrize("dim", [2, 3])
def test_literal_equals_tensor_formulation_elasticity(dim):
    if dim == 2:
        mic = ho.MicroCell(meshes.create_unit_square(6, 6), "elasticity", 3)
        verts = np.array([[0.1, 0.2, 0], [0.3, 0.25, 0], [0.15, 0.5, 0]])
        A, Dt = C.hooke_sin_2d(ufl), C.dtheta_test_stratified(ufl)
    else:
        mic = ho.MicroCell(meshes.create_unit_cube(3, 3, 3), "elasticity", 3)
        verts = np.array([[0.1, 0.2, 0.3], [0.3, 0.25, 0.3], [0.15, 0.5, 0.35], [0.2, 0.3, 0.6]])
        A, Dt = C.hooke_smooth_3d(ufl), C.dtheta_shear_3d(ufl)
    M = lambda x: np.asarray(Dt(x))[..., 0]
    S1 = ho.local_stiffness_literal(mic, A, verts, 1 / 64, M)
    c = verts.mean(axis=0)
    S2 = ho.local_stiffness_from_tensor(ho.cell_tensor(mic, A, c, M(c)), verts, "elasticity")
    assert np.abs(S1 - S2).max() < 1e-12 * np.abs(S2).max()


def test_x_only_coefficient():
    """test_integration_poisson.py:398-473: A = 1.1 + x0 -> A_hom(c_T) = (1.1 + c_T0) I exactly."""
    mic = ho.MicroCell(meshes.create_unit_square(7, 7), "poisson", 1)
    Ah = ho.cell_tensor(mic, C.x_only(ufl), [0.37, 0.1, 0.0])
    assert np.allclose(Ah, 1.47 * np.eye(2), atol=1e-14)


def test_laminate_closed_form():
    """scalar a(y0), any M: A_hom = <a>(I - n n^T) + <1/a>^-1 n n^T, n = M e0/|M e0|;
    exact for the P1 discretisation when the jumps sit on grid lines (laminate.py:101-102, n % 4 == 0)."""
    mic = ho.MicroCell(meshes.create_unit_square(8, 8), "poisson", 0)
    x = np.array([0.3, 0.6, 0.0])
    M = np.asarray(C.dtheta_wavy(ufl)(x))[..., 0]
    Ah = ho.cell_tensor(mic, C.laminate(ufl), x, M)
    nvec = M[:, 0] / np.linalg.norm(M[:, 0])
    arith, harm = 2.525, 1.0 / (0.5 / 5 + 0.5 / 0.05)
    expected = arith * (np.eye(2) - np.outer(nvec, nvec)) + harm * np.outer(nvec, nvec)
    assert np.allclose(Ah, expected, rtol=1e-12, atol=1e-13)


def test_constant_hooke_equals_plain_fem():
    """test_integration_linear_elasticity.py:205-322: constant Hooke tensor -> the HMM matrix equals
    the plain P1 elasticity stiffness matrix (reference tolerance 1e-4 relative)."""
    macro = meshes.create_box([0, 0, 0], [1.0, 0.2, 0.2], [4, 2, 2])
    mic = ho.MicroCell(meshes.create_unit_cube(3, 3, 3), "elasticity", 0)
    A = C.hooke_const_3d(ufl)
    K = ho.assemble_macro(macro, mic, A, 1.0).toarray()
    # plain FEM: S_loc = |T| e(phi_i):C:e(phi_j)
    Cc = np.asarray(A(np.zeros(3), np.zeros((3, 1))))[..., 0]
    Kf = np.zeros_like(K)
    for nodes in macro.cells:
        verts = macro.x[nodes]
        G = ho.p1_gradients(verts)
        vol = ho.simplex_volume(verts)
        E = []
        for a in range(4):
            for k in range(3):
                g = np.zeros((3, 3))
                g[k, :] = G[:, a]
                E.append(0.5 * (g + g.T))
        E = np.array(E)
        S = vol * np.einsum("ijkl,akl,bij->ab", Cc, E, E)
        dofs = ho.unroll_dofs(nodes, 3)
        Kf[np.ix_(dofs, dofs)] += S
    assert np.linalg.norm(K - Kf) / np.linalg.norm(Kf) < 1e-12


def test_analytic_example_1_discrete_convergence():
    """SURVEY.md 8c.1 (independent survey computation): the discrete A_hom[0,0] of
    A = 1/(2+cos 2 pi y0) converges like O(h^2) to 1/2; n = 60 gives 0.500061199."""
    mic = ho.MicroCell(meshes.create_unit_square(60, 60), "poisson", 3)
    Ah = ho.cell_tensor(mic, C.analytic1(ufl), [0.1, 0.9, 0.0])
    assert abs(Ah[0, 0] - 0.500061199) < 2e-9
    assert abs(Ah[1, 1] - 1 / math.sqrt(3)) < 1e-9


def test_hmm_equals_periodic_class_formula():
    """BasePeriodicHMM (hmm.py:1199-1245, 1274-1279): A_hom[p,q] = 1/|Y| int A (e_q + grad chi_q) . e_p --
    the non-symmetrised formula of the periodic class -- equals the oracle's energy form."""
    mic = ho.MicroCell(meshes.create_unit_square(12, 12), "poisson", 3)
    A = C.full_tensor_2d(ufl)
    x = np.array([0.3, 0.7, 0.0])
    Ah, chis = ho.cell_tensor(mic, A, x, return_correctors=True)
    Abar = mic.element_coefficient(A, x)
    _, B = mic.assemble(Abar, np.eye(2))
    ne = len(mic.vol)
    for q in range(2):
        tot = np.broadcast_to(np.eye(2)[q], (ne, 2)) + mic.field_of(B, chis[q])
        for p in range(2):
            flux = np.einsum("eij,ej->ei", Abar, tot)  # A (e_q + grad chi_q) per element
            val = float(np.einsum("e,ei->", mic.vol, flux * np.eye(2)[p]))
            assert abs(val / mic.Y - Ah[p, q]) < 1e-12
