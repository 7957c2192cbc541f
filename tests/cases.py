"""Parity cases shared by the CPU-emulation tests (`not gpu`) and the CUDA tests (`gpu`).

Each case names a coefficient from tests/coefficients.py (the expressions the reference's tests
and examples use), a micro mesh size and optionally a stratification Jacobian.  ``build``
returns everything both back ends need: the traced coefficient program (product side), the
numpy callables (oracle side) and the quadrature table.
"""
from dataclasses import dataclass

import numpy as np

import coefficients as Cf
from hommx_b200 import codegen, micro, quadrature
from hommx_b200 import ufl as pufl
from oracle import hmm_oracle as ho
from oracle import meshes as omesh
from oracle import npufl


@dataclass
class Case:
    name: str
    dim: int
    kind: int  # 0 poisson, 1 elasticity
    n: int
    coeff: str
    dtheta: str = None
    rtol: float = 1e-10
    tol: float = 1e-10  # relative parity tolerance on A_hom (north star: 1e-10)
    threads: int = None
    heavy: bool = False  # too slow for the CPU emulation


CASES = [
    Case("p2_smooth_n15", 2, 0, 15, "smooth_sin"),
    Case("p2_smooth_n16_c1", 2, 0, 16, "smooth_sin"),  # BASELINE config 1 micro cell
    Case("p2_analytic1_n15", 2, 0, 15, "analytic1"),
    Case("p2_analytic2_n15", 2, 0, 15, "analytic2"),
    Case("p2_periodic_n8", 2, 0, 8, "periodic_only"),
    Case("p2_smooth_n8", 2, 0, 8, "smooth_sin"),
    Case("p2_laminate_wavy_n8", 2, 0, 8, "laminate", "dtheta_wavy"),
    Case("p3_smooth_n4", 3, 0, 4, "smooth_sin"),
    Case("p2_xonly_n7", 2, 0, 7, "x_only"),
    Case("p2_laminate_wavy_n16", 2, 0, 16, "laminate", "dtheta_wavy"),
    Case("p2_laminate_wavy_n32_c2", 2, 0, 32, "laminate", "dtheta_wavy"),  # BASELINE config 2
    Case("p2_inclusion_n16", 2, 0, 16, "inclusion", "dtheta_inclusion"),
    Case("p2_fulltensor_strat_n9", 2, 0, 9, "full_tensor_2d", "dtheta_test_stratified"),
    Case("p2_smooth_strat_n12", 2, 0, 12, "smooth_sin", "dtheta_test_stratified"),
    Case("p3_smooth_n8_c3", 3, 0, 8, "smooth_sin"),  # BASELINE config 3
    Case("p3_fulltensor_shear_n5", 3, 0, 5, "full_tensor_3d", "dtheta_shear_3d"),
    Case("p3_fulltensor_n6", 3, 0, 6, "full_tensor_3d", threads=64),
    Case("p3_smooth_n6_lines", 3, 0, 6, "smooth_sin"),  # default threads: 96, of which 72 own a piece of a grid line
    Case("p2_inclusion_n24_lines", 2, 0, 24, "inclusion", "dtheta_inclusion"),  # 160 threads, 144 line pieces
    Case("p3_smooth_n10_lines", 3, 0, 10, "smooth_sin", threads=224, heavy=True),  # 5 nodes per thread: two register chunks in the set-up
    Case("p3_smooth_n12", 3, 0, 12, "smooth_sin", heavy=True),
    Case("p2_inclusion_n64", 2, 0, 64, "inclusion", heavy=True),
    Case("p3_fulltensor_n10_l2", 3, 0, 10, "full_tensor_3d", heavy=True),  # 4 atoms x 6000 elements: atoms in L2
    Case("e2_hooke_sin_n6", 2, 1, 6, "hooke_sin_2d"),
    Case("e2_hooke_sin_strat_n7", 2, 1, 7, "hooke_sin_2d", "dtheta_test_stratified"),
    Case("e3_hooke_const_n3", 3, 1, 3, "hooke_const_3d"),
    Case("e3_hooke_smooth_n4", 3, 1, 4, "hooke_smooth_3d"),
    Case("e3_hooke_smooth_shear_n3", 3, 1, 3, "hooke_smooth_3d", "dtheta_shear_3d"),
    Case("e3_fibre_rot_n4", 3, 1, 4, "hooke_fibre_3d", "dtheta_rotation_3d"),
    Case("e3_cubic_shear_n4", 3, 1, 4, "cubic_3d", "dtheta_shear_3d"),  # anisotropic tensor, 3 atoms, all y axes
    Case("e3_fibre_rot_n4_blocks", 3, 1, 4, "hooke_fibre_3d", "dtheta_rotation_3d", threads=192),  # 2x2x1 block sweep
    Case("e3_hooke_smooth_shear_n6", 3, 1, 6, "hooke_smooth_3d", "dtheta_shear_3d"),  # block sweep, partial warp
    Case("e3_fibre_rot_n8_c4", 3, 1, 8, "hooke_fibre_3d", "dtheta_rotation_3d", heavy=True),  # BASELINE config 4
    Case("e3_fibre_rot_n10_l2", 3, 1, 10, "hooke_fibre_3d", "dtheta_rotation_3d", rtol=1e-9, heavy=True),  # vectors in L2
]
BY_NAME = {c.name: c for c in CASES}


def program(case):
    A = getattr(Cf, case.coeff)(pufl)
    Dt = getattr(Cf, case.dtheta)(pufl) if case.dtheta else None
    return codegen.build_program(A, case.dim, case.kind, Dt)


def tables(case, prog):
    st = micro.default_structure(case.dim, case.n)
    return micro.quadrature_table(st, *quadrature.default_rule(case.dim, prog.degree))


def points(case, n_pts=3, seed=1):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.05, 0.95, (n_pts, 3))
    if case.dim == 2:
        x[:, 2] = 0.0
    return x


def oracle_degree(case):
    """Quadrature degree of the cell-problem forms by the oracle's OWN estimator (oracle/ufldegree.py), not the
    product's ``prog.degree``: a wrong degree rule in hommx_b200/ufl.py shows up as a parity failure."""
    from oracle import ufldegree

    return ufldegree.form_degree(getattr(Cf, case.coeff)(ufldegree), case.dim)


def oracle_cell(case, prog=None):
    m = omesh.create_unit_square(case.n, case.n) if case.dim == 2 else omesh.create_unit_cube(case.n, case.n, case.n)
    return ho.MicroCell(m, "poisson" if case.kind == 0 else "elasticity", oracle_degree(case))


def oracle_tensor(case, mic, x):
    A = getattr(Cf, case.coeff)(npufl)
    M = None
    if case.dtheta:
        M = np.asarray(getattr(Cf, case.dtheta)(npufl)(np.asarray(x, float)))[..., 0]
    return ho.cell_tensor(mic, A, x, M)


def random_simplices(dim, n, seed=3):
    """A few well-shaped macro cells: (cell_nodes (n, dim+1), node_xyz (n*(dim+1), 3))."""
    rng = np.random.default_rng(seed)
    xyz = []
    for _ in range(n):
        c = rng.uniform(0.2, 0.8, 3)
        ref = np.zeros((dim + 1, 3))
        ref[1:, :dim] = np.eye(dim) * 0.1
        v = c + ref + rng.uniform(-0.02, 0.02, (dim + 1, 3))
        if dim == 2:
            v[:, 2] = 0.0
        xyz.append(v)
    xyz = np.concatenate(xyz)
    cells = np.arange(n * (dim + 1), dtype=np.int32).reshape(n, dim + 1)
    return cells, xyz


def kernel_available(case):
    """Elasticity cases are skipped until csrc/hmx_cell_elasticity.cuh exists."""
    import os

    from hommx_b200 import native

    return case.kind == 0 or os.path.exists(os.path.join(native.CSRC, "hmx_cell_elasticity.cuh"))
