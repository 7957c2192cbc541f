"""Coefficient callables shared by the oracle tests and the CUDA parity tests.

Each factory takes the ``ufl``-like namespace to build with, so the SAME
source text is evaluated by two independent back ends: ``oracle.npufl`` (numpy,
the checker) and ``hommx_b200.ufl`` (symbolic -> CUDA, the product).  The
expressions are the ones the reference's tests and examples use (file:line in
each docstring, relative to /root/reference).
"""
import math

# the BASELINE workloads' coefficients live in the package (bench.py and examples/ use them too): one source text
from hommx_b200.workloads import (  # noqa: F401
    circle_indicator as _circle,
    dtheta_rotation_3d,
    dtheta_wavy,
    hooke,
    hooke_fibre_3d,
    hooke_spheres_3d,
    laminate,
    smooth_sin,
)




def analytic1(ufl):
    """test/integration/test_integration_poisson.py:124-125"""

    def A(x, y):
        return 1 / (2 + ufl.cos(2 * ufl.pi * y[0]))

    return A


def analytic2(ufl):
    """test/integration/test_integration_poisson.py:149-150"""

    def A(x, y):
        return 0.33 + 0.15 * (ufl.sin(2 * ufl.pi * x[0]) + ufl.sin(2 * ufl.pi * y[0]))

    return A


def periodic_only(ufl):
    """test/integration/test_integration_poisson.py:191-196"""

    def A(x, y):
        return 2.0 + ufl.sin(2 * ufl.pi * y[0])

    return A


def x_only(ufl):
    """test/integration/test_integration_poisson.py:401-402"""

    def A(x, y):
        return 1.1 + x[0]

    return A






def inclusion(ufl):
    """examples/diffusion/inclusion.py:117-118"""

    def A(x, y):
        return ufl.conditional(_circle(ufl, y[0], y[1]), 0.001, 0.1)

    return A


def full_tensor_2d(ufl):
    """matrix-valued, x- and y-dependent, symmetric (not in the reference's tests; exercises as_matrix)"""

    def A(x, y):
        a00 = 1.1 + x[0] + ufl.sin(2 * ufl.pi * y[0])
        a01 = 0.2 * ufl.cos(2 * ufl.pi * y[1])
        a11 = 2 + ufl.sin(2 * ufl.pi * (y[0] + y[1]))
        return ufl.as_matrix([[a00, a01], [a01, a11]])

    return A


def full_tensor_3d(ufl):
    def A(x, y):
        a = 1.5 + x[1] + ufl.sin(2 * ufl.pi * y[0]) * ufl.cos(2 * ufl.pi * y[2])
        b = 2.0 + 0.5 * ufl.cos(2 * ufl.pi * y[1])
        c = 1.0 + 0.3 * ufl.sin(2 * ufl.pi * (y[0] + y[1] + y[2]))
        o = 0.1 * ufl.sin(2 * ufl.pi * y[1])
        return ufl.as_matrix([[a, o, 0.05], [o, b, 0.0], [0.05, 0.0, c]])

    return A


# ---- stratification Jacobians (transposed): M[p, i] = d theta_i / d x_p ----
def dtheta_test_stratified(ufl, theta_factor=0.2):
    """test/integration/test_integration_poisson.py:498-508"""

    def Dtheta(x):
        arg_0 = ufl.pi / 2 * x[0]
        arg_1 = ufl.pi / 2 * x[1]
        f = theta_factor * ufl.cos(arg_0) * ufl.cos(arg_1)
        df_dx0 = -theta_factor * (ufl.pi / 2) * ufl.sin(arg_0) * ufl.cos(arg_1)
        df_dx1 = -theta_factor * (ufl.pi / 2) * ufl.cos(arg_0) * ufl.sin(arg_1)
        return ufl.as_matrix(
            [[1 - x[1] * df_dx0, f + x[0] * df_dx0], [-f - x[1] * df_dx1, 1 + x[0] * df_dx1]]
        )

    return Dtheta




def dtheta_inclusion(ufl):
    """examples/diffusion/inclusion.py:128-134"""

    def Dtheta(x):
        D = ufl.as_matrix([[1, 0.5 * 2 * ufl.pi * ufl.cos(2 * ufl.pi * x[1])], [0, 1]])
        return ufl.transpose(D)

    return Dtheta




def dtheta_shear_3d(ufl):
    def Dtheta(x):
        return ufl.as_matrix(
            [[1.0 + 0.2 * x[0], 0.3, 0.0], [-0.1, 1.1, 0.2 * ufl.sin(x[1])], [0.05, 0.0, 0.9 + 0.1 * x[2]]]
        )

    return Dtheta


# ---- elasticity ----


def hooke_sin_2d(ufl):
    """test_integration_linear_elasticity.py:78-82"""
    return hooke(ufl, 2, lambda x, y: 5 + 4.5 * ufl.sin(2 * ufl.pi * y[0]), lambda x, y: 1.25)


def hooke_const_3d(ufl):
    """test_integration_linear_elasticity.py:228-232"""
    return hooke(ufl, 3, lambda x, y: 1, lambda x, y: 1.25)




def hooke_smooth_3d(ufl):
    return hooke(
        ufl,
        3,
        lambda x, y: 2 + x[0] + ufl.sin(2 * ufl.pi * y[0]) * ufl.cos(2 * ufl.pi * y[1]),
        lambda x, y: 1.25 + 0.5 * ufl.cos(2 * ufl.pi * y[2]),
    )


def cubic_3d(ufl):
    """Anisotropic (cubic symmetry) elasticity tensor with three independent, y- and x-dependent moduli, written
    component by component: exercises the general 21-component path of the coefficient front end (the reference's
    tests only use isotropic Hooke tensors)."""

    def A(x, y):
        c11 = 3.0 + ufl.sin(2 * ufl.pi * y[0]) * ufl.cos(2 * ufl.pi * y[1])
        c12 = 1.0 + 0.3 * x[0]
        c44 = 0.8 + 0.2 * ufl.cos(2 * ufl.pi * y[2])
        d = lambda a, b: 1.0 if a == b else 0.0  # noqa: E731

        def comp(i, j, k, l):
            iso = c12 * (d(i, j) * d(k, l)) + c44 * (d(i, k) * d(j, l) + d(i, l) * d(j, k))
            if i == j == k == l:
                return iso + (c11 - c12 - 2 * c44)
            return iso

        return ufl.as_tensor([[[[comp(i, j, k, l) for l in range(3)] for k in range(3)] for j in range(3)] for i in range(3)])

    return A
