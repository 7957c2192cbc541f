"""Simplex quadrature tables used by the oracle.

The reference integrates every form with basix's default scheme at the degree
UFL estimates (call sites /root/reference/src/hommx/hmm.py:
645-667,364).  basix is not in this image; tables that could be restated are
restated, the rest fall back to a collapsed Gauss-Jacobi rule of the same
degree (what basix itself uses for ``QuadratureType.gauss_jacobi``).  Points
are on the reference simplex (vertices 0, e_1, .., e_d); weights sum to 1/d!.
"""
import numpy as np
from scipy.special import roots_jacobi


def _gauss_jacobi_simplex(dim, degree):
    m = (degree + 2) // 2
    pts1 = []
    for a in range(dim):
        xa, wa = roots_jacobi(m, float(a), 0.0)
        pts1.append((0.5 * (xa + 1.0), wa * 0.5 ** (a + 1)))
    if dim == 2:
        (x0, w0), (x1, w1) = pts1[0], pts1[1]
        P, W = [], []
        for i in range(m):
            for j in range(m):
                P.append([x1[i], x0[j] * (1.0 - x1[i])])
                W.append(w1[i] * w0[j])
        return np.array(P), np.array(W)
    (x0, w0), (x1, w1), (x2, w2) = pts1
    P, W = [], []
    for i in range(m):
        for j in range(m):
            for k in range(m):
                P.append([x2[i], x1[j] * (1.0 - x2[i]), x0[k] * (1.0 - x1[j]) * (1.0 - x2[i])])
                W.append(w2[i] * w1[j] * w0[k])
    return np.array(P), np.array(W)


def simplex_rule(dim, degree):
    """Return (points (nq, dim), weights (nq,)) exact to ``degree``."""
    if degree <= 1:
        return np.full((1, dim), 1.0 / (dim + 1)), np.array([1.0 / (2 if dim == 2 else 6)])
    if dim == 2 and degree == 2:
        p = np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]])
        return p, np.full(3, 1 / 6)
    if dim == 2 and degree in (3, 4):
        a, wa = 0.4459484909159649, 0.11169079483900574
        b, wb = 0.09157621350977074, 0.05497587182766094
        p = np.array(
            [[a, a], [a, 1 - 2 * a], [1 - 2 * a, a], [b, b], [b, 1 - 2 * b], [1 - 2 * b, b]]
        )
        return p, np.array([wa, wa, wa, wb, wb, wb])
    if dim == 3 and degree == 2:
        a, b = 0.1381966011250105, 0.5854101966249685
        p = np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]])
        return p, np.full(4, 1 / 24)
    return _gauss_jacobi_simplex(dim, degree)
