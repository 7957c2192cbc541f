"""Default simplex quadrature tables for hosts without basix.

FFCx integrates the reference's forms with basix's default scheme at the degree UFL estimates
(/root/reference/src/hommx/hmm.py:645-667).  The cell kernel takes the table as
a runtime input (``hmx_desc.qp/qw``) so a DOLFINx host can pass ``basix.make_quadrature``
output verbatim; these are the defaults otherwise.  Points live on the reference simplex
(vertices 0, e_1..e_d); weights sum to 1/d!.
"""
from __future__ import annotations

import numpy as np
from scipy.special import roots_jacobi


def _collapsed(dim, degree):
    """Gauss-Jacobi rule collapsed onto the simplex (basix ``QuadratureType.gauss_jacobi``)."""
    m = degree // 2 + 1
    rules = [roots_jacobi(m, float(a), 0.0) for a in range(dim)]
    nodes = [(0.5 * (x + 1.0), w / 2.0 ** (a + 1)) for a, (x, w) in enumerate(rules)]
    idx = np.indices((m,) * dim).reshape(dim, -1).T
    pts = np.empty((len(idx), dim))
    wts = np.ones(len(idx))
    for r, ix in enumerate(idx):
        # outermost coordinate uses the highest Jacobi weight
        scale = 1.0
        for lvl, i in enumerate(ix):
            a = dim - 1 - lvl
            pts[r, lvl] = nodes[a][0][i] * scale
            wts[r] *= nodes[a][1][i]
            scale *= 1.0 - nodes[a][0][i]
    return pts, wts


def default_rule(dim, degree):
    """(points (nq, dim), weights (nq,)) exact for polynomials up to ``degree``."""
    if dim not in (2, 3):
        raise ValueError("Topology should be 3D or 2D")
    fact = 2.0 if dim == 2 else 6.0
    if degree <= 1:
        return np.full((1, dim), 1.0 / (dim + 1)), np.array([1.0 / fact])
    if dim == 2 and degree == 2:
        return np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6)
    if dim == 2 and degree <= 4:  # 6-point Strang-Fix / Dunavant rule (basix Xiao-Gimbutas, degree 3-4)
        a, b = 0.4459484909159649, 0.09157621350977074
        wa, wb = 0.11169079483900574, 0.05497587182766094
        pts = [[a, a], [a, 1 - 2 * a], [1 - 2 * a, a], [b, b], [b, 1 - 2 * b], [1 - 2 * b, b]]
        return np.array(pts), np.array([wa] * 3 + [wb] * 3)
    if dim == 3 and degree == 2:
        a, b = 0.1381966011250105, 0.5854101966249685
        return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 1 / 24)
    return _collapsed(dim, degree)
