"""Host-side description of the periodic micro cell for the CUDA kernels.

The reference builds a periodic function space on ``msh_micro`` with dolfinx_mpc
(/root/reference/src/hommx/hmm.py:178-183, cell_problem.py:16-300).  The fast path is the
structured unit box (all of the reference's tests and examples use ``create_unit_square`` /
``create_unit_cube``): ``detect_structure`` recognises it from ``geometry.x`` and the cell
list, and rejects anything else -- there is no fallback for unstructured micro meshes.
"""
from __future__ import annotations

import itertools

import numpy as np

from .mesh import as_simplex_mesh

# walk order of the axes per element type, as in csrc/hmx_cell_common.cuh (kuhn_axis)
KUHN_AXES = {
    2: ((0, 1), (1, 0)),
    3: ((0, 1, 2), (0, 2, 1), (2, 0, 1), (1, 0, 2), (2, 1, 0), (1, 2, 0)),
}


def kuhn_vertices(dim):
    """(T, dim+1, dim) integer offsets of the path vertices of each element type."""
    out = []
    for axes in KUHN_AXES[dim]:
        p = np.zeros(dim, dtype=np.int64)
        verts = [p.copy()]
        for a in axes:
            p[a] += 1
            verts.append(p.copy())
        out.append(verts)
    return np.array(out)


class MicroStructure:
    """n cells per axis and, per element type, the vertex order the mesh uses (it fixes where a
    non-symmetric quadrature rule puts its points)."""

    def __init__(self, dim, n, vertex_order):
        self.dim, self.n = dim, n
        self.vertex_order = vertex_order  # (T, dim+1, dim) cube-local integer offsets, mesh order


def detect_structure(msh_micro):
    msh = as_simplex_mesh(msh_micro)
    d = msh.dim
    if d not in (2, 3):
        raise ValueError("Topology should be 3D or 2D")  # hmm.py:104-105
    x = msh.x[:, :d]
    lo, hi = x.min(axis=0), x.max(axis=0)
    if not (np.allclose(lo, 0.0) and np.allclose(hi, 1.0)):
        raise ValueError("the micro mesh must be the unit box [0,1]^d")
    nn = round(len(x) ** (1.0 / d)) - 1
    if nn < 2 or (nn + 1) ** d != len(x):
        raise ValueError("the micro mesh must be a structured n^d grid with n >= 2 (no fallback for general meshes)")
    ij = np.rint(x * nn).astype(np.int64)
    if not np.allclose(ij / nn, x, atol=1e-12):
        raise ValueError("micro mesh vertices are not on a uniform grid")
    T = 2 if d == 2 else 6
    cells = msh.cells
    if len(cells) != T * nn**d:
        raise ValueError("micro mesh is not the standard simplicial split of a grid")
    cv = ij[cells]  # (nc, d+1, d)
    origin = cv.min(axis=1)
    local = cv - origin[:, None, :]
    if local.max() > 1:
        raise ValueError("micro mesh cells span more than one grid cube")
    # identify the type of every cell by its vertex set
    ref = kuhn_vertices(d)
    code = (local * (2 ** np.arange(d))).sum(axis=2)  # corner id per vertex
    key = np.sort(code, axis=1)
    ref_key = np.sort((ref * (2 ** np.arange(d))).sum(axis=2), axis=1)
    types = np.full(len(cells), -1)
    for t in range(T):
        types[(key == ref_key[t]).all(axis=1)] = t
    if (types < 0).any():
        raise ValueError("micro mesh cells are not the right-diagonal / main-diagonal split")
    order = np.zeros((T, d + 1, d), dtype=np.int64)
    for t in range(T):
        sel = np.nonzero(types == t)[0]
        if len(sel) != nn**d:
            raise ValueError("micro mesh does not contain every element type once per cube")
        order[t] = local[sel[0]]
        if not (local[sel] == order[t]).all():
            raise ValueError("micro mesh lists the vertices of equal elements in different orders")
    return MicroStructure(d, nn, order)


def default_structure(dim, n):
    """Structure of ``create_unit_square(n, n)`` / ``create_unit_cube(n, n, n)``."""
    from . import mesh

    base = detect_structure(mesh.create_unit_square(2, 2) if dim == 2 else mesh.create_unit_cube(2, 2, 2))
    return MicroStructure(dim, int(n), base.vertex_order)


def quadrature_table(structure, points, weights):
    """Map a reference-simplex rule to cube-local points per element type.

    Returns ``qp`` (T, nq, dim) in units of h relative to the cube origin and ``qw`` (nq,)
    normalised to sum 1 (element means)."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, structure.dim)
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    v = structure.vertex_order.astype(np.float64)  # (T, d+1, d)
    edges = v[:, 1:, :] - v[:, :1, :]  # (T, d, d): row j = v_{j+1} - v_0
    qp = v[:, :1, :] + np.einsum("qj,tjk->tqk", pts, edges)
    return np.ascontiguousarray(qp), np.ascontiguousarray(w / w.sum())
