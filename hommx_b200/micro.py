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


# ----------------------------------------------------------------------------
# general periodic micro meshes: tables of the element-list kernel
# ----------------------------------------------------------------------------
def periodic_node_map(msh):
    """(node -> periodic node id, number of periodic nodes) of a simplicial mesh of the unit box whose boundary nodes
    match: every node on a max-face is identified with the node obtained by moving each max-coordinate to the
    min-coordinate (cell_problem.py:38-300 of the reference: faces -> opposite face, edges -> the min/min edge, far
    corner -> origin).  Raises ValueError when a boundary node has no partner."""
    d = msh.dim
    x = msh.x[:, :d]
    lo, hi = x.min(axis=0), x.max(axis=0)
    if not (np.allclose(lo, 0.0) and np.allclose(hi, 1.0)):
        raise ValueError("the micro mesh must be the unit box [0,1]^d")
    target = x.copy()
    for k in range(d):
        target[np.isclose(x[:, k], hi[k]), k] = lo[k]
    key = {tuple(np.round(p, 10)): i for i, p in enumerate(x)}
    try:
        master = np.array([key[tuple(np.round(t, 10))] for t in target], dtype=np.int64)
    except KeyError as e:
        raise ValueError(f"micro mesh boundary nodes do not match periodically (no partner for {e.args[0]})") from None
    uniq, inv = np.unique(master, return_inverse=True)
    return inv.astype(np.int32), len(uniq)


class ElementListTables:
    """Arrays of ``struct MicroMesh`` (csrc/hmx_cell_common.cuh) for a general periodic micro mesh."""

    def __init__(self, msh_micro, points, weights):
        msh = as_simplex_mesh(msh_micro)
        d = self.dim = msh.dim
        if d not in (2, 3):
            raise ValueError("Topology should be 3D or 2D")  # hmm.py:104-105
        x = msh.x[:, :d]
        cells = np.asarray(msh.cells, dtype=np.int64)
        self.node2per, self.n_nodes = periodic_node_map(msh)
        nv = d + 1
        v = x[cells]  # (E, d+1, d)
        J = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))  # columns = edge vectors
        det = np.linalg.det(J)
        if (np.abs(det) < 1e-14).any():
            raise ValueError("degenerate micro mesh cell")
        self.elem_vol = np.ascontiguousarray(np.abs(det) / (2.0 if d == 2 else 6.0))
        Jinv = np.linalg.inv(J)
        grad = np.empty((len(cells), nv, d))
        grad[:, 1:, :] = Jinv
        grad[:, 0, :] = -Jinv.sum(axis=1)
        self.elem_grad = np.ascontiguousarray(grad)
        pts = np.asarray(points, dtype=np.float64).reshape(-1, d)
        w = np.asarray(weights, dtype=np.float64).reshape(-1)
        self.qw = np.ascontiguousarray(w / w.sum())
        self.nq = len(w)
        self.elem_yq = np.ascontiguousarray(v[:, :1, :] + np.einsum("eij,qj->eqi", J, pts))
        pn = self.node2per[cells].astype(np.int64)  # (E, d+1)
        self.elem_nodes = np.ascontiguousarray(pn, dtype=np.int32)
        self.n_elem = E = len(cells)
        # block pattern and the fixed-order contribution lists
        e_idx, a_idx, b_idx = np.meshgrid(np.arange(E), np.arange(nv), np.arange(nv), indexing="ij")
        rows, cols = pn[e_idx, a_idx].ravel(), pn[e_idx, b_idx].ravel()
        src = ((e_idx * nv + a_idx) * nv + b_idx).ravel()
        order = np.lexsort((src, cols, rows))
        rows, cols, src = rows[order], cols[order], src[order]
        first = np.ones(len(rows), dtype=bool)
        first[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
        starts = np.nonzero(first)[0]
        self.nnzb = len(starts)
        self.col = np.ascontiguousarray(cols[starts], dtype=np.int32)
        self.blk_ptr = np.ascontiguousarray(np.append(starts, len(rows)), dtype=np.int32)
        self.blk_src = np.ascontiguousarray(src, dtype=np.int32)
        brow = rows[starts]
        self.row_ptr = np.ascontiguousarray(np.searchsorted(brow, np.arange(self.n_nodes + 1)), dtype=np.int32)
        self.diag = np.ascontiguousarray(np.nonzero(brow == cols[starts])[0], dtype=np.int32)
        if len(self.diag) != self.n_nodes:
            raise ValueError("micro mesh has nodes that belong to no cell")
        nsrc = (np.arange(E)[:, None] * nv + np.arange(nv)[None, :]).ravel()
        nrow = pn.ravel()
        order = np.lexsort((nsrc, nrow))
        self.node_src = np.ascontiguousarray(nsrc[order], dtype=np.int32)
        self.node_ptr = np.ascontiguousarray(np.searchsorted(nrow[order], np.arange(self.n_nodes + 1)), dtype=np.int32)

    def scratch_doubles(self, natoms, bs, nrhs):
        """Per-CTA global scratch of the element-list kernel (mirrors ``element_list_scratch``)."""
        return max(1, natoms) * self.n_elem + self.nnzb * bs * bs + self.n_nodes * bs * bs + 5 * nrhs * self.n_nodes * bs
