"""Structured simplicial meshes with DOLFINx's vertex/cell conventions.

Restates what the reference obtains from ``dolfinx.mesh.create_unit_square``,
``create_rectangle`` (DiagonalType.right), ``create_unit_cube`` and
``create_box`` (used at e.g. /root/reference/examples/hmm.py:35-36,
examples/hmm_3d.py:33-35, test/integration/test_integration_poisson.py:76-113).
DOLFINx renumbers cells and vertices afterwards; every comparison in this repo
therefore keys on coordinates, never on indices.
"""
import numpy as np


class Mesh:
    """Plain container: ``x`` (N,3) vertex coordinates (z = 0 in 2-D, as
    ``msh.geometry.x`` in DOLFINx), ``cells`` (n_cells, dim+1) vertex ids."""

    def __init__(self, x, cells, dim, shape, p0, p1):
        self.x = x
        self.cells = cells
        self.dim = dim
        self.shape = tuple(shape)
        self.p0 = np.asarray(p0, float)
        self.p1 = np.asarray(p1, float)


def create_rectangle(p0, p1, n):
    nx, ny = n
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")  # iy major, ix minor
    x = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1)
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (iy * (nx + 1) + ix).ravel()
    v1 = v0 + 1
    v2 = v0 + (nx + 1)
    v3 = v1 + (nx + 1)
    cells = np.empty((2 * nx * ny, 3), dtype=np.int64)
    cells[0::2] = np.stack([v0, v1, v3], axis=1)
    cells[1::2] = np.stack([v0, v2, v3], axis=1)
    return Mesh(x, cells, 2, (nx, ny), p0, p1)


def create_unit_square(nx, ny):
    return create_rectangle([0.0, 0.0], [1.0, 1.0], [nx, ny])


def create_box(p0, p1, n):
    nx, ny, nz = n
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    v0 = (iz * sz + iy * sy + ix * sx).ravel()
    v1, v2, v3 = v0 + sx, v0 + sy, v0 + sx + sy
    v4, v5, v6, v7 = v0 + sz, v1 + sz, v2 + sz, v3 + sz
    tets = [
        (v0, v1, v3, v7),
        (v0, v1, v7, v5),
        (v0, v5, v7, v4),
        (v0, v3, v2, v7),
        (v0, v6, v4, v7),
        (v0, v2, v6, v7),
    ]
    cells = np.empty((6 * nx * ny * nz, 4), dtype=np.int64)
    for k, t in enumerate(tets):
        cells[k::6] = np.stack(t, axis=1)
    return Mesh(x, cells, 3, (nx, ny, nz), p0, p1)


def create_unit_cube(nx, ny, nz):
    return create_box([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [nx, ny, nz])


def periodic_master_map(mesh):
    """Slave -> master vertex map of the unit box.

    Restates /root/reference/src/hommx/cell_problem.py:38-300: every vertex on
    a max-face is identified with the vertex obtained by moving each
    max-coordinate to the min-coordinate (faces -> opposite face, doubly
    constrained edges -> the min/min edge, far corner -> origin).  Returns
    ``master`` (N,) with master[v] == v for unconstrained vertices.
    """
    x = mesh.x
    d = mesh.dim
    lo = x.min(axis=0)
    hi = x.max(axis=0)
    target = x.copy()
    for k in range(d):
        on_max = np.isclose(x[:, k], hi[k])
        target[on_max, k] = lo[k]
    # look up by rounded coordinates
    key = {tuple(np.round(p, 12)): i for i, p in enumerate(x)}
    master = np.array([key[tuple(np.round(t, 12))] for t in target], dtype=np.int64)
    return master
