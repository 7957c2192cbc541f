"""Structured simplicial meshes in DOLFINx's layout, for hosts without DOLFINx.

The reference takes ``dolfinx.mesh.Mesh`` objects (``msh``, ``msh_micro``:
/root/reference/src/hommx/hmm.py:63-69; built with ``create_unit_square`` /
``create_rectangle`` / ``create_unit_cube`` / ``create_box`` in its tests and examples,
e.g. examples/hmm.py:35-36, examples/hmm_3d.py:33-35).  DOLFINx is not installable in the
build image, so these generators produce the same vertex coordinates and cell splits
(right diagonal in 2-D, the six tetrahedra around the main diagonal in 3-D) in a plain
container exposing the few attributes the HMM classes read.  A real ``dolfinx.mesh.Mesh``
is accepted wherever a ``SimplexMesh`` is (see ``as_simplex_mesh``).
"""
from __future__ import annotations

import numpy as np


class _Geometry:
    def __init__(self, x, dofmap):
        self.x = x  # (N, 3), z = 0 in 2-D like msh.geometry.x
        self.dofmap = dofmap  # (n_cells, dim+1) vertex ids per cell


class _Topology:
    def __init__(self, dim):
        self.dim = dim


class SimplexMesh:
    """P1 simplex mesh: ``geometry.x`` (N,3), ``geometry.dofmap`` (n_cells, dim+1), ``topology.dim``."""

    def __init__(self, x, cells, dim):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[1] == 2:
            x = np.concatenate([x, np.zeros((len(x), 1))], axis=1)
        self.geometry = _Geometry(x, np.ascontiguousarray(cells, dtype=np.int32))
        self.topology = _Topology(int(dim))

    # shorthands
    @property
    def x(self):
        return self.geometry.x

    @property
    def cells(self):
        return self.geometry.dofmap

    @property
    def dim(self):
        return self.topology.dim

    @property
    def num_cells(self):
        return len(self.geometry.dofmap)

    @property
    def num_nodes(self):
        return len(self.geometry.x)


def as_simplex_mesh(msh):
    """Accept a SimplexMesh or anything with DOLFINx's ``geometry.x`` / ``geometry.dofmap`` /
    ``topology.dim`` (P1 geometry)."""
    if isinstance(msh, SimplexMesh):
        return msh
    try:
        x = np.asarray(msh.geometry.x)
        cells = np.asarray(msh.geometry.dofmap)
        dim = int(msh.topology.dim)
    except AttributeError as e:
        raise TypeError("expected a SimplexMesh or a dolfinx.mesh.Mesh") from e
    return SimplexMesh(x, cells, dim)


def _lattice(p0, p1, n):
    axes = [np.linspace(float(a), float(b), int(k) + 1) for a, b, k in zip(p0, p1, n)]
    grids = np.meshgrid(*axes, indexing="ij")  # axis 0 = x fastest handled by the transpose below
    # x varies fastest, then y, then z
    return np.stack([g.transpose(*reversed(range(len(n)))).ravel() for g in grids], axis=1)


def create_rectangle(p0, p1, n):
    """Right-diagonal split: each square (v0 v1 / v2 v3) gives triangles (v0,v1,v3), (v0,v2,v3)."""
    nx, ny = int(n[0]), int(n[1])
    x = _lattice(p0, p1, (nx, ny))
    j, i = np.divmod(np.arange(nx * ny), nx)
    v0 = j * (nx + 1) + i
    quad = np.stack([v0, v0 + 1, v0 + nx + 1, v0 + nx + 2], axis=1)
    cells = quad[:, [[0, 1, 3], [0, 2, 3]]].reshape(-1, 3)
    return SimplexMesh(x, cells, 2)


def create_unit_square(nx, ny):
    return create_rectangle((0.0, 0.0), (1.0, 1.0), (nx, ny))


_TETS = ((0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7))


def create_box(p0, p1, n):
    """Six tetrahedra per hexahedron, all containing the diagonal v0-v7."""
    nx, ny, nz = (int(k) for k in n)
    x = _lattice(p0, p1, (nx, ny, nz))
    c = np.arange(nx * ny * nz)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    sy, sz = nx + 1, (nx + 1) * (ny + 1)
    v0 = k * sz + j * sy + i
    corner = np.stack([v0, v0 + 1, v0 + sy, v0 + sy + 1, v0 + sz, v0 + sz + 1, v0 + sz + sy, v0 + sz + sy + 1], axis=1)
    cells = corner[:, np.array(_TETS)].reshape(-1, 4)
    return SimplexMesh(x, cells, 3)


def create_unit_cube(nx, ny, nz):
    return create_box((0.0, 0.0, 0.0), (1.0, 1.0, 1.0), (nx, ny, nz))
