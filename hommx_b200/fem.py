"""The few FEM host objects the HMM classes need when DOLFINx is not installed: a P1 function
space on a ``SimplexMesh``, nodal functions, Dirichlet conditions and the macro load vector.

These stand in for ``dolfinx.fem.functionspace`` / ``Function`` / ``dirichletbc`` /
``locate_dofs_geometrical`` as the reference uses them (/root/reference/src/hommx/hmm.py:
124-135, 276-287, 453-480, 598-639).  They are host-side set-up and post-processing, not the hot
path.
"""
from __future__ import annotations

import numpy as np

from . import ufl
from .mesh import as_simplex_mesh
from .quadrature import default_rule


class _Vector:
    def __init__(self, n):
        self.array = np.zeros(n)

    def scatter_forward(self):  # single-process stand-in
        pass


class FunctionSpace:
    """Continuous P1 (``("Lagrange", 1)``) space, scalar (bs=1) or vector (bs=dim)."""

    def __init__(self, msh, bs=1):
        self.mesh = as_simplex_mesh(msh)
        self.bs = int(bs)

    @property
    def num_dofs(self):
        return self.mesh.num_nodes * self.bs

    def tabulate_dof_coordinates(self):
        return self.mesh.x


class Function:
    def __init__(self, V):
        self.function_space = V
        self.x = _Vector(V.num_dofs)

    def copy(self):
        g = Function(self.function_space)
        g.x.array[:] = self.x.array
        return g


class DirichletBC:
    """``value`` is a scalar, a vector of length bs, or a Function; ``dofs`` are blocked (node) ids
    for a full-space condition, or unrolled dof ids when ``unrolled=True`` (sub-space condition)."""

    def __init__(self, value, dofs, V, unrolled=False):
        self.g = value
        self.V = V
        dofs = np.asarray(dofs, dtype=np.int64)
        bs = V.bs
        if unrolled or bs == 1:
            self.dofs = dofs
            if isinstance(value, Function):
                self.values = value.x.array[dofs]
            else:
                self.values = np.broadcast_to(np.asarray(value, dtype=np.float64), dofs.shape).copy()
        else:
            self.dofs = (dofs[:, None] * bs + np.arange(bs)[None, :]).ravel()
            if isinstance(value, Function):
                self.values = value.x.array[self.dofs]
            else:
                v = np.asarray(value, dtype=np.float64)
                self.values = np.tile(v, len(dofs)) if v.ndim else np.full(len(self.dofs), float(v))


def dirichletbc(value, dofs, V, unrolled=False):
    return DirichletBC(value, dofs, V, unrolled)


def locate_dofs_geometrical(V, marker):
    """Blocked dof (node) ids whose coordinates satisfy ``marker(x)`` with x of shape (3, N)."""
    return np.nonzero(np.asarray(marker(V.mesh.x.T), dtype=bool))[0]


def boundary_nodes(msh):
    """Nodes on the bounding box of the mesh (the default condition of PoissonHMM, hmm.py:598-636)."""
    msh = as_simplex_mesh(msh)
    x = msh.x
    on = np.zeros(len(x), dtype=bool)
    for k in range(msh.dim):
        on |= np.isclose(x[:, k], x[:, k].min()) | np.isclose(x[:, k], x[:, k].max())
    return np.nonzero(on)[0]


def assemble_load(V, f):
    """b_i = int f . phi_i dx  (hmm.py:131-133, 445-450).  ``f(x)`` is traced with
    ``hommx_b200.ufl`` and integrated with the rule of the UFL-estimated degree."""
    msh, bs = V.mesh, V.bs
    d = msh.dim
    val = ufl.call_traced(f, [("coord", "x", d)])
    if isinstance(val, ufl.Tensor):
        comps = [val.data[k] for k in range(val.data.shape[0])]
    elif isinstance(val, (list, tuple, np.ndarray)):
        comps = [ufl.Expr.wrap(v) for v in val]
    else:
        comps = [ufl.Expr.wrap(val)]
    if len(comps) != bs:
        raise ValueError(f"f must have {bs} component(s), got {len(comps)}")
    deg = max(ufl.estimate_degree(c, {"x": 1}) for c in comps) + 1  # times the P1 test function
    qp, qw = default_rule(d, deg)
    X = msh.x[:, :d]
    v = X[msh.cells]  # (nc, d+1, d)
    J = np.transpose(v[:, 1:] - v[:, :1], (0, 2, 1))
    detJ = np.abs(np.linalg.det(J))
    phi = np.concatenate([1.0 - qp.sum(axis=1, keepdims=True), qp], axis=1)  # (nq, d+1)
    xq = v[:, :1, :] + np.einsum("eij,qj->eqi", J, qp)
    env = {("x", k): xq[:, :, k] for k in range(d)}
    b = np.zeros(V.num_dofs)
    for k, c in enumerate(comps):
        fv = np.broadcast_to(np.asarray(ufl.evaluate(c, env), dtype=np.float64), xq.shape[:2])
        contrib = np.einsum("eq,q,qa,e->ea", fv, qw, phi, detJ)
        np.add.at(b, msh.cells.astype(np.int64) * bs + k, contrib)
    return b
