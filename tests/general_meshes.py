"""General (non-structured) periodic micro meshes of the unit box for the element-list kernel tests:
meshes the structured detector of hommx_b200.micro rejects but the reference's periodic constraint accepts
(cell_problem.py:16-35: any mesh whose boundary nodes match)."""
import numpy as np

from hommx_b200 import mesh


def perturbed(dim, n, seed=0, amp=0.3):
    """The structured mesh with every node moved by up to amp * h / 2 per axis: interior nodes freely, face nodes
    within their face and identically on the opposite face (so that the boundary still matches), corners stay."""
    m = mesh.create_unit_square(n, n) if dim == 2 else mesh.create_unit_cube(n, n, n)
    x = m.x.copy()
    ij = np.rint(x[:, :dim] * n).astype(np.int64)
    rng = np.random.default_rng(seed)
    # one displacement per PERIODIC grid node, zero along every axis in which the node lies on the boundary
    disp = rng.uniform(-0.5, 0.5, (n,) * dim + (dim,)) * amp / n
    per = ij % n
    d = disp[tuple(per[:, a] for a in range(dim))]
    on_bnd = (ij == 0) | (ij == n)
    d[on_bnd] = 0.0
    x[:, :dim] += d
    return mesh.SimplexMesh(x, m.cells.copy(), dim)


def flipped_2d(n):
    """n x n squares, the diagonal direction alternating like a chess board (a 'crossed-free union jack' pattern)."""
    xs = np.linspace(0.0, 1.0, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    x = np.stack([X.ravel(), Y.ravel()], axis=1)
    idx = lambda i, j: j * (n + 1) + i  # noqa: E731
    cells = []
    for j in range(n):
        for i in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i, j + 1), idx(i + 1, j + 1)
            if (i + j) % 2 == 0:
                cells += [[a, b, d], [a, d, c]]
            else:
                cells += [[a, b, c], [b, d, c]]
    return mesh.SimplexMesh(x, np.array(cells), 2)


def refined_corner_2d(n):
    """A structured n x n mesh whose first square is split into four triangles around its centre (one extra node,
    mixed element sizes)."""
    m = mesh.create_unit_square(n, n)
    x = m.x.copy()
    cells = [list(c) for c in m.cells]
    X = x[:, :2]
    h = 1.0 / n
    inside = [k for k, c in enumerate(cells) if (X[c].max(axis=0) <= h + 1e-12).all()]
    corner = sorted({v for k in inside for v in cells[k]}, key=lambda v: (X[v][1], X[v][0]))  # a, b, c, d
    cells = [c for k, c in enumerate(cells) if k not in inside]
    centre = len(x)
    x = np.concatenate([x, [[h / 2, h / 2, 0.0]]])
    a, b, c, d = corner
    cells += [[a, b, centre], [b, d, centre], [d, c, centre], [c, a, centre]]
    return mesh.SimplexMesh(x, np.array(cells), 2)
