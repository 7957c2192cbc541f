"""numpy-backed stand-in for the handful of UFL names the reference's
coefficient callables use (``A(x, y)``, ``Dtheta_transpose(x)``, ``f(x)``:
/root/reference/src/hommx/hmm.py:131,198,757; enumerated from
test/integration/*.py and examples/**/*.py).  Oracle-only: the product has its
own symbolic layer (hommx_b200/ufl.py) and never imports this module.

Conventions: ``x`` is a float array of shape (3,) (``fem.Constant`` of shape
(3,), hmm.py:190-192); ``y`` is an array of shape (gdim, npts) so that
``y[0]`` is the array of first coordinates.  Tensor-valued results carry the
point axis last.
"""
import numpy as np

pi = np.pi
sin, cos, acos, sqrt, exp, ln, tan = np.sin, np.cos, np.arccos, np.sqrt, np.exp, np.log, np.tan


def conditional(cond, a, b):
    return np.where(cond, a, b)


def lt(a, b):
    return a < b


def gt(a, b):
    return a > b


def _pt(v):
    """broadcast to a trailing point axis"""
    v = np.asarray(v, dtype=float)
    return v if v.ndim else v.reshape(1)


def as_vector(comps):
    comps = [_pt(c) for c in comps]
    n = max(c.shape[-1] for c in comps)
    return np.stack([np.broadcast_to(c, (n,)) for c in comps], axis=0)


def as_matrix(rows):
    rows = [[_pt(c) for c in r] for r in rows]
    n = max(c.shape[-1] for r in rows for c in r)
    return np.stack([np.stack([np.broadcast_to(c, (n,)) for c in r], axis=0) for r in rows], axis=0)


def transpose(m):
    m = np.asarray(m)
    return np.swapaxes(m, 0, 1)


class Index:
    _count = 0

    def __init__(self):
        Index._count += 1
        self.id = Index._count


def indices(n):
    return tuple(Index() for _ in range(n))


class Labeled:
    """array with named leading axes + trailing point axis (index notation)."""

    __array_ufunc__ = None  # make ndarray * Labeled defer to Labeled.__rmul__

    def __init__(self, data, labels):
        self.data = data  # shape (*dims, npts|1)
        self.labels = tuple(labels)

    def _aligned(self, other):
        labels = list(self.labels) + [l for l in other.labels if l not in self.labels]

        def expand(t):
            src = list(t.labels)
            d = t.data
            # move existing axes to label order, insert singleton axes for missing labels
            perm = [src.index(l) for l in labels if l in src]
            d = np.transpose(d, perm + [d.ndim - 1])
            shape, it = [], iter(d.shape[:-1])
            for l in labels:
                shape.append(next(it) if l in src else 1)
            return d.reshape(shape + [d.shape[-1]])

        return expand(self), expand(other), labels

    def __mul__(self, other):
        if isinstance(other, Labeled):
            a, b, labels = self._aligned(other)
            return Labeled(a * b, labels)
        return Labeled(self.data * _pt(other), self.labels)

    __rmul__ = __mul__

    def __add__(self, other):
        a, b, labels = self._aligned(other)
        assert set(self.labels) == set(other.labels), "index sets must agree in a sum"
        return Labeled(a + b, labels)

    def __neg__(self):
        return Labeled(-self.data, self.labels)

    def __sub__(self, other):
        return self + (-other)


class _Tensor(np.ndarray):
    def __getitem__(self, idx):
        if isinstance(idx, tuple) and any(isinstance(i, Index) for i in idx):
            return Labeled(np.asarray(self)[..., None], idx)
        return super().__getitem__(idx)


def Identity(d):
    return np.eye(d).view(_Tensor)


def _stack_nested(expr):
    """nested lists of scalars / point arrays -> array with the point axis last (leaves broadcast)"""
    if isinstance(expr, (list, tuple)):
        parts = [_stack_nested(e) for e in expr]
        n = max(p.shape[-1] for p in parts)
        parts = [np.broadcast_to(p, p.shape[:-1] + (n,)) for p in parts]
        return np.stack(parts, axis=0)
    return _pt(expr)


def as_tensor(expr, indices=None):
    if indices is None:
        return _stack_nested(expr) if isinstance(expr, (list, tuple)) else np.asarray(expr)
    src = list(expr.labels)
    perm = [src.index(l) for l in indices]
    return np.transpose(expr.data, perm + [expr.data.ndim - 1])
