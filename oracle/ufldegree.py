"""Quadrature-degree estimation for the cell-problem forms, restated independently of the product.

TEST INFRASTRUCTURE (oracle): only tests/, bench.py's CPU legs and __graft_entry__.smoke() import it.

FFCx integrates each form with the basix default rule of the degree UFL estimates for the integrand
(``ufl.algorithms.estimate_total_polynomial_degree``).  The forms of the hot path
(/root/reference/src/hommx/hmm.py:644-667, 759-789, 891-922, 1032-1067) are products of the coefficient
``A(x_macro, y)`` with gradients of P1 functions (degree 0 on affine simplices) and the constant Jacobian
``Dtheta(x_macro)`` (degree 0), so the degree of every form is the degree of ``A``:

* ``fem.Constant`` (the macro point ``x``) -> 0, ``SpatialCoordinate`` of an affine mesh (``y``) -> 1, numbers -> 0;
* sum -> max, product -> sum, division -> sum (UFL adds numerator and denominator degrees);
* ``sin cos tan acos sqrt exp ln`` -> argument + 2;  ``f ** p``: non-negative integer p -> degree * p, else + 2;
* ``conditional(c, t, f)`` -> max(t, f) (the condition is ignored);  tensors -> max over the components.

This module evaluates the SAME coefficient callables (tests/coefficients.py) with values that carry nothing but
that degree -- a third back end next to numpy (oracle/npufl.py) and the product's tracer (hommx_b200/ufl.py, whose
own estimate lives in ``estimate_degree`` there).  ``form_degree(A, dim)`` is what the oracle's MicroCell is built with.
"""
import math
import numbers

pi = math.pi


class Deg:
    """A scalar expression known only by its polynomial degree in the micro coordinate."""

    def __init__(self, d):
        self.d = int(d)

    @staticmethod
    def of(v):
        if isinstance(v, Deg):
            return v.d
        if isinstance(v, numbers.Number):
            return 0
        if isinstance(v, Tensor):
            return v.degree()
        raise TypeError(f"cannot take the degree of {type(v).__name__}")

    def _sum(self, o):
        return Deg(max(self.d, Deg.of(o)))

    def _prod(self, o):
        return Deg(self.d + Deg.of(o))

    __add__ = __radd__ = __sub__ = __rsub__ = _sum
    __mul__ = __rmul__ = _prod
    __truediv__ = __rtruediv__ = _prod  # UFL: degree(numerator) + degree(denominator)

    def __neg__(self):
        return Deg(self.d)

    def __pow__(self, p):
        if isinstance(p, numbers.Integral) and p >= 0:
            return Deg(self.d * int(p))
        return Deg(self.d + 2)

    def __rpow__(self, base):
        return Deg(self.d + 2)

    # comparisons build a condition, whose degree never matters
    def _cond(self, o):
        return Cond()

    __lt__ = __gt__ = __le__ = __ge__ = _cond


class Cond:
    pass


def _math(v):
    return Deg(Deg.of(v) + 2)


sin = cos = tan = acos = asin = atan = sqrt = exp = ln = _math


def conditional(c, t, f):
    return Deg(max(Deg.of(t), Deg.of(f)))


def lt(a, b):
    return Cond()


gt = le = ge = lt


class Index:
    pass


def indices(n):
    return tuple(Index() for _ in range(n))


class Tensor:
    """Nested lists of Deg / numbers; indexing with Index objects or integers gives the maximal degree reachable."""

    def __init__(self, comps):
        self.comps = comps

    def degree(self):
        def walk(c):
            if isinstance(c, (list, tuple)):
                return max(walk(k) for k in c)
            return Deg.of(c)

        return walk(self.comps)

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        c = self.comps
        for k in idx:
            if isinstance(k, Index):
                return Deg(self.degree())  # a free index: any component
            c = c[k]
        return c if not isinstance(c, (list, tuple)) else Tensor(c)

    @property
    def T(self):
        return self


class Identity(Tensor):
    def __init__(self, dim):
        super().__init__([[1.0 if i == j else 0.0 for j in range(dim)] for i in range(dim)])

    def __getitem__(self, idx):
        return Deg(0)


def as_vector(comps):
    return Tensor(list(comps))


def as_matrix(rows):
    return Tensor([list(r) for r in rows])


def as_tensor(expr, indices=None):
    if indices is not None:
        return Tensor([expr])  # index notation: one representative component carries the degree
    return Tensor(expr)


def transpose(m):
    return m


class Coordinate:
    def __init__(self, degree):
        self._d = degree

    def __getitem__(self, k):
        return Deg(self._d)


def form_degree(A, dim):
    """Estimated quadrature degree of the cell-problem forms for the coefficient callable ``A(x, y)`` (built with this
    module as its ``ufl`` namespace)."""
    x = Coordinate(0)  # fem.Constant of shape (3,), hmm.py:190-192
    y = Coordinate(1)  # ufl.SpatialCoordinate(micro mesh), affine simplices
    return Deg.of(A(x, y))
