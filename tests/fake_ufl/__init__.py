"""A stand-in for UFL's expression classes -- TEST FIXTURE ONLY.

UFL (fenics-ufl) is not installable in the build image, so ``hommx_b200.ufl.from_ufl`` -- the translator that lets a
user script written against the real ``ufl`` module keep working (hmm.py:186-198 hands ``A`` a ``fem.Constant`` and a
``ufl.SpatialCoordinate``) -- is exercised against this module.  It reproduces what the translator relies on and
nothing else: UFL's node CLASS NAMES, ``ufl_operands``, ``ufl_shape``, ``MultiIndex.indices()``, ``Index.count()``,
``FixedIndex.__int__``, ``IndexSum.dimension()``, ``ScalarValue.value()``, and the way UFL's operators build the graph
(``a - b`` = ``Sum(a, Product(-1, b))``, ``A[i, j] * v[j]`` = ``IndexSum(Product(Indexed, Indexed), MultiIndex)``,
``as_tensor(e, (i, j))`` = ``ComponentTensor``, scalar * tensor = ``ComponentTensor(Product(scalar, Indexed))``),
following ufl/exproperators.py, ufl/algebra.py, ufl/tensors.py and ufl/indexsum.py of UFL 2024.2.  No evaluation, no
forms, no differentiation."""
import math
import numbers

pi = math.pi


class Expr:
    ufl_operands = ()
    ufl_shape = ()
    ufl_free_indices = ()  # sorted tuple of Index counts
    ufl_index_dimensions = ()

    # -- operators (ufl/exproperators.py) -----------------------------------
    def __add__(self, o):
        return Sum(self, as_ufl(o))

    def __radd__(self, o):
        return Sum(as_ufl(o), self)

    def __sub__(self, o):
        return Sum(self, -as_ufl(o))

    def __rsub__(self, o):
        return Sum(as_ufl(o), -self)

    def __neg__(self):
        return _mult(IntValue(-1), self)

    def __pos__(self):
        return self

    def __mul__(self, o):
        return _mult(self, as_ufl(o))

    def __rmul__(self, o):
        return _mult(as_ufl(o), self)

    def __truediv__(self, o):
        o = as_ufl(o)
        if self.ufl_shape:
            ii = indices(len(self.ufl_shape))
            return as_tensor(Division(self[ii], o), ii)
        return Division(self, o)

    def __rtruediv__(self, o):
        return Division(as_ufl(o), self)

    def __pow__(self, o):
        return Power(self, as_ufl(o))

    def __rpow__(self, o):
        return Power(as_ufl(o), self)

    def __abs__(self):
        return Abs(self)

    def __lt__(self, o):
        return LT(self, as_ufl(o))

    def __gt__(self, o):
        return GT(self, as_ufl(o))

    def __le__(self, o):
        return LE(self, as_ufl(o))

    def __ge__(self, o):
        return GE(self, as_ufl(o))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        idx = tuple(FixedIndex(i) if isinstance(i, numbers.Integral) else i for i in idx)
        if len(idx) != len(self.ufl_shape):
            raise NotImplementedError("fake_ufl: slices / partial indexing")
        e = Indexed(self, MultiIndex(idx))
        # a repeated free index is summed (ufl/exproperators.py _getitem)
        counts = [i.count() for i in idx if isinstance(i, Index)]
        for c in sorted(set(counts)):
            if counts.count(c) > 1:
                k = [i for i in idx if isinstance(i, Index) and i.count() == c][0]
                e = IndexSum(e, MultiIndex((k,)))
        return e

    @property
    def T(self):
        return transpose(self)

    def __len__(self):
        if len(self.ufl_shape) != 1:
            raise NotImplementedError("len() of a non-vector")
        return self.ufl_shape[0]

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]

    def __bool__(self):
        raise TypeError("UFL conditions have no truth value")


def _free(*ops):
    """Merged free indices and their dimensions of the operands."""
    m = {}
    for o in ops:
        m.update(zip(o.ufl_free_indices, o.ufl_index_dimensions))
    keys = tuple(sorted(m))
    return keys, tuple(m[k] for k in keys)


class Operator(Expr):
    def __init__(self, *ops):
        self.ufl_operands = tuple(ops)
        self.ufl_free_indices, self.ufl_index_dimensions = _free(*ops)


# -- terminals -------------------------------------------------------------------
class ScalarValue(Expr):
    def __init__(self, v):
        self._value = v

    def value(self):
        return self._value


class IntValue(ScalarValue):
    pass


class FloatValue(ScalarValue):
    pass


class Zero(Expr):
    def __init__(self, shape=()):
        self.ufl_shape = tuple(shape)


class Identity(Expr):
    def __init__(self, dim):
        self.ufl_shape = (dim, dim)


class Mesh:
    """ufl.Mesh(coordinate_element): only the geometric dimension is kept."""

    def __init__(self, element):
        name = element if isinstance(element, str) else getattr(element, "cell_name", None) or str(element)
        self._gdim = 2 if "triangle" in name else 3

    def geometric_dimension(self):
        return self._gdim


class SpatialCoordinate(Expr):
    def __init__(self, domain):
        self.ufl_shape = (domain.geometric_dimension(),)


class Constant(Expr):
    def __init__(self, domain, shape=()):
        self.ufl_shape = tuple(shape)


def as_ufl(v):
    if isinstance(v, Expr):
        return v
    if isinstance(v, numbers.Integral):
        return IntValue(int(v))
    if isinstance(v, numbers.Real):
        return FloatValue(float(v))
    raise ValueError(f"Invalid type conversion: {v!r} can not be converted to any UFL type.")


# -- indices -----------------------------------------------------------------------
class FixedIndex:
    def __init__(self, v):
        self._value = int(v)

    def __int__(self):
        return self._value


class Index:
    _next = 0

    def __init__(self):
        Index._next += 1
        self._count = Index._next

    def count(self):
        return self._count


def indices(n):
    return tuple(Index() for _ in range(n))


class MultiIndex(Expr):
    def __init__(self, idx):
        self._indices = tuple(idx)

    def indices(self):
        return self._indices

    def __iter__(self):
        return iter(self._indices)

    def __len__(self):
        return len(self._indices)


class Indexed(Operator):
    def __init__(self, expr, mi):
        self.ufl_operands = (expr, mi)
        m = dict(zip(expr.ufl_free_indices, expr.ufl_index_dimensions))
        for i, d in zip(mi.indices(), expr.ufl_shape):
            if isinstance(i, Index):
                m[i.count()] = d
        keys = tuple(sorted(m))
        self.ufl_free_indices, self.ufl_index_dimensions = keys, tuple(m[k] for k in keys)


class IndexSum(Operator):
    def __init__(self, summand, mi):
        self.ufl_operands = (summand, mi)
        (i,) = mi.indices()
        m = dict(zip(summand.ufl_free_indices, summand.ufl_index_dimensions))
        self._dimension = m.pop(i.count())
        keys = tuple(sorted(m))
        self.ufl_free_indices, self.ufl_index_dimensions = keys, tuple(m[k] for k in keys)
        self.ufl_shape = summand.ufl_shape

    def dimension(self):
        return self._dimension


class ComponentTensor(Operator):
    def __init__(self, expr, mi):
        self.ufl_operands = (expr, mi)
        m = dict(zip(expr.ufl_free_indices, expr.ufl_index_dimensions))
        self.ufl_shape = tuple(m.pop(i.count()) for i in mi.indices())
        keys = tuple(sorted(m))
        self.ufl_free_indices, self.ufl_index_dimensions = keys, tuple(m[k] for k in keys)


class ListTensor(Operator):
    def __init__(self, *ops):
        super().__init__(*ops)
        self.ufl_shape = (len(ops),) + tuple(ops[0].ufl_shape)


def _nest(v):
    if isinstance(v, (list, tuple)):
        return ListTensor(*[_nest(e) for e in v])
    return as_ufl(v)


def as_tensor(expressions, indices=None):  # noqa: A002 - UFL's own parameter name
    if indices is None:
        return expressions if isinstance(expressions, Expr) else _nest(expressions)
    if isinstance(indices, Index):
        indices = (indices,)
    return ComponentTensor(as_ufl(expressions), MultiIndex(tuple(indices)))


as_vector = as_tensor
as_matrix = as_tensor


class Transposed(Operator):
    def __init__(self, a):
        super().__init__(a)
        self.ufl_shape = (a.ufl_shape[1], a.ufl_shape[0])


def transpose(a):
    return Transposed(a)


# -- algebra -------------------------------------------------------------------------
class Sum(Operator):
    def __init__(self, a, b):
        if a.ufl_shape != b.ufl_shape:
            raise ValueError("Can't add expressions with different shapes.")
        super().__init__(a, b)
        self.ufl_shape = a.ufl_shape


class Product(Operator):
    """Scalar-valued product; repeated free indices are summed by ``_mult``."""


class Division(Operator):
    pass


class Power(Operator):
    pass


class Abs(Operator):
    pass


def _mult(a, b):
    """ufl/exproperators.py _mult: scalar * tensor is a ComponentTensor, matrix * vector / matrix * matrix contract the
    inner axis, repeated free indices are summed."""
    s1, s2 = a.ufl_shape, b.ufl_shape
    if s1 and s2:  # matrix-vector / matrix-matrix
        if len(s1) != 2 or len(s2) not in (1, 2):
            raise ValueError("Invalid ranks in product.")
        ii, k = indices(1), Index()
        jj = indices(len(s2) - 1)
        p = _mult(a[ii + (k,)], b[(k,) + jj])
        return as_tensor(p, ii + jj)
    if s1 or s2:
        t, s = (a, b) if s1 else (b, a)
        ii = indices(len(t.ufl_shape))
        return as_tensor(_mult(s, t[ii]), ii)
    p = Product(a, b)
    for c in sorted(set(a.ufl_free_indices) & set(b.ufl_free_indices)):
        k = Index()
        k._count = c
        p = IndexSum(p, MultiIndex((k,)))
    return p


def _math(name):
    cls = type(name, (Operator,), {})

    def f(a):
        return cls(as_ufl(a))

    f.__name__ = name.lower()
    return cls, f


Sin, sin = _math("Sin")
Cos, cos = _math("Cos")
Tan, tan = _math("Tan")
Acos, acos = _math("Acos")
Asin, asin = _math("Asin")
Atan, atan = _math("Atan")
Sqrt, sqrt = _math("Sqrt")
Exp, exp = _math("Exp")
Ln, ln = _math("Ln")


# -- conditions ----------------------------------------------------------------------------
class Condition(Operator):
    pass


LT = type("LT", (Condition,), {})
GT = type("GT", (Condition,), {})
LE = type("LE", (Condition,), {})
GE = type("GE", (Condition,), {})
EQ = type("EQ", (Condition,), {})
NE = type("NE", (Condition,), {})
AndCondition = type("AndCondition", (Condition,), {})
OrCondition = type("OrCondition", (Condition,), {})
NotCondition = type("NotCondition", (Condition,), {})


def lt(a, b):
    return LT(as_ufl(a), as_ufl(b))


def gt(a, b):
    return GT(as_ufl(a), as_ufl(b))


def le(a, b):
    return LE(as_ufl(a), as_ufl(b))


def ge(a, b):
    return GE(as_ufl(a), as_ufl(b))


def eq(a, b):
    return EQ(as_ufl(a), as_ufl(b))


def ne(a, b):
    return NE(as_ufl(a), as_ufl(b))


def And(a, b):
    return AndCondition(a, b)


def Or(a, b):
    return OrCondition(a, b)


def Not(a):
    return NotCondition(a)


class Conditional(Operator):
    def __init__(self, c, t, f):
        if not isinstance(c, Condition):
            raise ValueError("Expecting condition as first argument.")
        super().__init__(c, t, f)
        self.ufl_shape = t.ufl_shape


def conditional(c, t, f):
    return Conditional(c, as_ufl(t), as_ufl(f))


class MinValue(Operator):
    pass


class MaxValue(Operator):
    pass


def min_value(a, b):
    return MinValue(as_ufl(a), as_ufl(b))


def max_value(a, b):
    return MaxValue(as_ufl(a), as_ufl(b))


# -- compound tensor algebra ---------------------------------------------------------------
class Dot(Operator):
    def __init__(self, a, b):
        super().__init__(a, b)
        self.ufl_shape = a.ufl_shape[:-1] + b.ufl_shape[1:]


class Inner(Operator):
    pass


class Outer(Operator):
    def __init__(self, a, b):
        super().__init__(a, b)
        self.ufl_shape = a.ufl_shape + b.ufl_shape


def dot(a, b):
    a, b = as_ufl(a), as_ufl(b)
    return a * b if not a.ufl_shape and not b.ufl_shape else Dot(a, b)


def inner(a, b):
    a, b = as_ufl(a), as_ufl(b)
    return a * b if not a.ufl_shape and not b.ufl_shape else Inner(a, b)


def outer(a, b):
    return Outer(as_ufl(a), as_ufl(b))
